"""Element tables and parameter names needed by the hot path.

Only what the in-scope feature classes read: atomic mass and Pauling electronegativity per
atomic number (same values as the tables the reference looks weights up in,
/root/reference/nearl/constants.py:92-138 -- they are physical data; tests/test_host_api.py
compares them entry by entry when the reference tree is present), the element symbols, the
twenty residue names (constants.py:211-216) and the list of parameters a Feature inherits
from its Featurizer (constants.py:329-338).  Tables are stored as whitespace-separated rows
indexed by atomic number 1..118; "-" marks an element without a tabulated value.
"""

_SYMBOLS = """
  H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar
  K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr
  Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe
  Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu Hf
  Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn Fr Ra Ac Th
  Pa U Np Pu Am Cm Bk Cf Es Fm - - - - - - - -
  - - - - - - - - - -
"""

_MASS = """
  1.008 4.003 6.940 9.012 10.810 12.011 14.007 15.999 18.998 20.180
  22.990 24.305 26.982 28.085 30.974 32.060 35.450 39.948 39.098 40.078
  44.956 47.867 50.941 51.996 54.938 55.845 58.933 58.693 63.546 65.380
  69.723 72.630 74.922 78.971 79.904 83.798 85.468 87.620 88.906 91.224
  92.906 95.950 97.000 101.070 102.906 106.420 107.868 112.414 114.818 118.710
  121.760 127.600 126.904 131.293 132.905 137.327 138.905 140.116 140.908 144.242
  145.000 150.360 151.964 157.250 158.925 162.500 164.930 167.259 168.934 173.045
  174.967 178.490 180.948 183.840 186.207 190.230 192.217 195.084 196.967 200.592
  204.380 207.200 208.980 209.000 210.000 222.000 223.000 226.000 227.000 232.038
  231.036 238.029 237.000 244.000 243.000 247.000 247.000 251.000 252.000 257.000
  258.000 259.000 262.000 267.000 270.000 269.000 270.000 270.000 278.000 281.000
  281.000 285.000 286.000 289.000 289.000 293.000 293.000 294.000
"""

_ELECTRONEGATIVITY = """
  2.20 - 0.98 1.57 2.04 2.55 3.04 3.44 3.98 - 0.93 1.31
  1.61 1.90 2.19 2.58 3.16 - 0.82 1.00 1.36 1.54 1.63 1.66
  1.55 1.83 1.88 1.91 1.90 1.65 1.81 2.01 2.18 2.55 2.96 3.00
  0.82 0.95 1.22 1.33 1.60 2.16 1.90 2.20 2.28 2.20 1.93 1.69
  1.78 1.96 2.05 2.10 2.66 2.60 0.79 0.89 1.10 1.12 1.13 1.14
  - 1.17 - 1.20 - 1.22 1.23 1.24 1.25 - 1.27 1.30
  1.50 2.36 1.90 2.20 2.20 2.28 2.54 2.00 1.62 2.33 2.02 2.00
  2.20 - - - - - - - - - - -
  - - - - - - - - - - - -
  - - - - - - - - - -
"""


def _table(text, conv):
  return {z + 1: conv(tok) for z, tok in enumerate(text.split()) if tok != "-"}


ATOMICNR2SYMBOL = _table(_SYMBOLS, str)
SYMBOL2ATOMICNR = {sym: z for z, sym in ATOMICNR2SYMBOL.items()}
ATOMICMASS = _table(_MASS, float)
ELECNEG = _table(_ELECTRONEGATIVITY, float)

RES = "ALA ARG ASN ASP CYS GLU GLN GLY HIS ILE LEU LYS MET PHE PRO SER THR TRP TYR VAL".split()

# residue-classification labels of Featurizer.loop_by_residue (nearl/constants.py:218-237): single residues
# are numbered in RES order, residue pairs row-major over RES x RES (nearl/constants.py:240-325);
# protonation / disulfide variants are folded onto their parent residue
LAB2RES = {i: name for i, name in enumerate(RES)}
RES2LAB = {name: i for i, name in enumerate(RES)}
RES_PATCH = {
  "CYX": "CYS",
  "HID": "HIS", "HIE": "HIS", "HIP": "HIS",
  "HSD": "HIS", "HSE": "HIS", "HSP": "HIS",
}
RES2LAB_DUAL = {a + b: i * len(RES) + j for i, a in enumerate(RES) for j, b in enumerate(RES)}
LAB2RES_DUAL = {lab: name for name, lab in RES2LAB_DUAL.items()}

# superseded PDB entries (constants.py:327)
PDBCODE_SUPERCEDES = {"4dgo": "6qs5", "4otw": "6op9", "4v1c": "6iso", "5v8h": "6nfg", "5v8j": "6nfo", "6fim": "6fex", "6h7k": "6ibl",
                      "3m8t": "5wcm", "4n3l": "6eo8"}

# parameters a Feature inherits from the Featurizer when it does not set them itself
COMMON_FEATURE_PARMS = [
  "outfile",
  "cutoff", "sigma", "padding", "byres", "frame_offset",
  "hdf_compress_level",
  "frame_slice",
]

# backbone atom names used by Backbone / backboneness / sidechainness (features.py:846,1214,1219)
BACKBONE_ATOM_NAMES = ("C", "O", "CA", "HA", "N", "HN")
