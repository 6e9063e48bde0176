// nearl_b200 -- data-layout kernels in front of the gather kernels:
//   nrl_pack_legacy_kernel  [F,N,3] + [F*N] arrays of the reference's FFI -> packed float4
//   nrl_crop_pack_kernel    focus-box transform + crop + (by-residue) + selection + order-
//                           preserving compaction of raw trajectory frames, ON THE DEVICE
//                           (replaces the numpy host code of Feature.query / crop,
//                           /root/reference/nearl/features.py:92-123,426-483,1329-1368)
//   nrl_atom_norm_kernel    per-atom normalisation sum of the voxeliser (the first pass of
//                           frame_interp_global, /root/reference/src/voxelize.cu:79-108)
#pragma once
#include "common.cuh"
#include <cooperative_groups.h>
#include <float.h>

// ---- legacy layout ------------------------------------------------------------------
// perm (optional): for every frame a permutation of [0,N) that orders the frame's atoms by
// weight value (needed by distinct_count); hide_placeholder_weights: atoms whose weight is
// 99999.0f can never be counted by distinct_count (marching_observers.cu:118-138), they
// are moved out of every tile's reach.
__global__ void nrl_pack_legacy_kernel(const float *__restrict__ coords, const float *__restrict__ weights,
                                       const int *__restrict__ perm, const int F, const int N,
                                       const int hide_placeholder_weights, float4 *__restrict__ atoms,
                                       long long *__restrict__ off, int *__restrict__ cnt, int *__restrict__ pbase) {
  const long long total = (long long)F * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int f = (int)(i / N);
    const int j = (int)(i - (long long)f * N);
    const long long src = (long long)f * N + (perm ? perm[i] : j);
    float x = coords[src * 3], y = coords[src * 3 + 1], z = coords[src * 3 + 2];
    if (hide_placeholder_weights && weights[src] == NRL_PLACEHOLDER) x = FLT_MAX;
    atoms[i] = make_float4(x, y, z, __int_as_float((int)src));
    if (j == 0) { off[f] = (long long)f * N; cnt[f] = N; pbase[f] = f * N; }
  }
  if (N == 0)
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F; f += gridDim.x * blockDim.x) { off[f] = 0; cnt[f] = 0; pbase[f] = 0; }
}

// ---- device-side focus box -----------------------------------------------------------
struct NrlCropDev {
  float focal[3];           // FOCALPOINTS entry (float32, featurizer.py:352)
  float center[3];          // Feature.center (float32, features.py:312)
  float half_f32;           // float32(spacing/2): numpy weak-scalar rule for float32 coordinates
  double half_f64;          // spacing/2 for float64 coordinates
  double lo[3], hi[3];      // keep iff lo < x' < hi (features.py:113-122), thresholds as numpy compares them
  int byres;                // features.py:471-475
  int selection;            // row of the selection-mask table, -1: all atoms
  int max_atoms;            // DynamicFeature.MAX_ALLOWED_ATOMS (features.py:1279,1353-1359), 0: no cap
  int sort_table;           // row of the value-order permutation table, -1: topology order
  int atom_begin, atom_end; // slice of the topology this crop looks at
  int table_shift;          // per-atom tables are read at a + table_shift
  double focal64[3];        // focal point for float64 coordinates
  float lo32[3], hi32[3];   // float32 coordinates: the largest float <= lo / the smallest float >= hi -- for a float x,
                            // (double)x > lo <=> x > lo32 and (double)x < hi <=> x < hi32 (no FP64 on the streaming path)
};

struct NrlCropParams {
  const void *xyz;              // [F_all, N_all, 3] float32 or float64
  int n_atoms;
  const NrlCropDev *crops;
  const int *frame_crop;        // per output frame: crop index
  const int *frame_src;         // per output frame: trajectory frame index
  const unsigned char *sel;     // [n_sel, N_all]
  const int *resid;             // [N_all] dense residue index
  int n_res_words;
  const int *perms;             // [n_tabs, N_all]
  float4 *atoms;                // output, frame fd starts at frame_off[fd]
  const long long *frame_off;   // per output frame: first packed atom (capacity = the crop's slice length)
  long long *off;
  int *cnt;
  int *pbase;                   // per output frame: payload base (atom_begin + table_shift)
  unsigned char *keep;          // unused (kept for ABI stability of the struct)
  unsigned int *resflags;       // scratch [n_frames_out, n_res_words], zero-initialised
};

template <typename T> struct NrlXform;
template <> struct NrlXform<float> {
  // numpy: ((x - focal) + center) - float32(spacing/2), every step rounded to float32
  static __device__ __forceinline__ float apply(float x, int k, const NrlCropDev &cr) {
    return ((x - cr.focal[k]) + cr.center[k]) - cr.half_f32;
  }
};
template <> struct NrlXform<double> {
  // numpy: float64 coordinates promote the float32 focal point and centre exactly
  static __device__ __forceinline__ double apply(double x, int k, const NrlCropDev &cr) {
    return ((x - cr.focal64[k]) + (double)cr.center[k]) - cr.half_f64;
  }
};

#define NRL_CROP_NT 1024
#define NRL_CROP_NW (NRL_CROP_NT / 32)

// blockDim.x: any multiple of 32 up to NRL_CROP_NT (short slices of the topology get small CTAs)
template <typename T>
__global__ void __launch_bounds__(NRL_CROP_NT) nrl_crop_pack_kernel(const NrlCropParams p) {
  extern __shared__ unsigned int s_keep[];            // one bit per atom of the slice
  __shared__ int s_wcnt[NRL_CROP_NW];
  const int fd = blockIdx.x;
  const NrlCropDev &cr = p.crops[p.frame_crop[fd]];
  const int N = p.n_atoms;
  const int a0 = cr.atom_begin, n = cr.atom_end - cr.atom_begin;   // the slice [a0, a0 + n) of the topology
  const int shift = cr.table_shift;
  const T *__restrict__ xyz = reinterpret_cast<const T *>(p.xyz) + ((long long)p.frame_src[fd] * N + a0) * 3;
  unsigned int *__restrict__ resflags = p.resflags + (long long)fd * p.n_res_words;
  const unsigned char *__restrict__ sel = cr.selection >= 0 ? p.sel + (long long)cr.selection * N + a0 + shift : nullptr;
  const int *__restrict__ resid = p.resid + a0 + shift;
  const int nt = blockDim.x, nw = nt >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_pad = (n + 31) & ~31;

  // phase A: transform + strict two-sided crop per axis, in the coordinate dtype; a warp handles 32
  // consecutive atoms (384 contiguous bytes of float32 coordinates) and ballots their flags into one word
  // four atoms per thread and step: all twelve coordinate loads are issued before the first ballot (the
  // ballots keep the compiler from overlapping the loads of consecutive iterations on its own)
  for (int a0 = tid; a0 < n_pad; a0 += 4 * nt) {
    T cx[4], cy[4], cz[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = a0 + u * nt;
      const bool live = a < n;
      cx[u] = live ? xyz[3 * a] : (T)0; cy[u] = live ? xyz[3 * a + 1] : (T)0; cz[u] = live ? xyz[3 * a + 2] : (T)0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = a0 + u * nt;
      if (a >= n_pad) break;                  // warp-uniform: n_pad and the strides are multiples of 32
      bool in = false;
      if (a < n) {
        const T x = NrlXform<T>::apply(cx[u], 0, cr);
        const T y = NrlXform<T>::apply(cy[u], 1, cr);
        const T z = NrlXform<T>::apply(cz[u], 2, cr);
        in = ((double)x < cr.hi[0]) && ((double)x > cr.lo[0]) && ((double)y < cr.hi[1]) &&
             ((double)y > cr.lo[1]) && ((double)z < cr.hi[2]) && ((double)z > cr.lo[2]);
        if (cr.byres && in) {
          const int r = resid[a];
          atomicOr(&resflags[r >> 5], 1u << (r & 31));
        }
        if (!cr.byres && sel) in = in && (sel[a] != 0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (lane == 0) s_keep[a >> 5] = m;
    }
  }
  __syncthreads();
  if (cr.byres) {
    // whole residues of any kept atom (features.py:471-475), then the selection mask
    for (int a = tid; a < n_pad; a += nt) {
      bool k = false;
      if (a < n) {
        const int r = resid[a];
        k = (resflags[r >> 5] >> (r & 31)) & 1u;
        if (sel) k = k && (sel[a] != 0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (lane == 0) s_keep[a >> 5] = m;
    }
    __syncthreads();
  }
  // value-ordered emission: the permutation lists the atoms of the WHOLE topology (slices are not supported here)
  const int *__restrict__ perm = cr.sort_table >= 0 ? p.perms + (long long)cr.sort_table * N : nullptr;
  if (perm && cr.max_atoms > 0) {
    // the cap keeps the first max_atoms atoms in TOPOLOGY order; resolve it before the
    // value-ordered emission.  One warp, sequential words: this path is rare.
    if (warp == 0) {
      int seen = 0;
      for (int w0 = 0; w0 < (n_pad >> 5); w0 += 32) {
        const int wi = w0 + lane;
        unsigned m = wi < (n_pad >> 5) ? s_keep[wi] : 0u;
        int c = __popc(m), incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        int before = seen + incl - c;
        // drop the bits whose topology rank reaches the cap
        while (m && before + __popc(m) > cr.max_atoms) m &= ~(0x80000000u >> __clz(m));
        if (wi < (n_pad >> 5)) s_keep[wi] = m;
        seen += __shfl_sync(0xffffffffu, incl, 31);
      }
    }
    __syncthreads();
  }

  // phase B: order-preserving compaction in emission order, warp-contiguous segments
  const int seg = ((n + nw - 1) / nw + 31) & ~31;
  const int e_begin = min(warp * seg, n), e_end = min(e_begin + seg, n);
  int mycnt = 0;
  if (!perm) {
    if (e_begin < e_end)
      for (int w0 = (e_begin >> 5) + lane; w0 < ((e_end + 31) >> 5); w0 += 32) mycnt += __popc(s_keep[w0]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mycnt += __shfl_xor_sync(0xffffffffu, mycnt, o);
  } else {
    for (int e0 = e_begin; e0 < e_end; e0 += 32) {
      const int e = e0 + lane;
      bool k = false;
      if (e < e_end) { const int a = perm[e]; k = (s_keep[a >> 5] >> (a & 31)) & 1u; }
      mycnt += __popc(__ballot_sync(0xffffffffu, k));
    }
  }
  if (lane == 0) s_wcnt[warp] = mycnt;
  __syncthreads();
  int base = 0, total = 0;
  for (int w = 0; w < nw; ++w) { const int c = s_wcnt[w]; if (w < warp) base += c; total += c; }
  const int cap = (cr.max_atoms > 0 && !perm) ? cr.max_atoms : 0x7fffffff;
  const long long foff = p.frame_off[fd];
  float4 *out = p.atoms + foff;
  for (int e0 = e_begin; e0 < e_end; e0 += 32) {
    const int e = e0 + lane;
    bool k = false;
    int a = 0;
    unsigned m;
    if (!perm) {
      m = s_keep[e0 >> 5];                    // segments start on word boundaries
      k = (m >> lane) & 1u;
      a = e;
    } else {
      if (e < e_end) { a = perm[e]; k = (s_keep[a >> 5] >> (a & 31)) & 1u; }
      m = __ballot_sync(0xffffffffu, k);
    }
    if (m == 0u) continue;
    const int pos = base + __popc(m & ((1u << lane) - 1u));
    if (k && pos < cap) {
      const T x = NrlXform<T>::apply(xyz[3 * a], 0, cr);
      const T y = NrlXform<T>::apply(xyz[3 * a + 1], 1, cr);
      const T z = NrlXform<T>::apply(xyz[3 * a + 2], 2, cr);
      out[pos] = make_float4((float)x, (float)y, (float)z, __int_as_float(a0 + a + shift));   // commands.*: astype(float32)
    }
    base += __popc(m);
  }
  if (tid == 0) { p.off[fd] = foff; p.cnt[fd] = min(total, cap); p.pbase[fd] = a0 + shift; }
}

// Cluster form of the crop kernel for whole-topology crops in topology order: NRL_CROP_CLUSTER CTAs (one
// thread-block cluster) share a frame, each owns a contiguous chunk of the atoms, and the kept counts are
// exchanged through distributed shared memory so that the compaction stays order preserving.  One CTA per
// frame leaves most SMs idle when a trajectory has ~100 frames of 20 000 atoms; four per frame do not.
#define NRL_CROP_CLUSTER 4
#define NRL_CROP_CNT 512

template <typename T>
__global__ void __cluster_dims__(NRL_CROP_CLUSTER, 1, 1) __launch_bounds__(NRL_CROP_CNT)
nrl_crop_pack_cluster_kernel(const NrlCropParams p) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ unsigned int s_keep[];            // one bit per atom of this CTA's chunk
  __shared__ int s_wcnt[NRL_CROP_CNT / 32];
  __shared__ int s_cta_total;
  const int fd = blockIdx.x / NRL_CROP_CLUSTER;
  const int rank = (int)cluster.block_rank();
  const NrlCropDev &cr = p.crops[p.frame_crop[fd]];
  const int N = p.n_atoms;
  const int a0 = cr.atom_begin, n_all = cr.atom_end - cr.atom_begin;
  const int shift = cr.table_shift;
  // chunk of this CTA, on 32-atom boundaries
  const int chunk = (((n_all + NRL_CROP_CLUSTER - 1) / NRL_CROP_CLUSTER) + 31) & ~31;
  const int c0 = min(rank * chunk, n_all), n = min(c0 + chunk, n_all) - c0;
  const T *__restrict__ xyz = reinterpret_cast<const T *>(p.xyz) + ((long long)p.frame_src[fd] * N + a0 + c0) * 3;
  unsigned int *__restrict__ resflags = p.resflags + (long long)fd * p.n_res_words;
  const unsigned char *__restrict__ sel = cr.selection >= 0 ? p.sel + (long long)cr.selection * N + a0 + c0 + shift : nullptr;
  const int *__restrict__ resid = p.resid + a0 + c0 + shift;
  constexpr int nt = NRL_CROP_CNT, nw = nt >> 5;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_pad = (n + 31) & ~31;

  // four atoms per thread and step: all twelve coordinate loads are issued before the first ballot (the
  // ballots keep the compiler from overlapping the loads of consecutive iterations on its own)
  for (int a0 = tid; a0 < n_pad; a0 += 4 * nt) {
    T cx[4], cy[4], cz[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = a0 + u * nt;
      const bool live = a < n;
      cx[u] = live ? xyz[3 * a] : (T)0; cy[u] = live ? xyz[3 * a + 1] : (T)0; cz[u] = live ? xyz[3 * a + 2] : (T)0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int a = a0 + u * nt;
      if (a >= n_pad) break;                  // warp-uniform: n_pad and the strides are multiples of 32
      bool in = false;
      if (a < n) {
        const T x = NrlXform<T>::apply(cx[u], 0, cr);
        const T y = NrlXform<T>::apply(cy[u], 1, cr);
        const T z = NrlXform<T>::apply(cz[u], 2, cr);
        in = ((double)x < cr.hi[0]) && ((double)x > cr.lo[0]) && ((double)y < cr.hi[1]) &&
             ((double)y > cr.lo[1]) && ((double)z < cr.hi[2]) && ((double)z > cr.lo[2]);
        if (cr.byres && in) {
          const int r = resid[a];
          atomicOr(&resflags[r >> 5], 1u << (r & 31));
        }
        if (!cr.byres && sel) in = in && (sel[a] != 0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, in);
      if (lane == 0) s_keep[a >> 5] = m;
    }
  }
  if (cr.byres) {
    // a residue can straddle two chunks: every CTA of the cluster must have flagged its residues first
    __threadfence();
    cluster.sync();
    for (int a = tid; a < n_pad; a += nt) {
      bool k = false;
      if (a < n) {
        const int r = resid[a];
        k = (__ldcg(&resflags[r >> 5]) >> (r & 31)) & 1u;
        if (sel) k = k && (sel[a] != 0);
      }
      const unsigned m = __ballot_sync(0xffffffffu, k);
      if (lane == 0) s_keep[a >> 5] = m;
    }
  }
  __syncthreads();

  // order-preserving compaction: warp-contiguous segments inside the chunk, chunks in rank order
  const int seg = ((n + nw - 1) / nw + 31) & ~31;
  const int e_begin = min(warp * seg, n), e_end = min(e_begin + seg, n);
  int mycnt = 0;
  if (e_begin < e_end)
    for (int w0 = (e_begin >> 5) + lane; w0 < ((e_end + 31) >> 5); w0 += 32) mycnt += __popc(s_keep[w0]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mycnt += __shfl_xor_sync(0xffffffffu, mycnt, o);
  if (lane == 0) s_wcnt[warp] = mycnt;
  __syncthreads();
  int base = 0, cta_total = 0;
#pragma unroll
  for (int w = 0; w < nw; ++w) { const int c = s_wcnt[w]; if (w < warp) base += c; cta_total += c; }
  if (tid == 0) s_cta_total = cta_total;
  cluster.sync();
  int total = 0;
#pragma unroll
  for (int q = 0; q < NRL_CROP_CLUSTER; ++q) {
    const int c = *cluster.map_shared_rank(&s_cta_total, q);
    if (q < rank) base += c;
    total += c;
  }
  const int cap = cr.max_atoms > 0 ? cr.max_atoms : 0x7fffffff;
  const long long foff = p.frame_off[fd];
  float4 *out = p.atoms + foff;
  for (int e0 = e_begin; e0 < e_end; e0 += 32) {
    const unsigned m = s_keep[e0 >> 5];
    if (m == 0u) continue;
    const bool k = (m >> lane) & 1u;
    const int a = e0 + lane;
    const int pos = base + __popc(m & ((1u << lane) - 1u));
    if (k && pos < cap) {
      const T x = NrlXform<T>::apply(xyz[3 * a], 0, cr);
      const T y = NrlXform<T>::apply(xyz[3 * a + 1], 1, cr);
      const T z = NrlXform<T>::apply(xyz[3 * a + 2], 2, cr);
      out[pos] = make_float4((float)x, (float)y, (float)z, __int_as_float(a0 + c0 + a + shift));
    }
    base += __popc(m);
  }
  if (rank == 0 && tid == 0) { p.off[fd] = foff; p.cnt[fd] = min(total, cap); p.pbase[fd] = a0 + shift; }
  cluster.sync();     // s_cta_total must outlive the remote reads
}

// Streaming form of the crop kernel for whole frames in topology order without by-residue expansion -- the kernel that
// reads the raw trajectory, 12 (float32) or 24 (float64) bytes per atom and frame, the one large HBM stream of the path.
//   * up to 8 CTAs share a frame; every WARP owns a contiguous run of atoms (a multiple of 128) and walks it 128 atoms at
//     a time: lane l loads the 12 coordinates of atoms 4l .. 4l+3 as three 16-byte words (six for float64) -- whole
//     128-byte lines per warp instruction -- with the next group's loads in flight while the current group is tested;
//   * kept atoms are written at once, transformed and compacted inside the warp's run, to a scratch frame (`tmp`, laid
//     out like the output); nothing is read twice from the trajectory; float32 coordinates are tested against float
//     thresholds that mean the same as numpy's double ones (NrlCropDev::lo32 / hi32): no FP64 instruction on the path;
//   * a second, small kernel moves every run to its final place with coalesced 16-byte copies (L2 hits).  Its launch is
//     the only synchronisation between the runs of a frame: with a cluster barrier instead (kept counts over distributed
//     shared memory) a third of the warps' resident time was spent waiting for the slowest run of the cluster.
//   Measured (B200, 200 frames x 200 k atoms): 1.7 TB/s for the first form -> 3.1 TB/s with the cluster barrier; a ring of
//   TMA bulk copies into shared memory instead of the register prefetch was slower (2.7 TB/s: 72 KB of ring per CTA cost
//   more occupancy than the deeper prefetch bought).
// Order preserving, cap applied at the move.  Same arithmetic as nrl_crop_pack_kernel, so the same atoms in the same
// order: tests/test_featurizer_gpu.py::test_streaming_crop_keeps_the_same_atoms_in_the_same_order runs both forms.
#define NRL_CROPS_NT 256

template <typename T> struct NrlLoad12;
template <> struct NrlLoad12<float> {
  static __device__ __forceinline__ void vec(const float *q, float (&c)[12]) {
    const float4 *v = reinterpret_cast<const float4 *>(q);
    const float4 a = __ldg(v), b = __ldg(v + 1), d = __ldg(v + 2);
    c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    c[8] = d.x; c[9] = d.y; c[10] = d.z; c[11] = d.w;
  }
};
template <> struct NrlLoad12<double> {
  static __device__ __forceinline__ void vec(const double *q, double (&c)[12]) {
    const double2 *v = reinterpret_cast<const double2 *>(q);
#pragma unroll
    for (int k = 0; k < 6; ++k) { const double2 a = __ldg(v + k); c[2 * k] = a.x; c[2 * k + 1] = a.y; }
  }
};

// in-box test of a transformed atom, in the coordinate dtype
template <typename T> struct NrlInBox;
template <> struct NrlInBox<float> {
  float lo[3], hi[3];
  __device__ __forceinline__ explicit NrlInBox(const NrlCropDev &cr) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = cr.lo32[k]; hi[k] = cr.hi32[k]; }
  }
  __device__ __forceinline__ bool test(float x, float y, float z) const {
    return (x < hi[0]) && (x > lo[0]) && (y < hi[1]) && (y > lo[1]) && (z < hi[2]) && (z > lo[2]);
  }
};
template <> struct NrlInBox<double> {
  double lo[3], hi[3];
  __device__ __forceinline__ explicit NrlInBox(const NrlCropDev &cr) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { lo[k] = cr.lo[k]; hi[k] = cr.hi[k]; }
  }
  __device__ __forceinline__ bool test(double x, double y, double z) const {
    return (x < hi[0]) && (x > lo[0]) && (y < hi[1]) && (y > lo[1]) && (z < hi[2]) && (z > lo[2]);
  }
};
// the transform with its constants in registers (same operations, same order as NrlXform)
template <typename T> struct NrlXformReg;
template <> struct NrlXformReg<float> {
  float f[3], c[3], h;
  __device__ __forceinline__ explicit NrlXformReg(const NrlCropDev &cr) : h(cr.half_f32) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { f[k] = cr.focal[k]; c[k] = cr.center[k]; }
  }
  __device__ __forceinline__ float apply(float x, int k) const { return ((x - f[k]) + c[k]) - h; }
};
template <> struct NrlXformReg<double> {
  double f[3], c[3], h;
  __device__ __forceinline__ explicit NrlXformReg(const NrlCropDev &cr) : h(cr.half_f64) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { f[k] = cr.focal64[k]; c[k] = (double)cr.center[k]; }
  }
  __device__ __forceinline__ double apply(double x, int k) const { return ((x - f[k]) + c[k]) - h; }
};

// pass 1: test, transform, compact inside the warp's run of the scratch frame; per-warp kept counts to `run_cnt`
template <typename T>
__global__ void __launch_bounds__(NRL_CROPS_NT, (sizeof(T) == 4 ? 6 : 3)) nrl_crop_stream_kernel(const NrlCropParams p, float4 *__restrict__ tmp,
                                                                       int *__restrict__ run_cnt, const int csize) {
  constexpr int nw = NRL_CROPS_NT / 32;
  const int fd = blockIdx.x / csize;
  const int rank = blockIdx.x - fd * csize;
  const NrlCropDev &cr = p.crops[p.frame_crop[fd]];
  const int N = p.n_atoms;
  const int a0 = cr.atom_begin, n_all = cr.atom_end - cr.atom_begin;
  const int shift = cr.table_shift;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // run of this warp: [s0, s1) of the slice, runs are multiples of 128 atoms and follow each other in (rank, warp) order
  const int seg = (((n_all + csize * nw - 1) / (csize * nw)) + 127) & ~127;
  const int s0 = min((rank * nw + warp) * seg, n_all), s1 = min(s0 + seg, n_all);
  const T *__restrict__ xyz = reinterpret_cast<const T *>(p.xyz) + ((long long)p.frame_src[fd] * N + a0) * 3;
  const unsigned char *__restrict__ sel = cr.selection >= 0 ? p.sel + (long long)cr.selection * N + a0 + shift : nullptr;
  const bool aligned = (reinterpret_cast<uintptr_t>(xyz) & 15) == 0;     // s0 * 3 * sizeof(T) is a multiple of 16
  float4 *__restrict__ mine = tmp + p.frame_off[fd] + s0;                 // never more kept atoms than atoms in the run
  const NrlXformReg<T> xf(cr);
  const NrlInBox<T> box(cr);
  const int payload0 = a0 + shift;

  int run = 0;
  auto sweep = [&](const int s, const T (&c)[12], const bool whole) {
    const int a = s + 4 * lane;
    float4 rec[4];
    unsigned nib = 0u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const T x = xf.apply(c[3 * u], 0), y = xf.apply(c[3 * u + 1], 1), z = xf.apply(c[3 * u + 2], 2);
      bool in = box.test(x, y, z) && (whole || a + u < s1);
      if (sel && in) in = sel[a + u] != 0;
      rec[u] = make_float4((float)x, (float)y, (float)z, __int_as_float(payload0 + a + u));   // commands.*: astype(float32)
      nib |= (in ? 1u : 0u) << u;
    }
    const int cnt = __popc(nib);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    int pos = run + incl - cnt;
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if ((nib >> u) & 1u) mine[pos++] = rec[u];
    run += __shfl_sync(0xffffffffu, incl, 31);
  };
  // groups of 128 atoms whose 16-byte loads are all inside the run; the rest (an unaligned frame, the last group of the
  // slice) is read coordinate by coordinate
  const int s_vec = aligned ? s0 + ((s1 - s0) & ~127) : s0;
  if (s0 < s_vec) {
    T cur[12], nxt[12];
    NrlLoad12<T>::vec(xyz + 3 * (long long)(s0 + 4 * lane), cur);
    for (int s = s0; s < s_vec; s += 128) {
      if (s + 128 < s_vec) NrlLoad12<T>::vec(xyz + 3 * (long long)(s + 128 + 4 * lane), nxt);
      sweep(s, cur, true);
#pragma unroll
      for (int k = 0; k < 12; ++k) cur[k] = nxt[k];
    }
  }
  for (int s = s_vec; s < s1; s += 128) {
    const int a = s + 4 * lane;
    T c[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) c[k] = (a + k / 3 < s1) ? xyz[3 * (long long)a + k] : (T)0;
    sweep(s, c, false);
  }
  if (lane == 0) run_cnt[(long long)fd * csize * nw + rank * nw + warp] = run;
}

// pass 2: every warp moves its compact run to its place in the packed frame (coalesced 16-byte copies out of L2); the
// kernel boundary is the only synchronisation between the runs of a frame -- no warp waits, resident, for another one
__global__ void __launch_bounds__(NRL_CROPS_NT) nrl_crop_place_kernel(const NrlCropParams p, const float4 *__restrict__ tmp,
                                                                      const int *__restrict__ run_cnt, const int csize) {
  constexpr int nw = NRL_CROPS_NT / 32;
  const int fd = blockIdx.x / csize;
  const int rank = blockIdx.x - fd * csize;
  const NrlCropDev &cr = p.crops[p.frame_crop[fd]];
  const int n_all = cr.atom_end - cr.atom_begin;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg = (((n_all + csize * nw - 1) / (csize * nw)) + 127) & ~127;
  const int me = rank * nw + warp, n_runs = csize * nw;              // <= 64 runs per frame
  const int s0 = min(me * seg, n_all);
  const int *cnt = run_cnt + (long long)fd * n_runs;
  const int c0 = lane < n_runs ? cnt[lane] : 0, c1 = lane + 32 < n_runs ? cnt[lane + 32] : 0;
  int before = (lane < me ? c0 : 0) + (lane + 32 < me ? c1 : 0), total = c0 + c1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, o); total += __shfl_xor_sync(0xffffffffu, total, o); }
  const int run = cnt[me];
  const int cap = cr.max_atoms > 0 ? cr.max_atoms : 0x7fffffff;
  const long long foff = p.frame_off[fd];
  const float4 *__restrict__ mine = tmp + foff + s0;
  float4 *__restrict__ out = p.atoms + foff;
  for (int k = lane; k < run; k += 32)
    if (before + k < cap) out[before + k] = mine[k];
  if (me == 0 && lane == 0) { p.off[fd] = foff; p.cnt[fd] = min(total, cap); p.pbase[fd] = cr.atom_begin + cr.table_shift; }
}

// ---- per-atom normalisation of the voxeliser -------------------------------------------
// S_a = sum of the (unnormalised) Gaussian over the lattice points of the BUFFERED grid
// [-b, n+b-1]^3 strictly inside the cutoff (voxelize.cu:79-96); the 1/(sigma*sqrt(2pi))
// prefactor of gaussian_map_device cancels in G/S and is dropped.  One warp per packed atom,
// lanes over the flattened bounding cube, partial sums in double.  Writes 1/S (0 when the
// reference skips the atom: S == 0, voxelize.cu:109).
struct NrlNormParams {
  NrlFrames frames;
  const int *frame_list;     // frames to process
  int nx, ny, nz;
  float spacing, cutoff, k2;
  float *inv_norm;           // per frame at frames.off, addressed by payload - frames.pbase
};

#define NRL_NORM_MAXZ 96     // lattice points along an axis inside one cutoff sphere handled by the fast path

// in-cutoff predicate of the reference, bit for bit (voxelize.cu:92-95 as compiled)
__device__ __forceinline__ bool nrl_norm_in(const float az, const int z, const float s, const float t2, const float c2) {
  const float dz = az - (float)z * s;
  return fmaf(dz, dz, t2) < c2;
}

__global__ void __launch_bounds__(128) nrl_atom_norm_kernel(const NrlNormParams p, const int capacity) {
  // per warp: prefix of the z-plane factors, and per-axis tables of the x / y offsets and factors
  __shared__ float s_prefix[4][NRL_NORM_MAXZ + 1];
  __shared__ float s_dx[4][NRL_NORM_MAXZ], s_ex[4][NRL_NORM_MAXZ], s_dy2[4][NRL_NORM_MAXZ], s_ey[4][NRL_NORM_MAXZ];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warps_per_cta = blockDim.x >> 5;
  const int fd = p.frame_list[blockIdx.y];
  const int n = min(p.frames.cnt[fd], capacity);
  const long long off = p.frames.off[fd];
  const int pb = p.frames.pbase[fd];
  const float s = p.spacing, c = p.cutoff, c2 = c * c, k2 = p.k2;
  const int b = nrl_buff_dim_device(c, s);
  float *P = s_prefix[wid], *Dx = s_dx[wid], *Ex = s_ex[wid], *Dy2 = s_dy2[wid], *Ey = s_ey[wid];
  for (int i = blockIdx.x * warps_per_cta + wid; i < n; i += gridDim.x * warps_per_cta) {
    const float4 a = __ldg(p.frames.atoms + off + i);
    double sum = 0.0;
    if (isfinite(a.x) && isfinite(a.y) && isfinite(a.z)) {
      // lattice planes that can hold a point inside the cutoff, with a margin far above the float
      // rounding of the exact predicates that decide membership below
      const float is = 1.0f / s, cr = c * is;
      const float mx = 2e-3f + 2e-6f * fabsf(a.x * is), my = 2e-3f + 2e-6f * fabsf(a.y * is), mz = 2e-3f + 2e-6f * fabsf(a.z * is);
      const int x_lo = (int)fmaxf(ceilf(a.x * is - cr - mx), (float)(-b));
      const int x_hi = (int)fminf(floorf(a.x * is + cr + mx), (float)(p.nx + b - 1));
      const int y_lo = (int)fmaxf(ceilf(a.y * is - cr - my), (float)(-b));
      const int y_hi = (int)fminf(floorf(a.y * is + cr + my), (float)(p.ny + b - 1));
      const int z_lo = (int)fmaxf(ceilf(a.z * is - cr - mz), (float)(-b));
      const int z_hi = (int)fminf(floorf(a.z * is + cr + mz), (float)(p.nz + b - 1));
      const int ex = x_hi - x_lo + 1, ey = y_hi - y_lo + 1, ez = z_hi - z_lo + 1;
      if (ex > 0 && ey > 0 && ez > 0 && ez <= NRL_NORM_MAXZ && ex <= NRL_NORM_MAXZ && ey <= NRL_NORM_MAXZ) {
        // Separable form.  exp(-(dx^2+dy^2+dz^2)/2s^2) = E_x * E_y * E_z, and along z the lattice points
        // that pass the reference's float predicate form ONE interval (the predicate is monotone in
        // |dz|), so a row contributes E_x * E_y * (prefix[z_last+1] - prefix[z_first]).  The per-axis
        // offsets are tabulated exactly as the predicate needs them (dx, and dy*dy rounded once).
        __syncwarp();
        float carry = 0.0f;
        for (int j0 = 0; j0 < ez; j0 += 32) {
          const int j = j0 + lane;
          float v = 0.0f;
          if (j < ez) { const float dz = a.z - (float)(z_lo + j) * s; v = nrl_ex2_approx(dz * dz * k2); }
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const float u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
          if (j < ez) P[j + 1] = carry + v;
          carry += __shfl_sync(0xffffffffu, v, 31);
        }
        if (lane == 0) P[0] = 0.0f;
        for (int j = lane; j < ex; j += 32) { const float dx = a.x - (float)(x_lo + j) * s; Dx[j] = dx; Ex[j] = nrl_ex2_approx(dx * dx * k2); }
        for (int j = lane; j < ey; j += 32) { const float dy = a.y - (float)(y_lo + j) * s; const float d2 = dy * dy; Dy2[j] = d2; Ey[j] = nrl_ex2_approx(d2 * k2); }
        __syncwarp();
        const int rows = ex * ey;
        const float inv_s = 1.0f / s, zc = a.z * inv_s;
        // reach of float rounding, in lattice units, on the position of an interval end: coordinate
        // scaling (2 ulp of zc), the cancellation in c2 - t2 and the predicate's own last-bit fuzz
        const float eps = 5e-4f + 4e-6f * fabsf(zc), one_m_eps = 1.0f - eps;
        // rows q = lane, lane + 32, ... of the (x, y) rectangle, walked without a division per row
        const int step_x = 32 / ey, step_y = 32 - step_x * ey;
        int xi = lane / ey, yi = lane - xi * ey;
        float part = 0.0f;
        for (int q = lane; q < rows; q += 32) {
          const float dx = Dx[xi], dy2 = Dy2[yi];
          const float exy = Ex[xi] * Ey[yi];
          yi += step_y; xi += step_x;
          if (yi >= ey) { yi -= ey; ++xi; }
          const float t2 = fmaf(dx, dx, dy2);                     // voxelize.cu:92-94 as compiled, z term pending
          if (!(t2 < c2)) continue;
          const float gap = c2 - t2;
          const float r = nrl_sqrt_approx(gap) * inv_s;        // half-width of the interval in lattice units
          const float ul = zc - r, uh = zc + r;
          const float fl = ceilf(ul), fh = floorf(uh);
          int jl = (int)fl, jh = (int)fh;
          // The float predicate can only disagree with this real-number estimate when an end falls
          // within rounding reach of a lattice plane, or when the row is nearly tangent to the sphere
          // (gap of a few ulp).  Only then are the two ends settled on the exact predicate.
          const float el = fl - ul, eh = uh - fh;                // distances of the ends to the planes beside them
          const bool fragile = (el < eps) || (el > one_m_eps) || (eh < eps) || (eh > one_m_eps) || (r < 0.05f);
          if (fragile) {
            jl = max(jl, z_lo - 1); jh = min(jh, z_hi + 1);
            int guard = 0;
            while (jl > z_lo && nrl_norm_in(a.z, jl - 1, s, t2, c2) && guard < 8) { --jl; ++guard; }
            while (jl <= jh && !nrl_norm_in(a.z, jl, s, t2, c2) && guard < 16) { ++jl; ++guard; }
            while (jh < z_hi && nrl_norm_in(a.z, jh + 1, s, t2, c2) && guard < 24) { ++jh; ++guard; }
            while (jh >= jl && !nrl_norm_in(a.z, jh, s, t2, c2) && guard < 32) { --jh; ++guard; }
          }
          jl = max(jl, z_lo); jh = min(jh, z_hi);
          if (jl <= jh) part = fmaf(exy, P[jh - z_lo + 1] - P[jl - z_lo], part);
        }
        sum = (double)part;
      } else if (ex > 0 && ey > 0 && ez > 0) {
        // very fine lattices: plain sweep of the bounding cube
        const int plane = ey * ez;
        const float rcp_ez = 1.0f / (float)ez;
        for (int xi = 0; xi < ex; ++xi) {
          const float dx = a.x - (float)(x_lo + xi) * s;
          if (!(dx * dx < c2)) continue;
          for (int q = lane; q < plane; q += 32) {
            const int yi = __float2int_rd(((float)q + 0.5f) * rcp_ez);
            const int zi = q - yi * ez;
            const float dy = a.y - (float)(y_lo + yi) * s, dz = a.z - (float)(z_lo + zi) * s;
            const float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));        // voxelize.cu:92-94 as compiled
            if (d2 < c2) sum += (double)nrl_ex2_approx(d2 * k2);
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) p.inv_norm[off + (__float_as_int(a.w) - pb)] = sum > 0.0 ? (float)(1.0 / sum) : 0.0f;
  }
}
