// nearl_b200 -- shared device-side definitions.
//
// Arithmetic contract.  This translation unit is compiled with -fmad=false -ftz=true:
// the compiler never contracts a*b+c on its own, every fused multiply-add below is an
// explicit fmaf() that restates the FMA the reference binary executes (read from the
// PTX of /root/reference/src built with its own flags, -use_fast_math -O3), and
// denormals flush to zero as they do there.  The approximate division / square root /
// log2 the reference gets from -use_fast_math are issued explicitly (inline PTX) only
// where the reference's result feeds a discrete decision (buff_dim, entropy bins) or
// where following it is free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define NRL_PLACEHOLDER 99999.0f      // reference src/constants.h:9-10
#define NRL_MAX_FRAMES 512            // reference src/constants.h:11
#define NRL_DISTINCT_LIMIT 1000.0f    // reference src/constants.h:12
#define NRL_ENTROPY_BINS 16           // reference src/constants.h:34

// tile geometry of the gather kernels: 8x8x8 voxels, 4 consecutive z per thread
#define NRL_TX 8
#define NRL_TY 8
#define NRL_TZ 8
#define NRL_TILE_VOX (NRL_TX * NRL_TY * NRL_TZ)
#define NRL_VPT 4
#define NRL_MAX_CH 4
#define NRL_NT (NRL_TILE_VOX / NRL_VPT)   // 128 threads
#define NRL_NW (NRL_NT / 32)

enum NrlMode {
  NRL_MODE_VOXELIZE = 0,         // Gaussian deposit (frame_voxelize / density_flow)
  NRL_OBS_EXISTENCE = 1,
  NRL_OBS_DIRECT_COUNT = 2,
  NRL_OBS_DISTINCT_COUNT = 3,
  NRL_OBS_MEAN_DISTANCE = 11,
  NRL_OBS_CUMULATIVE_WEIGHT = 12,
  NRL_OBS_DENSITY = 13,
  NRL_OBS_DISPERSION = 14,
  NRL_OBS_ECCENTRICITY = 15,
  NRL_OBS_RADIUS_OF_GYRATION = 16,
  NRL_OBS_FUSED = 100,           // several observables of one focus box from ONE sweep (cell-list kernel only)
};

// what a fused sweep has to accumulate besides the in-range count (NrlTask::fuse_flags, the kernel's second template argument)
enum { NRL_FUSE_SEEN = 1,        // distinct_count: OR of the value-rank bits
       NRL_FUSE_WS = 2,          // cumulative_weight, density, mean_distance, eccentricity: sum of weights
       NRL_FUSE_R = 4,           // mean_distance: sum of distance * weight
       NRL_FUSE_COM = 8 };       // eccentricity: weighted coordinate sums

__device__ __forceinline__ float nrl_div_approx(float a, float b) {
  float r;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float nrl_rcp_approx(float a) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
// eccentricity_device (marching_observers.cu:289-298) as the reference binary evaluates it: ONE reciprocal of the weight
// sum, each centre-of-mass division fused with the subtraction from the observer (d = fma(rcp, -com, o), no rounded
// quotient in between), squares summed x, y, z, approximate square root
__device__ __forceinline__ float nrl_eccentricity(float ox, float oy, float oz, float cx, float cy, float cz, float ws);
__device__ __forceinline__ float nrl_sqrt_approx(float a) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float nrl_eccentricity(float ox, float oy, float oz, float cx, float cy, float cz, float ws) {
  const float rcp = nrl_rcp_approx(ws);
  const float dx = fmaf(rcp, -cx, ox), dy = fmaf(rcp, -cy, oy), dz = fmaf(rcp, -cz, oz);
  return nrl_sqrt_approx(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
}
__device__ __forceinline__ float nrl_lg2_approx(float a) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float nrl_ex2_approx(float a) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

// reference src/voxelize.cu:50  `const int buff_dim = (cutoff + spacing) / spacing;`
// compiled to add.ftz / div.approx.ftz / cvt.rzi -- reproduced instruction for instruction
__device__ __forceinline__ int nrl_buff_dim_device(float cutoff, float spacing) {
  float q = nrl_div_approx(cutoff + spacing, spacing);
  int r;
  asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(r) : "f"(q));
  return r;
}

// One unit of work for the tile kernels: one (feature, frame window, focus box).
// The lattice has extents (nx, ny, nz) with z fastest: flat = (x*ny + y)*nz + z.  For the
// reference's index math (src/voxelize.cu:113-115, src/marching_observers.cu:416-420:
// x = g/d0/d1, y = g/d0%d1, z = g%d0) that is (nx,ny,nz) = (dims[2], dims[1], dims[0]),
// the plain C layout for the cubic grids every reference config uses.
struct NrlTask {
  int nx, ny, nz;
  int ntx, nty, ntz;        // tiles per axis (ntx: of the task's slab, see tx0)
  int tx0;                  // first tile along x: a task may cover only the slab [tx0, tx0 + ntx) of its grid (the grid
                            // of one task split over several GPUs); ch_out is shifted so that the slab is stored compactly
  int tile_begin;           // first work item of this task in the launch
  int n_frames;             // frames observed (already clamped to NRL_MAX_FRAMES)
  // A work item produces up to NRL_MAX_CH output grids ("channels") from ONE sweep: several reducers
  // of the same per-frame value, and -- voxeliser only -- up to NRL_MAX_CH weight tables deposited
  // together (the geometry, the cutoff tests and the exponentials are shared).
  int n_ch;                 // channels
  int n_w;                  // weight slots (1 unless voxeliser with several weight tables)
  int ch_agg[NRL_MAX_CH];   // per channel: 0 = single frame written as is, 1..8 reducer id
  int ch_w[NRL_MAX_CH];     // per channel: weight slot whose per-frame values it reduces
  int ch_mode[NRL_MAX_CH];  // NRL_OBS_FUSED: the observable of every channel
  int fuse_flags;           // NRL_OBS_FUSED: NRL_FUSE_* needed by the channels (wtab[0]: weights, wtab[1]: value ranks)
  long long ch_out[NRL_MAX_CH];   // per channel: first output float
  int mode;                 // NrlMode
  float spacing;
  float cutoff;
  float k2;                 // voxelize: -0.5*log2(e)/sigma^2
  float sigma;
  double sphere_volume;     // density: (4/3) pi cutoff^3 evaluated in double, used as a float (marching_observers.cu:200-201)
  long long frame0;         // first frame descriptor
  const float *wtab[NRL_MAX_CH];   // per weight slot: table addressed by the atom's index payload
  const float *inv_norm;    // voxelize: 1/S per packed atom (parallel to atoms), else null
};

struct NrlFrames {
  const float4 *atoms;      // packed (x, y, z, index-payload) per atom
  const long long *off;     // per frame: first packed atom
  const int *cnt;           // per frame: number of packed atoms (device-produced)
  const int *pbase;         // per frame: payload of the frame's first POSSIBLE atom; per-atom side arrays of a frame
                            // (1/S of the voxeliser) are addressed by payload - pbase, whatever order the atoms are stored in
};

struct NrlTileParams {
  const NrlTask *tasks;
  int n_tasks;
  int n_work;               // total tiles over all tasks
  NrlFrames frames;
  float *out;
  float4 *cand_scratch;     // per-CTA overflow candidate list
  long long cand_scratch_stride;   // float4 per CTA
  float *ser_scratch;       // per-CTA overflow series
  long long ser_scratch_stride;    // floats per CTA
  int cand_cap;             // candidates that fit in shared memory
  int ser_cap;              // series floats that fit in shared memory
};
