// nearl_b200 -- cell-list form of the gather kernels (K1 voxeliser, K2 observers, fused K3).
//
//   nrl_bin_kernel        counting sort of every packed frame into a uniform cell list (cells of
//                         about one tile edge, z fastest), order preserving inside a cell, so the
//                         result is deterministic.  Atoms that no voxel can reach are dropped,
//                         atoms with a NaN coordinate go to a bucket every tile visits (the
//                         observers count them as in range, marching_observers.cu:40).
//   nrl_cell_tile_kernel  one CTA per 8x8x8 tile of one task, walking the frames of the window.
//                         The atoms a tile can see are the cells of <= 32 (cx, cy) columns; each
//                         column is ONE contiguous run of the binned frame and is staged into
//                         shared memory by a 1-D bulk asynchronous copy (TMA engine,
//                         cp.async.bulk -> UBLKCP) completing on an mbarrier.  The runs of frame
//                         f+1 are in flight while frame f is swept.  Every warp compacts, from the
//                         staged cells, the private list of atoms that can reach its own 8x4x4
//                         block (exact sphere/box test), resolves their weights once, and sweeps
//                         that list with the reference's squared-distance arithmetic.  Per-frame
//                         values go to a shared-memory time series that is reduced in place by
//                         the reference's reducer.  No global atomics, no [F,G] intermediate, no
//                         per-frame launch, no full-frame scan per tile.
//
// Replaces: frame_interp_global      /root/reference/src/voxelize.cu:38-130 (deposit half)
//           marching_observer_global /root/reference/src/marching_observers.cu:405-429
//           observables 1,2,3,11,12,13,15   /root/reference/src/marching_observers.cu:30-299
//           gridwise_aggregation_global /root/reference/src/gpuutils.cu:48-77 (fused)
// (dispersion, radius of gyration and distinct_count over more than 32 distinct values keep the
//  scan kernels of tile_kernels.cuh: they need the whole candidate list at once / in value order)
#pragma once
#include "agg_kernels.cuh"
#include "common.cuh"
#include "reducers.cuh"
#include "tile_kernels.cuh"

struct NrlCellGrid {          // one per frame descriptor; the frames of a crop share the geometry
  float ox, oy, oz;           // lower corner of cell (0,0,0)
  float inv_h;                // 1 / cell edge along x and y
  int ncx, ncy, ncz;          // cells per axis, cell = (cx*ncy + cy)*ncz + cz
  float inv_hz;               // 1 / cell edge along z: a tile reads whole z-runs of cells, finer cells trim the run ends
  long long cs_off;           // this frame's row of cell_start: ncell + 2 ints (prefix, NaN bucket last)
  float bx, by, bz;           // far corner of the box spanned by the voxel centres (near corner: the origin)
  float reach2;               // squared reach: an atom farther than this from that box touches no voxel and is dropped
};

// Cell of an atom.  Monotone in every coordinate (float rounding is monotone), and the tile kernel
// pushes the corners of its search box through the same expression, so a tile never misses a cell.
__device__ __forceinline__ float nrl_cell_coord(const float x, const float o, const float inv_h) { return (x - o) * inv_h; }
__device__ __forceinline__ int nrl_cell_of(const float4 a, const NrlCellGrid &g, const int ncell) {
  if (a.x != a.x || a.y != a.y || a.z != a.z) return ncell;                 // NaN bucket
  const float fx = nrl_cell_coord(a.x, g.ox, g.inv_h), fy = nrl_cell_coord(a.y, g.oy, g.inv_h), fz = nrl_cell_coord(a.z, g.oz, g.inv_hz);
  if (!(fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < (float)g.ncx && fy < (float)g.ncy && fz < (float)g.ncz)) return -1;
  // corners of the padded box: inside the cell grid, yet out of every voxel's reach (reach2 carries the margin)
  const float qx = fmaxf(fmaxf(-a.x, a.x - g.bx), 0.0f), qy = fmaxf(fmaxf(-a.y, a.y - g.by), 0.0f), qz = fmaxf(fmaxf(-a.z, a.z - g.bz), 0.0f);
  if (fmaf(qz, qz, fmaf(qy, qy, qx * qx)) > g.reach2) return -1;
  return ((int)fx * g.ncy + (int)fy) * g.ncz + (int)fz;
}

struct NrlBinParams {
  NrlFrames in;               // packed frames in emission order
  const NrlCellGrid *grids;   // per frame descriptor
  const int *frame_list;      // frames to bin
  float4 *atoms_out;          // binned atoms; frame fd starts at in.off[fd]
  int *cnt_out;               // per frame: atoms binned (dropped atoms excluded)
  int *cell_start;
  int capacity;
};

#define NRL_BIN_NT 1024
#define NRL_BIN_BIG_CELL 48

__global__ void __launch_bounds__(NRL_BIN_NT) nrl_bin_kernel(const NrlBinParams p) {
  extern __shared__ int s_cell[];                 // ncell + 2 counters / cursors
  __shared__ int s_warp[NRL_BIN_NT / 32];
  __shared__ int s_carry, s_big;
  const int fd = p.frame_list[blockIdx.x];
  const NrlCellGrid g = p.grids[fd];
  const int ncell = g.ncx * g.ncy * g.ncz;
  const int n = min(p.in.cnt[fd], p.capacity);
  const long long off = p.in.off[fd];
  const float4 *__restrict__ src = p.in.atoms + off;
  float4 *dst = p.atoms_out + off;
  int *cs = p.cell_start + g.cs_off;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < ncell + 2; i += NRL_BIN_NT) s_cell[i] = 0;
  if (tid == 0) { s_carry = 0; s_big = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += NRL_BIN_NT) {
    const int c = nrl_cell_of(src[i], g, ncell);
    if (c >= 0) atomicAdd(&s_cell[c], 1);
  }
  __syncthreads();
  // exclusive prefix over the ncell + 1 buckets, 1024 at a time
  for (int base = 0; base < ncell + 1; base += NRL_BIN_NT) {
    const int i = base + tid;
    const int v = i < ncell + 1 ? s_cell[i] : 0;
    if (v > NRL_BIN_BIG_CELL) s_big = 1;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int wbase = 0, chunk = 0;
#pragma unroll
    for (int w = 0; w < NRL_BIN_NT / 32; ++w) { const int t = s_warp[w]; if (w < warp) wbase += t; chunk += t; }
    const int excl = s_carry + wbase + incl - v;
    if (i < ncell + 1) { s_cell[i] = excl; cs[i] = excl; }
    __syncthreads();
    if (tid == 0) s_carry += chunk;
    __syncthreads();
  }
  const int total = s_carry;
  if (tid == 0) { cs[ncell + 1] = total; p.cnt_out[fd] = total; }
  // scatter (the order inside a cell is settled below)
  for (int i = tid; i < n; i += NRL_BIN_NT) {
    const float4 a = src[i];
    const int c = nrl_cell_of(a, g, ncell);
    if (c >= 0) dst[atomicAdd(&s_cell[c], 1)] = a;
  }
  __syncthreads();
  // s_cell[c] is now the END of cell c.  Emission order inside every cell: the payload grows with the
  // emission index, so a small cell is insertion-sorted by it ...
  for (int c = tid; c < ncell + 1; c += NRL_BIN_NT) {
    const int b = c ? s_cell[c - 1] : 0, e = s_cell[c];
    if (e - b > NRL_BIN_BIG_CELL || e - b < 2) continue;
    for (int k = b + 1; k < e; ++k) {
      const float4 key = dst[k];
      const int kp = __float_as_int(key.w);
      int j = k - 1;
      while (j >= b && __float_as_int(dst[j].w) > kp) { dst[j + 1] = dst[j]; --j; }
      dst[j + 1] = key;
    }
  }
  // ... and the atoms of an (unphysically) crowded cell are placed by counting their predecessors
  if (s_big) {
    for (int i = tid; i < n; i += NRL_BIN_NT) {
      const float4 a = src[i];
      const int c = nrl_cell_of(a, g, ncell);
      if (c < 0) continue;
      const int b = c ? s_cell[c - 1] : 0, e = s_cell[c];
      if (e - b <= NRL_BIN_BIG_CELL) continue;
      int rank = 0;
      for (int j = 0; j < i; ++j) rank += (nrl_cell_of(src[j], g, ncell) == c) ? 1 : 0;
      dst[b + rank] = a;
    }
  }
}

// ---- the cell-list tile kernel -------------------------------------------------------------------
#define NRL_MAX_RUNS 32
#define NRL_STAGE_CAP 640          // atoms staged per frame (10 KB); fuller frames are read straight from global memory

struct NrlCellTileParams {
  const NrlTask *tasks;
  int n_tasks;
  int n_work;
  NrlFrames frames;              // BINNED frames
  const NrlCellGrid *grids;      // per frame descriptor
  const int *cell_start;
  float *out;
  float *ser_scratch;            // per-CTA overflow series
  long long ser_scratch_stride;
  int ser_cap;                   // series floats that fit in shared memory
  int wcap;                      // records per warp
};

// every channel reduced by mean / max / min: the reducers run as recurrences over the frames, no series is kept
__host__ __device__ __forceinline__ bool nrl_streaming_task(const NrlTask &t) {
  if (t.n_ch < 1 || t.ch_agg[0] == 0) return false;
  if (t.mode != NRL_OBS_FUSED && t.n_ch != 1) return false;      // single-observable kernels keep one recurrence (registers)
  for (int ch = 0; ch < t.n_ch; ++ch)
    if (!(t.ch_agg[ch] == 1 || t.ch_agg[ch] == 5 || t.ch_agg[ch] == 6)) return false;
  return true;
}
// rows of the per-tile time series: one per weight slot, or one per channel of a fused sweep
__host__ __device__ __forceinline__ int nrl_series_rows(const NrlTask &t) { return t.mode == NRL_OBS_FUSED ? t.n_ch : t.n_w; }

template <int MODE, int NWT> struct NrlRecStride {   // float4 per record
  static constexpr int value = (MODE == NRL_MODE_VOXELIZE) ? (NWT > 1 ? 3 : 2) : (MODE == NRL_OBS_FUSED && (NWT & NRL_FUSE_SEEN)) ? 2 : 1;
};
// per-frame values a sweep produces: one per weight table (voxeliser), one per channel (fused observers), else one
template <int MODE, int NWT> struct NrlValCount {
  static constexpr int value = (MODE == NRL_OBS_FUSED) ? NRL_MAX_CH : NWT;
};

// accumulators of one thread's four voxels; a frame may be swept in several rounds
template <int MODE, int NWT> struct NrlAcc {
  static constexpr int NF0 = (MODE == NRL_OBS_FUSED) ? 1 : NWT;
  float f0[NF0][NRL_VPT];   // vox: deposits / exist, count: flag, count / mean_distance: sum r*w / weights: sum w / ecc: sum w*x
  float f1[NRL_VPT], f2[NRL_VPT], f3[NRL_VPT], f4[NRL_VPT];
  int n[NRL_VPT];
  unsigned seen[NRL_VPT];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) {
#pragma unroll
      for (int w = 0; w < NF0; ++w) f0[w][k] = 0.0f;
      f1[k] = f2[k] = f3[k] = f4[k] = 0.0f; n[k] = 0; seen[k] = 0u;
    }
  }
};

// one round: the warp's `nrec` records (a multiple of four, padded with far-away records)
template <int MODE, int NWT>
__device__ __forceinline__ bool nrl_rec_sweep(const float4 *rec, const int nrec, const NrlTask &t, const float c2, const float px,
                                              const float py, const float (&pz)[NRL_VPT], NrlAcc<MODE, NWT> &A) {
  constexpr int S = NrlRecStride<MODE, NWT>::value;
  for (int r0 = 0; r0 < nrec; r0 += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 a = rec[(r0 + u) * S];
      if constexpr (MODE == NRL_MODE_VOXELIZE) {
        const float4 ezq = rec[(r0 + u) * S + 1];
        const float part = nrl_vox_partial(a.x, a.y, px, py);
        const float exy = nrl_ex2_approx(part * t.k2);
        const float ez[NRL_VPT] = {ezq.x, ezq.y, ezq.z, ezq.w};
        if constexpr (NWT > 1) {
          // same association as the single-table form below, so a feature gets the same bits whether or not
          // it shares its sweep with other weight tables
          const float4 ww = rec[(r0 + u) * S + 2];
          const float wxy[4] = {exy * ww.x, exy * ww.y, exy * ww.z, exy * ww.w};
#pragma unroll
          for (int k = 0; k < NRL_VPT; ++k) {
            const float dz = a.z - pz[k];
            const float d2 = fmaf(dz, dz, part);
            if (d2 < c2) {                                               // voxelize.cu:121-125
#pragma unroll
              for (int w = 0; w < NWT; ++w) A.f0[w][k] = fmaf(wxy[w], ez[k], A.f0[w][k]);
            }
          }
        } else {
          const float wxy = exy * a.w;
#pragma unroll
          for (int k = 0; k < NRL_VPT; ++k) {
            const float dz = a.z - pz[k];
            const float d2 = fmaf(dz, dz, part);
            if (d2 < c2) A.f0[0][k] = fmaf(wxy, ez[k], A.f0[0][k]);
          }
        }
      } else if constexpr (MODE == NRL_OBS_FUSED) {
        // several observables of the same atoms: one distance test, every accumulator the channels need
        // (f0: sum r*w, f1: sum w, f2..f4: sum w*(x, y, z); each with the arithmetic of its single-mode form)
        unsigned mask = 0u;
        if constexpr ((NWT & NRL_FUSE_SEEN) != 0) mask = (unsigned)__float_as_int(rec[(r0 + u) * S + 1].x);
        const float part = nrl_obs_partial(px, py, a.x, a.y);
#pragma unroll
        for (int k = 0; k < NRL_VPT; ++k) {
          const float dz = pz[k] - a.z;
          const float d2 = fmaf(dz, dz, part);
          if (!(d2 > c2)) {
            // with a rank mask in the group, its presence bit (bit 31, set for every atom) is the "some atom in range"
            // flag and the counter is not kept (a direct count never shares a group with a rank mask)
            if constexpr ((NWT & NRL_FUSE_SEEN) == 0) A.n[k] += 1;
            if constexpr ((NWT & NRL_FUSE_SEEN) != 0) A.seen[k] |= mask;
            if constexpr ((NWT & NRL_FUSE_R) != 0) A.f0[0][k] = fmaf(nrl_sqrt_approx(d2), a.w, A.f0[0][k]);
            if constexpr ((NWT & NRL_FUSE_WS) != 0) A.f1[k] = A.f1[k] + a.w;
            if constexpr ((NWT & NRL_FUSE_COM) != 0) {
              A.f2[k] = fmaf(a.w, a.x, A.f2[k]);
              A.f3[k] = fmaf(a.w, a.y, A.f3[k]);
              A.f4[k] = fmaf(a.w, a.z, A.f4[k]);
            }
          }
        }
      } else {
        const float part = nrl_obs_partial(px, py, a.x, a.y);
#pragma unroll
        for (int k = 0; k < NRL_VPT; ++k) {
          const float dz = pz[k] - a.z;
          const float d2 = fmaf(dz, dz, part);
          if (!(d2 > c2)) {                                              // marching_observers.cu:40,77,127,162,185
            if constexpr (MODE == NRL_OBS_EXISTENCE) A.f0[0][k] = 1.0f;
            else if constexpr (MODE == NRL_OBS_DIRECT_COUNT) A.f0[0][k] = A.f0[0][k] + 1.0f;
            else if constexpr (MODE == NRL_OBS_DISTINCT_COUNT) A.seen[k] |= (unsigned)__float_as_int(a.w);
            else if constexpr (MODE == NRL_OBS_MEAN_DISTANCE) {          // :160-167
              A.f0[0][k] = fmaf(nrl_sqrt_approx(d2), a.w, A.f0[0][k]);
              A.f1[k] = A.f1[k] + a.w;
              A.n[k] += 1;
            } else if constexpr (MODE == NRL_OBS_CUMULATIVE_WEIGHT || MODE == NRL_OBS_DENSITY) {
              A.f0[0][k] = A.f0[0][k] + a.w;                             // :183-189
            } else if constexpr (MODE == NRL_OBS_ECCENTRICITY) {         // :279-287
              A.n[k] += 1;
              A.f0[0][k] = fmaf(a.w, a.x, A.f0[0][k]);
              A.f1[k] = fmaf(a.w, a.y, A.f1[k]);
              A.f2[k] = fmaf(a.w, a.z, A.f2[k]);
              A.f3[k] = A.f3[k] + a.w;
            }
          }
        }
      }
    }
    if constexpr (MODE == NRL_OBS_EXISTENCE) {
      // existence saturates: once every voxel of the warp has seen an atom nothing can change
      if (__all_sync(0xffffffffu, (A.f0[0][0] + A.f0[0][1]) + (A.f0[0][2] + A.f0[0][3]) == 4.0f)) return true;
    }
  }
  return false;
}

template <int MODE, int NWT>
__device__ __forceinline__ void nrl_rec_finish(const NrlAcc<MODE, NWT> &A, const NrlTask &t, const float px, const float py,
                                               const float (&pz)[NRL_VPT], float (&val)[NrlValCount<MODE, NWT>::value][NRL_VPT]) {
  if constexpr (MODE == NRL_OBS_FUSED) {
#pragma unroll
    for (int ch = 0; ch < NRL_MAX_CH; ++ch) {
      const int m = ch < t.n_ch ? t.ch_mode[ch] : 0;
#pragma unroll
      for (int k = 0; k < NRL_VPT; ++k) {
        float v = 0.0f;
        bool any;
        if constexpr ((NWT & NRL_FUSE_SEEN) != 0) any = A.seen[k] != 0u; else any = A.n[k] > 0;
        if (m == NRL_OBS_EXISTENCE) v = any ? 1.0f : 0.0f;
        else if (m == NRL_OBS_DIRECT_COUNT) v = (float)A.n[k];
        else if (m == NRL_OBS_DISTINCT_COUNT) v = fminf((float)__popc(A.seen[k] & 0x7fffffffu), NRL_DISTINCT_LIMIT);
        else if (m == NRL_OBS_MEAN_DISTANCE) v = any ? nrl_div_approx(A.f0[0][k], A.f1[k]) : 0.0f;
        else if (m == NRL_OBS_CUMULATIVE_WEIGHT) v = A.f1[k];
        else if (m == NRL_OBS_DENSITY) v = nrl_div_approx(A.f1[k], (float)t.sphere_volume);      // float volume, fast-math division (:199-201)
        else if (m == NRL_OBS_ECCENTRICITY) {
          if (any && A.f1[k] != 0.0f) {
            v = nrl_eccentricity(px, py, pz[k], A.f2[k], A.f3[k], A.f4[k], A.f1[k]);
          } else {
            v = t.cutoff;
          }
        }
        val[ch][k] = v;
      }
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < NRL_VPT; ++k) {
    if constexpr (MODE == NRL_MODE_VOXELIZE) {
#pragma unroll
      for (int w = 0; w < NrlAcc<MODE, NWT>::NF0; ++w) val[w][k] = A.f0[w][k];
    } else if constexpr (MODE == NRL_OBS_DISTINCT_COUNT) {
      val[0][k] = fminf((float)__popc(A.seen[k]), NRL_DISTINCT_LIMIT);
    } else if constexpr (MODE == NRL_OBS_MEAN_DISTANCE) {
      val[0][k] = A.n[k] > 0 ? nrl_div_approx(A.f0[0][k], A.f1[k]) : 0.0f;
    } else if constexpr (MODE == NRL_OBS_DENSITY) {
      val[0][k] = nrl_div_approx(A.f0[0][k], (float)t.sphere_volume);        // float volume, fast-math division (:199-201)
    } else if constexpr (MODE == NRL_OBS_ECCENTRICITY) {
      if (A.n[k] > 0 && A.f3[k] != 0.0f) {                               // :289-298
        val[0][k] = nrl_eccentricity(px, py, pz[k], A.f0[0][k], A.f1[k], A.f2[k], A.f3[k]);
      } else {
        val[0][k] = t.cutoff;
      }
    } else {
      val[0][k] = A.f0[0][k];
    }
  }
}

template <int MODE, int NWT>
__global__ void __launch_bounds__(NRL_NT, 5) nrl_cell_tile_kernel(const NrlCellTileParams p) {
  constexpr int S = NrlRecStride<MODE, NWT>::value;
  extern __shared__ __align__(16) unsigned char nrl_smem[];
  // layout: [stage: NRL_STAGE_CAP float4][records: NRL_NW * wcap * S float4][series: ser_cap floats]
  float4 *s_stage = reinterpret_cast<float4 *>(nrl_smem);
  float4 *s_rec = s_stage + NRL_STAGE_CAP;
  float4 *s_ser = s_rec + (size_t)NRL_NW * p.wcap * S;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_run_first[NRL_MAX_RUNS], s_run_last[NRL_MAX_RUNS];    // cells of a run (frame independent)
  __shared__ int s_run_beg[NRL_MAX_RUNS], s_run_dst[NRL_MAX_RUNS + 1];   // per frame: first binned atom, first staged slot
  __shared__ int s_total;                                                 // per frame: atoms of all runs
  __shared__ int s_n_runs;

  // a warp owns a compact 8x4x4 block of the tile: 32 lanes x 4 z-consecutive voxels, so a record
  // carries ONE float4 of per-plane Gaussian factors and the dilated volume (the list length) stays small
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lx = lane >> 2, ly = (lane & 3) + 4 * (warp & 1), lzq = warp >> 1;
  float4 *my_rec = s_rec + (size_t)warp * p.wcap * S;
  float4 *g_ser = reinterpret_cast<float4 *>(p.ser_scratch + (long long)blockIdx.x * p.ser_scratch_stride);
  const int wcap = p.wcap;

  if (tid == 0) nrl_mbar_init(&s_bar, 1);
  __syncthreads();
  unsigned phase = 0;

  for (int work = blockIdx.x; work < p.n_work; work += gridDim.x) {
    int lo = 0, hi = p.n_tasks - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (__ldg(&p.tasks[mid].tile_begin) <= work) lo = mid; else hi = mid - 1;
    }
    const NrlTask &t = p.tasks[lo];
    const int tile = work - t.tile_begin;
    const int nx = t.nx, ny = t.ny, nz = t.nz;
    const int tzi = tile % t.ntz, tyi = (tile / t.ntz) % t.nty, txi = t.tx0 + tile / (t.ntz * t.nty);
    const int x0 = txi * NRL_TX, y0 = tyi * NRL_TY, z0 = tzi * NRL_TZ;
    const float s = t.spacing, c = t.cutoff;
    const float c2 = c * c;
    const int nF = t.n_frames;
    const long long fd0 = t.frame0;

    // the warp's block of voxel centres and the inflated squared cutoff of the candidate test
    NrlTileBox wbox;
    wbox.x0 = (float)min(x0, nx - 1) * s;                 wbox.x1 = (float)min(x0 + NRL_TX - 1, nx - 1) * s;
    wbox.y0 = (float)min(y0 + 4 * (warp & 1), ny - 1) * s; wbox.y1 = (float)min(y0 + 4 * (warp & 1) + 3, ny - 1) * s;
    wbox.z0 = (float)min(z0 + 4 * lzq, nz - 1) * s;        wbox.z1 = (float)min(z0 + 4 * lzq + 3, nz - 1) * s;
    wbox.r2 = c2 * 1.0005f + 1e-30f;

    const int ix = x0 + lx, iy = y0 + ly, iz = z0 + lzq * NRL_VPT;
    const float px = (float)ix * s, py = (float)iy * s;
    float pz[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) pz[k] = (float)(iz + k) * s;

    constexpr bool FUSED = MODE == NRL_OBS_FUSED;
    constexpr int NTAB = FUSED ? 2 : NWT;                 // weight tables a record reads
    constexpr int NV = NrlValCount<MODE, NWT>::value;     // per-frame values of a sweep
    const int n_w = t.n_w;
    const int n_ch = t.n_ch;
    const float *wtabs[NTAB];
#pragma unroll
    for (int q = 0; q < NTAB; ++q) wtabs[q] = t.wtab[q];
    const bool streaming = nrl_streaming_task(t);
    constexpr int NRACC = FUSED ? NRL_MAX_CH : 1;         // recurrences a thread may carry (see nrl_streaming_task)
    float racc[NRACC][NRL_VPT];
#pragma unroll
    for (int ch = 0; ch < NRACC; ++ch)
#pragma unroll
      for (int k = 0; k < NRL_VPT; ++k) racc[ch][k] = 0.0f;
    const bool ser_in_smem = (long long)nrl_series_rows(t) * nF * NRL_TILE_VOX <= (long long)p.ser_cap;
    float4 *ser = ser_in_smem ? s_ser : g_ser;

    // ---- runs of this tile: one per (cx, cy) cell column its search box touches, plus the NaN bucket
    const NrlCellGrid g = p.grids[fd0];
    const int ncell = g.ncx * g.ncy * g.ncz;
    if (warp == 0) {
      const float reach = c * 1.001f + 1e-3f;
      const float bx1 = (float)min(x0 + NRL_TX - 1, nx - 1) * s, by1 = (float)min(y0 + NRL_TY - 1, ny - 1) * s,
                  bz1 = (float)min(z0 + NRL_TZ - 1, nz - 1) * s;
      const int cxl = max((int)floorf(nrl_cell_coord((float)x0 * s - reach, g.ox, g.inv_h)), 0);
      const int cxh = min((int)floorf(nrl_cell_coord(bx1 + reach, g.ox, g.inv_h)), g.ncx - 1);
      const int cyl = max((int)floorf(nrl_cell_coord((float)y0 * s - reach, g.oy, g.inv_h)), 0);
      const int cyh = min((int)floorf(nrl_cell_coord(by1 + reach, g.oy, g.inv_h)), g.ncy - 1);
      const int czl = max((int)floorf(nrl_cell_coord((float)z0 * s - reach, g.oz, g.inv_hz)), 0);
      const int czh = min((int)floorf(nrl_cell_coord(bz1 + reach, g.oz, g.inv_hz)), g.ncz - 1);
      const int nxs = max(cxh - cxl + 1, 0), nys = max(cyh - cyl + 1, 0);
      const int nreg = (czh >= czl) ? min(nxs * nys, NRL_MAX_RUNS - 1) : 0;   // the host sizes cells so that nxs*nys <= 31
      if (lane < nreg) {
        const int cx = cxl + lane / nys, cy = cyl + lane % nys;
        s_run_first[lane] = (cx * g.ncy + cy) * g.ncz + czl;
        s_run_last[lane] = (cx * g.ncy + cy) * g.ncz + czh;
      } else if (lane == nreg) {
        s_run_first[lane] = ncell;
        s_run_last[lane] = ncell;
      }
      if (lane == 0) s_n_runs = nreg + 1;
      __syncwarp();
    }

    // Per-frame descriptors are affine in the frame index inside a crop (the host lays the frames of a crop out
    // back to back): two loads per tile instead of a dependent load per frame.
    const long long off0 = p.frames.off[fd0], cs0 = p.grids[fd0].cs_off;
    const long long off_step = nF > 1 ? p.frames.off[fd0 + 1] - off0 : 0, cs_step = nF > 1 ? p.grids[fd0 + 1].cs_off - cs0 : 0;
    const int pb0 = p.frames.pbase[fd0], pb_step = nF > 1 ? p.frames.pbase[fd0 + 1] - pb0 : 0;

    // Staging of frame f (warp 0), software-pipelined so that no load sits on the critical path of the frame loop:
    // fetch_runs(f) only ISSUES the loads of the frame's cell-prefix entries (two per run), issue(f) consumes them
    // one frame later: sizes -> slots by a warp scan, one bulk copy per run.
    int nx_beg = 0, nx_end = 0;
    auto fetch_runs = [&](const int f) {
      const int *row = p.cell_start + (cs0 + (long long)f * cs_step);
      nx_beg = 0; nx_end = 0;
      if (lane < s_n_runs) { nx_beg = __ldg(row + s_run_first[lane]); nx_end = __ldg(row + s_run_last[lane] + 1); }
    };
    auto issue = [&](const int f) {
      const int beg = nx_beg, len = nx_end - nx_beg;
      int incl = len;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      const int dst = incl - len;
      s_run_beg[lane] = beg;
      s_run_dst[lane] = dst;
      if (lane == 31) { s_run_dst[32] = total; s_total = total; }
      __syncwarp();
      const bool staged = total <= NRL_STAGE_CAP;
      if (lane == 0) nrl_mbar_expect_tx(&s_bar, staged ? (uint32_t)total * 16u : 0u);
      __syncwarp();
      if (staged && len > 0)
        nrl_bulk_g2s(s_stage + dst, p.frames.atoms + (off0 + (long long)f * off_step) + beg, (uint32_t)len * 16u, &s_bar);
    };
    if (warp == 0) {
      fetch_runs(0);
      issue(0);
      if (nF > 1) fetch_runs(1);
    }

    for (int f = 0; f < nF; ++f) {
      const long long aoff = off0 + (long long)f * off_step;
      nrl_mbar_wait(&s_bar, phase & 1u);
      ++phase;
      const int total = s_total;
      const int nan_begin = s_run_dst[s_n_runs - 1];       // the last run is the NaN bucket
      const bool staged = total <= NRL_STAGE_CAP;
      const float4 *gat = p.frames.atoms + aoff;
      const float *inv = (MODE == NRL_MODE_VOXELIZE) ? t.inv_norm + aoff - (pb0 + f * pb_step) : nullptr;    // addressed by payload

      // position of staged / flat slot i in the binned frame
      auto locate = [&](const int i) {
        int a = 0, b = s_n_runs - 1;
        while (a < b) { const int m = (a + b + 1) >> 1; if (s_run_dst[m] <= i) a = m; else b = m - 1; }
        return s_run_beg[a] + (i - s_run_dst[a]);
      };

      NrlAcc<MODE, NWT> A;
      A.clear();
      int nrec = 0;
      bool saturated = false;
      // Second half of the compaction + one sweep round.  The compaction below only moves (x, y, z,
      // payload) records; here every lane resolves the weights of its share of the records with
      // independent gathers (one exposed L2 round trip per round instead of one per 32 staged atoms),
      // the voxeliser adds 1/S and its four per-plane Gaussian factors, then the list is padded to a
      // multiple of four with far-away records and swept.
      auto flush = [&]() {
        __syncwarp();
        if constexpr (FUSED) {
          if constexpr (NWT != 0) {
            for (int r0 = 0; r0 < nrec; r0 += 128) {
              float4 a[4];
              float w[4], rk[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * 32 + lane;
                a[u] = r < nrec ? my_rec[(size_t)r * S] : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * 32 + lane;
                const int payload = __float_as_int(a[u].w);
                w[u] = (r < nrec && (NWT & (NRL_FUSE_WS | NRL_FUSE_R | NRL_FUSE_COM)) != 0) ? __ldg(wtabs[0] + payload) : 0.0f;
                rk[u] = (r < nrec && (NWT & NRL_FUSE_SEEN) != 0) ? __ldg(wtabs[1] + payload) : -1.0f;
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int r = r0 + u * 32 + lane;
                if (r >= nrec) continue;
                float4 *slot = my_rec + (size_t)r * S;
                a[u].w = w[u];
                slot[0] = a[u];
                if constexpr ((NWT & NRL_FUSE_SEEN) != 0) {
                  const int rank = (int)rk[u];
                  // bit 31: "an atom" (ranks of a fused group are <= 30, checked by the host side of the call)
                  slot[1] = make_float4(__int_as_float((int)((rank >= 0 ? (1u << (rank & 31)) : 0u) | 0x80000000u)), 0.f, 0.f, 0.f);
                }
              }
            }
          }
        } else if constexpr (NrlModeTraits<MODE>::kNeedsWeight) {
          for (int r0 = 0; r0 < nrec; r0 += 128) {
            float4 a[4];
            float w[4][NWT];
            float iv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * 32 + lane;
              a[u] = r < nrec ? my_rec[(size_t)r * S] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * 32 + lane;
              const int payload = __float_as_int(a[u].w);
#pragma unroll
              for (int q = 0; q < NWT; ++q) w[u][q] = (r < nrec && q < n_w) ? __ldg(wtabs[q] + payload) : 0.0f;
              if constexpr (MODE == NRL_MODE_VOXELIZE)
                iv[u] = r < nrec ? __ldg(inv + payload) : 0.0f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int r = r0 + u * 32 + lane;
              if (r >= nrec) continue;
              float4 *slot = my_rec + (size_t)r * S;
              if constexpr (MODE == NRL_MODE_VOXELIZE) {
#pragma unroll
                for (int q = 0; q < NWT; ++q) w[u][q] = w[u][q] * iv[u];
                float e[NRL_VPT];
#pragma unroll
                for (int k = 0; k < NRL_VPT; ++k) { const float dz = a[u].z - pz[k]; e[k] = nrl_ex2_approx(dz * dz * t.k2); }
                slot[1] = make_float4(e[0], e[1], e[2], e[3]);
                if constexpr (NWT > 1) slot[2] = make_float4(w[u][0], w[u][1 % NWT], w[u][2 % NWT], w[u][3 % NWT]);
              }
              if constexpr (MODE == NRL_OBS_DISTINCT_COUNT) {
                // the weight table of this mode holds the RANK of the atom's value among the distinct values
                // (-1: the 99999.0 placeholder, never counted, marching_observers.cu:118-138)
                const int rank = (int)w[u][0];
                w[u][0] = __int_as_float(rank >= 0 ? (int)(1u << (rank & 31)) : 0);
              }
              a[u].w = w[u][0];
              slot[0] = a[u];
            }
          }
        }
        if (lane < ((4 - (nrec & 3)) & 3)) {
          float4 *slot = my_rec + (size_t)(nrec + lane) * S;
          slot[0] = make_float4(NRL_FAR, NRL_FAR, NRL_FAR, 0.f);
#pragma unroll
          for (int q = 1; q < S; ++q) slot[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        nrec = (nrec + 3) & ~3;
        __syncwarp();
        if (!saturated) saturated = nrl_rec_sweep<MODE, NWT>(my_rec, nrec, t, c2, px, py, pz, A);
        __syncwarp();
        nrec = 0;
      };

      // ---- the warp's own candidate records: (x, y, z, payload) of the staged atoms that can reach its block
      for (int c0 = 0; c0 < total; c0 += 32) {
        const int i = c0 + lane;
        int gi = 0;
        float4 a = make_float4(NRL_FAR, NRL_FAR, NRL_FAR, 0.f);
        if (i < total) {
          if (staged) a = s_stage[i];
          else { gi = locate(i); a = __ldg(gat + gi); }
        }
        const bool keep = i < total && (i >= nan_begin || nrl_near_box(a, wbox));   // the NaN bucket is in range of everything
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (m == 0u) continue;
        if (nrec + __popc(m) > wcap - 4) flush();
        if (keep) {
          float4 *slot = my_rec + (size_t)(nrec + __popc(m & ((1u << lane) - 1u))) * S;
          slot[0] = a;
        }
        nrec += __popc(m);
      }
      __syncthreads();                       // every warp is done with the staged cells
      if (warp == 0 && f + 1 < nF) {
        issue(f + 1);
        if (f + 2 < nF) fetch_runs(f + 2);
      }
      flush();

      float val[NV][NRL_VPT];
      nrl_rec_finish<MODE, NWT>(A, t, px, py, pz, val);
      // value row of channel ch: its own (fused observers), its weight slot (voxeliser), else the only one
      auto row_of = [&](const int ch) { return FUSED ? ch : (NWT > 1 ? t.ch_w[ch] : 0); };

      if (t.ch_agg[0] == 0) {
        // static channels: the single frame is the result
        if (ix < nx && iy < ny) {
          for (int ch = 0; ch < n_ch; ++ch) {
            float *o = p.out + t.ch_out[ch] + ((long long)ix * ny + iy) * nz;
            const int row = row_of(ch);
#pragma unroll
            for (int q = 0; q < NV; ++q)
              if (q == row) {
#pragma unroll
                for (int k = 0; k < NRL_VPT; ++k)
                  if (iz + k < nz) o[iz + k] = val[q][k];
              }
          }
        }
      } else if (streaming) {
        // mean / max / min need no series: the reference's own sequential recurrences (sum_device,
        // max_device, min_device, gpuutils.cuh:11-42) run as the frames pass, one per channel
#pragma unroll
        for (int ch = 0; ch < NRACC; ++ch) {
          if (ch >= n_ch) break;
          const int agg = t.ch_agg[ch], row = row_of(ch);
#pragma unroll
          for (int k = 0; k < NRL_VPT; ++k) {
            float x = 0.0f;
#pragma unroll
            for (int q = 0; q < NV; ++q) if (q == row) x = val[q][k];
            if (agg == 1) racc[ch][k] = racc[ch][k] + x;
            else if (f == 0) racc[ch][k] = x;
            else if (agg == 5) { if (x > racc[ch][k]) racc[ch][k] = x; }
            else { if (x < racc[ch][k]) racc[ch][k] = x; }
          }
        }
      } else {
        const int rows = nrl_series_rows(t);
#pragma unroll
        for (int q = 0; q < NV; ++q)
          if (q < rows) ser[((size_t)q * nF + f) * NRL_NT + tid] = make_float4(val[q][0], val[q][1], val[q][2], val[q][3]);
      }
    }

    if (streaming) {
      if (ix < nx && iy < ny) {
#pragma unroll
        for (int ch = 0; ch < NRACC; ++ch) {
          if (ch >= n_ch) break;
          float *o = p.out + t.ch_out[ch] + ((long long)ix * ny + iy) * nz;
#pragma unroll
          for (int k = 0; k < NRL_VPT; ++k)
            if (iz + k < nz) o[iz + k] = t.ch_agg[ch] == 1 ? nrl_div_approx(racc[ch][k], (float)nF) : racc[ch][k];   // mean_device :45-50
        }
      }
    } else if (t.ch_agg[0] != 0) {
      // K3 fused: every thread reduces the series of its own four voxels (it wrote them), once per
      // channel.  The median sorts its series in place, so channels that need it go last.
      for (int pass = 0; pass < 2; ++pass)
        for (int ch = 0; ch < n_ch; ++ch) {
          const int agg = t.ch_agg[ch];
          if ((agg == 3) != (pass == 1)) continue;
          const int row = FUSED ? ch : t.ch_w[ch];
          const float4 r = nrl_reduce4(ser + (size_t)row * nF * NRL_NT + tid, NRL_NT, nF, agg);
          if (ix < nx && iy < ny) {
            float *o = p.out + t.ch_out[ch] + ((long long)ix * ny + iy) * nz;
            if (iz + 0 < nz) o[iz + 0] = r.x;
            if (iz + 1 < nz) o[iz + 1] = r.y;
            if (iz + 2 < nz) o[iz + 2] = r.z;
            if (iz + 3 < nz) o[iz + 3] = r.w;
          }
        }
    }
    __syncthreads();     // the run table of the next tile is written by warp 0
  }
}
