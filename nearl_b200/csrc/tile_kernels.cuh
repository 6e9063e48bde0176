// nearl_b200 -- the gather kernels of the hot path (K1 voxeliser, K2 observers, fused K3).
//
// One CTA owns an 8x8x8 tile of voxels of one task and walks the frames of the task's
// window.  Per frame it (1) scans the frame's packed atoms with coalesced float4 loads
// and compacts, ORDER-PRESERVING, the atoms whose cutoff sphere can touch the tile into
// a shared-memory candidate list, (2) lets every thread sweep that list for its four
// z-consecutive voxels with the reference's exact squared-distance arithmetic, and
// (3) parks the per-frame value in a shared-memory time series.  After the last frame
// the series is reduced in place by the reference's reducer (K3) and one float per voxel
// is written.  No global atomics, no [F,G] intermediate in HBM, no per-frame launch.
//
// Replaces: frame_interp_global      /root/reference/src/voxelize.cu:38-130 (deposit half)
//           marching_observer_global /root/reference/src/marching_observers.cu:405-429
//           the nine observables     /root/reference/src/marching_observers.cu:30-355
//           gridwise_aggregation_global /root/reference/src/gpuutils.cu:48-77 (fused)
#pragma once
#include "common.cuh"
#include "reducers.cuh"

// ---- squared distances, exactly as the reference binary evaluates them -------------
// voxeliser (src/voxelize.cu:121-123): d = atom - gridpoint, fma(dz,dz, fma(dx,dx, dy*dy))
__device__ __forceinline__ float nrl_vox_partial(float ax, float ay, float gx, float gy) {
  const float dx = ax - gx, dy = ay - gy;
  return fmaf(dx, dx, dy * dy);
}
// observers (src/gpuutils.cuh:338-345): d = observer - atom, fma(dz,dz, fma(dy,dy, dx*dx))
__device__ __forceinline__ float nrl_obs_partial(float ox, float oy, float ax, float ay) {
  const float dx = ox - ax, dy = oy - ay;
  return fmaf(dy, dy, dx * dx);
}
__device__ __forceinline__ float nrl_obs_d2(float ox, float oy, float oz, float ax, float ay, float az) {
  const float dz = oz - az;
  return fmaf(dz, dz, nrl_obs_partial(ox, oy, ax, ay));
}

#define NRL_SCAN_ILP 4
#define NRL_FAR 3.0e38f

// Candidate test of the scan: squared distance from the atom to the box spanned by the tile's
// voxel centres against a slightly inflated cutoff.  Conservative for the exact float predicates
// of the sweep (the margin dwarfs their rounding), ~30 % tighter than box-plus-cutoff on every
// axis.  An atom with a NaN coordinate must survive the scan: the observers count it as in range for
// every voxel (d2 > c2 is false, marching_observers.cu:40), the voxeliser's d2 < c2 rejects it later.
struct NrlTileBox { float x0, x1, y0, y1, z0, z1, r2; };
__device__ __forceinline__ bool nrl_near_box(const float4 a, const NrlTileBox &b) {
  const float dx = fmaxf(fmaxf(b.x0 - a.x, a.x - b.x1), 0.0f);
  const float dy = fmaxf(fmaxf(b.y0 - a.y, a.y - b.y1), 0.0f);
  const float dz = fmaxf(fmaxf(b.z0 - a.z, a.z - b.z1), 0.0f);
  return !(fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > b.r2);
}
// the same, but an atom with ANY NaN coordinate is a candidate wherever its other coordinates are
// (fmaxf drops the NaN axis only); inf - inf false positives are settled by the exact predicate later
__device__ __forceinline__ bool nrl_near_box_nan(const float4 a, const NrlTileBox &b) {
  const float sum = (a.x + a.y) + a.z;
  return nrl_near_box(a, b) || (sum != sum);
}

// shared-memory record per candidate: the atom (x, y, z, weight) and, for the voxeliser, the
// eight per-plane factors exp(-(z_atom - z_plane)^2 / 2 sigma^2) of the tile's z planes
// (+ one float4 of per-slot weights when several weight tables are deposited together)
template <int MODE, int NWT> struct NrlCandStride {
  static constexpr int value = (MODE == NRL_MODE_VOXELIZE) ? (NWT > 1 ? 4 : 3) : 1;
};

template <int MODE> struct NrlModeTraits {
  static constexpr bool kNeedsWeight = (MODE != NRL_OBS_EXISTENCE && MODE != NRL_OBS_DIRECT_COUNT);
};

// Loop over the candidates of a sweep.  INDEXED: `idx` is the warp's sub-list of candidate numbers,
// padded to a multiple of four with the number of a far-away sentinel record; four numbers come in
// one 8-byte shared load and the four candidate records are fetched back to back (ILP over the
// dependent index -> record loads).  Inside BODY the candidate number is `ci`.
#define NRL_CAND_LOOP(...)                                                                     \
  if constexpr (INDEXED) {                                                                     \
    for (int c4 = 0; c4 < ncand; c4 += 4) {                                                    \
      const ushort4 q4 = *reinterpret_cast<const ushort4 *>(idx + c4);                         \
      const int ii[4] = {q4.x, q4.y, q4.z, q4.w};                                              \
      _Pragma("unroll") for (int u = 0; u < 4; ++u) { const int ci = ii[u]; __VA_ARGS__ }      \
    }                                                                                          \
  } else {                                                                                     \
    _Pragma("unroll 2") for (int ci = 0; ci < ncand; ++ci) { __VA_ARGS__ }                     \
  }

// ---- per-frame sweep of the candidate list for the 4 voxels of one thread ----------
// `idx` (optional): the warp's own sub-list of candidate numbers, see the tile kernel
template <int MODE, int NWT, bool INDEXED, typename CandPtr>
__device__ __forceinline__ void nrl_sweep(CandPtr cand, const unsigned short *idx, const int ncand, const NrlTask &t, const float c2,
                                          const float px, const float py, const float (&pz)[NRL_VPT],
                                          float (&vals)[NWT][NRL_VPT]) {
  float (&val)[NRL_VPT] = vals[0];
  if constexpr (MODE == NRL_MODE_VOXELIZE && NWT > 1) {
    // several weight tables at once: geometry, cutoff test and exponentials are shared
    float acc[NWT][NRL_VPT];
#pragma unroll
    for (int w = 0; w < NWT; ++w)
#pragma unroll
      for (int k = 0; k < NRL_VPT; ++k) acc[w][k] = 0.0f;
    const float k2 = t.k2;
    const int zq = 1 + (threadIdx.x & 1);
    NRL_CAND_LOOP(
      const float4 a = cand[ci * 4];
      const float4 ezq = cand[ci * 4 + zq];
      const float4 ww = cand[ci * 4 + 3];
      const float part = nrl_vox_partial(a.x, a.y, px, py);
      const float exy = nrl_ex2_approx(part * k2);
      const float ez[NRL_VPT] = {ezq.x, ezq.y, ezq.z, ezq.w};
      const float wxy[4] = {exy * ww.x, exy * ww.y, exy * ww.z, exy * ww.w};     // same association as the single-table form
_Pragma("unroll")
      for (int k = 0; k < NRL_VPT; ++k) {
        const float dz = a.z - pz[k];
        const float d2 = fmaf(dz, dz, part);
        if (d2 < c2) {
_Pragma("unroll")
          for (int w = 0; w < NWT; ++w) acc[w][k] = fmaf(wxy[w], ez[k], acc[w][k]);
        }
      }
    )
#pragma unroll
    for (int w = 0; w < NWT; ++w)
#pragma unroll
      for (int k = 0; k < NRL_VPT; ++k) vals[w][k] = acc[w][k];
  } else if constexpr (MODE == NRL_MODE_VOXELIZE) {
    // Gaussian in separable form: exp(-d2/2s^2) = E_xy * E_z; E_z of the tile's z planes was
    // tabulated per candidate by the scan, so one exponential per (candidate, thread) remains.
    // The in/out decision still uses the reference's exact d2 (voxelize.cu:121-125).
    float acc[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) acc[k] = 0.0f;
    const float k2 = t.k2;
    const int zq = 1 + (threadIdx.x & 1);
    NRL_CAND_LOOP(
      const float4 a = cand[ci * 3];
      const float4 ezq = cand[ci * 3 + zq];
      const float part = nrl_vox_partial(a.x, a.y, px, py);
      const float wxy = nrl_ex2_approx(part * k2) * a.w;
      const float ez[NRL_VPT] = {ezq.x, ezq.y, ezq.z, ezq.w};
_Pragma("unroll")
      for (int k = 0; k < NRL_VPT; ++k) {
        const float dz = a.z - pz[k];
        const float d2 = fmaf(dz, dz, part);
        if (d2 < c2) acc[k] = fmaf(wxy, ez[k], acc[k]);
      }
    )
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) val[k] = acc[k];
  } else if constexpr (MODE == NRL_OBS_EXISTENCE || MODE == NRL_OBS_DIRECT_COUNT) {
    float acc[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) acc[k] = 0.0f;
    if constexpr (MODE == NRL_OBS_EXISTENCE && INDEXED) {
      // existence saturates: once every voxel of the warp has seen an atom the rest of the list cannot
      // change the result (inside a protein that happens within the first few candidates)
      for (int c4 = 0; c4 < ncand; c4 += 4) {
        const ushort4 q4 = *reinterpret_cast<const ushort4 *>(idx + c4);
        const int ii[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 a = cand[ii[u]];
          const float part = nrl_obs_partial(px, py, a.x, a.y);
#pragma unroll
          for (int k = 0; k < NRL_VPT; ++k) {
            const float dz = pz[k] - a.z;
            const float d2 = fmaf(dz, dz, part);
            if (!(d2 > c2)) acc[k] = 1.0f;                              // marching_observers.cu:40
          }
        }
        if (__all_sync(0xffffffffu, (acc[0] + acc[1]) + (acc[2] + acc[3]) == 4.0f)) break;
      }
    } else {
      NRL_CAND_LOOP(
        const float4 a = cand[ci];
        const float part = nrl_obs_partial(px, py, a.x, a.y);
_Pragma("unroll")
        for (int k = 0; k < NRL_VPT; ++k) {
          const float dz = pz[k] - a.z;
          const float d2 = fmaf(dz, dz, part);
          if (!(d2 > c2)) {                                             // marching_observers.cu:40,77
            if constexpr (MODE == NRL_OBS_EXISTENCE) acc[k] = 1.0f; else acc[k] = acc[k] + 1.0f;
          }
        }
      )
    }
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) val[k] = acc[k];
  } else if constexpr (MODE == NRL_OBS_DISTINCT_COUNT) {
    // candidates arrive sorted by weight value (the pack stage orders atoms by value and
    // the compaction preserves order), so "distinct" is "differs from the last in-range
    // value"; the reference's 1000-slot table (marching_observers.cu:118-140) caps the
    // count at 1000 whatever the visiting order.
    float cnt[NRL_VPT], last[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) { cnt[k] = 0.0f; last[k] = __int_as_float(0x7fc00000); }
    NRL_CAND_LOOP(
      const float4 a = cand[ci];
      const float part = nrl_obs_partial(px, py, a.x, a.y);
_Pragma("unroll")
      for (int k = 0; k < NRL_VPT; ++k) {
        const float dz = pz[k] - a.z;
        const float d2 = fmaf(dz, dz, part);
        if (!(d2 > c2)) {
          if (a.w != last[k]) cnt[k] = cnt[k] + 1.0f;
          last[k] = a.w;
        }
      }
    )
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) val[k] = fminf(cnt[k], NRL_DISTINCT_LIMIT);
  } else if constexpr (MODE == NRL_OBS_MEAN_DISTANCE) {
    float r[NRL_VPT], ws[NRL_VPT];
    int n[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) { r[k] = 0.0f; ws[k] = 0.0f; n[k] = 0; }
    NRL_CAND_LOOP(
      const float4 a = cand[ci];
      const float part = nrl_obs_partial(px, py, a.x, a.y);
_Pragma("unroll")
      for (int k = 0; k < NRL_VPT; ++k) {
        const float dz = pz[k] - a.z;
        const float d2 = fmaf(dz, dz, part);
        if (!(d2 > c2)) {                                               // :160-167
          r[k] = fmaf(nrl_sqrt_approx(d2), a.w, r[k]);
          ws[k] = ws[k] + a.w;
          n[k] += 1;
        }
      }
    )
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) val[k] = n[k] > 0 ? nrl_div_approx(r[k], ws[k]) : 0.0f;
  } else if constexpr (MODE == NRL_OBS_CUMULATIVE_WEIGHT || MODE == NRL_OBS_DENSITY) {
    float ws[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) ws[k] = 0.0f;
    NRL_CAND_LOOP(
      const float4 a = cand[ci];
      const float part = nrl_obs_partial(px, py, a.x, a.y);
_Pragma("unroll")
      for (int k = 0; k < NRL_VPT; ++k) {
        const float dz = pz[k] - a.z;
        const float d2 = fmaf(dz, dz, part);
        if (!(d2 > c2)) ws[k] = ws[k] + a.w;                            // :183-189
      }
    )
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) {
      if constexpr (MODE == NRL_OBS_DENSITY) val[k] = nrl_div_approx(ws[k], (float)t.sphere_volume);  // float volume, fast-math division (:199-201)
      else val[k] = ws[k];
    }
  } else if constexpr (MODE == NRL_OBS_ECCENTRICITY || MODE == NRL_OBS_RADIUS_OF_GYRATION) {
    float cx[NRL_VPT], cy[NRL_VPT], cz[NRL_VPT], ws[NRL_VPT];
    int n[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) { cx[k] = cy[k] = cz[k] = ws[k] = 0.0f; n[k] = 0; }
    NRL_CAND_LOOP(
      const float4 a = cand[ci];
      const float part = nrl_obs_partial(px, py, a.x, a.y);
_Pragma("unroll")
      for (int k = 0; k < NRL_VPT; ++k) {
        const float dz = pz[k] - a.z;
        const float d2 = fmaf(dz, dz, part);
        if (!(d2 > c2)) {                                               // :279-287, :322-332
          n[k] += 1;
          cx[k] = fmaf(a.w, a.x, cx[k]);
          cy[k] = fmaf(a.w, a.y, cy[k]);
          cz[k] = fmaf(a.w, a.z, cz[k]);
          ws[k] = ws[k] + a.w;
        }
      }
    )
    if constexpr (MODE == NRL_OBS_ECCENTRICITY) {
#pragma unroll
      for (int k = 0; k < NRL_VPT; ++k) {
        if (n[k] > 0 && ws[k] != 0.0f) {                                // :289-298
          val[k] = nrl_eccentricity(px, py, pz[k], cx[k], cy[k], cz[k], ws[k]);
        } else {
          val[k] = t.cutoff;
        }
      }
    } else {
      float m0[NRL_VPT], m1[NRL_VPT], m2[NRL_VPT], r[NRL_VPT];
      bool live[NRL_VPT];
#pragma unroll
      for (int k = 0; k < NRL_VPT; ++k) {                               // :334-343
        live[k] = (n[k] > 1) && (ws[k] != 0.0f);
        m0[k] = nrl_div_approx(cx[k], ws[k]);
        m1[k] = nrl_div_approx(cy[k], ws[k]);
        m2[k] = nrl_div_approx(cz[k], ws[k]);
        r[k] = 0.0f;
      }
      NRL_CAND_LOOP(
        const float4 a = cand[ci];
        const float part = nrl_obs_partial(px, py, a.x, a.y);
_Pragma("unroll")
        for (int k = 0; k < NRL_VPT; ++k) {
          const float dz = pz[k] - a.z;
          const float d2 = fmaf(dz, dz, part);
          if (!(d2 > c2)) r[k] = fmaf(a.w, nrl_obs_d2(m0[k], m1[k], m2[k], a.x, a.y, a.z), r[k]);
        }
      )
#pragma unroll
      for (int k = 0; k < NRL_VPT; ++k)
        val[k] = live[k] ? nrl_sqrt_approx(nrl_div_approx(r[k], ws[k])) : 0.0f;
    }
  } else if constexpr (MODE == NRL_OBS_DISPERSION) {
    // pairs (j<k) of in-range atoms in candidate (= atom) order, marching_observers.cu:228-243
#pragma unroll 1
    for (int v = 0; v < NRL_VPT; ++v) {
      float r = 0.0f, ws = 0.0f;
      for (int j = 0; j < ncand; ++j) {
        const float4 a = cand[INDEXED ? (int)idx[j] : j];
        if (nrl_obs_d2(px, py, pz[v], a.x, a.y, a.z) > c2) continue;
        for (int m = j + 1; m < ncand; ++m) {
          const float4 b = cand[INDEXED ? (int)idx[m] : m];
          if (nrl_obs_d2(px, py, pz[v], b.x, b.y, b.z) > c2) continue;
          const float d2 = nrl_obs_d2(a.x, a.y, a.z, b.x, b.y, b.z);
          const float ww = a.w * b.w;
          r = fmaf(ww, nrl_sqrt_approx(d2), r);
          ws = ws + ww;
        }
      }
      val[v] = (ws == 0.0f) ? 0.0f : nrl_div_approx(r, ws);             // :244-248
    }
  }
}

// ---- the tile kernel ---------------------------------------------------------------
template <int MODE, int NWT>
__global__ void __launch_bounds__(NRL_NT) nrl_tile_kernel(const NrlTileParams p) {
  extern __shared__ __align__(16) unsigned char nrl_smem[];
  float4 *s_cand = reinterpret_cast<float4 *>(nrl_smem);
  float4 *s_ser = reinterpret_cast<float4 *>(nrl_smem + (size_t)p.cand_cap * NrlCandStride<MODE, NWT>::value * sizeof(float4));
  __shared__ int s_wcnt[NRL_NW];
  // per-warp sub-lists (candidate numbers) live behind the series in dynamic shared memory
  unsigned short *s_idx = reinterpret_cast<unsigned short *>(
      nrl_smem + (size_t)p.cand_cap * NrlCandStride<MODE, NWT>::value * sizeof(float4) + (size_t)p.ser_cap * sizeof(float));

  // a warp owns a compact 4x4x8 block of the tile (32 lanes x 4 z-consecutive voxels): of all
  // 128-voxel shapes it has the smallest cutoff-dilated volume, i.e. the shortest candidate list
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lzq = lane & 1, ly = ((lane >> 1) & 3) + 4 * (warp & 1), lx = ((lane >> 3) & 3) + 4 * (warp >> 1);
  unsigned short *my_idx = s_idx + (size_t)warp * p.cand_cap;
  float4 *g_cand = p.cand_scratch + (long long)blockIdx.x * p.cand_scratch_stride;
  float4 *g_ser = reinterpret_cast<float4 *>(p.ser_scratch + (long long)blockIdx.x * p.ser_scratch_stride);

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

    // box spanned by the centres of the tile's voxels, and the inflated squared cutoff of the scan
    NrlTileBox box;
    box.x0 = (float)x0 * s; box.x1 = (float)min(x0 + NRL_TX - 1, nx - 1) * s;
    box.y0 = (float)y0 * s; box.y1 = (float)min(y0 + NRL_TY - 1, ny - 1) * s;
    box.z0 = (float)z0 * s; box.z1 = (float)min(z0 + NRL_TZ - 1, nz - 1) * s;
    box.r2 = c2 * 1.0005f + 1e-30f;
    NrlTileBox wbox = box;      // the same test against the warp's own 4x4x8 block
    wbox.x0 = (float)min(x0 + 4 * (warp >> 1), nx - 1) * s; wbox.x1 = (float)min(x0 + 4 * (warp >> 1) + 3, nx - 1) * s;
    wbox.y0 = (float)min(y0 + 4 * (warp & 1), ny - 1) * s;  wbox.y1 = (float)min(y0 + 4 * (warp & 1) + 3, ny - 1) * s;

    const int ix = x0 + lx, iy = y0 + ly, iz = z0 + lzq * NRL_VPT;
    const float px = (float)ix * s, py = (float)iy * s;
    float pz[NRL_VPT];
#pragma unroll
    for (int k = 0; k < NRL_VPT; ++k) pz[k] = (float)(iz + k) * s;

    const int n_w = t.n_w;
    const float *wtabs[NWT];
#pragma unroll
    for (int q = 0; q < NWT; ++q) wtabs[q] = t.wtab[q];
    const bool ser_in_smem = (long long)n_w * nF * NRL_TILE_VOX <= (long long)p.ser_cap;
    float4 *ser = ser_in_smem ? s_ser : g_ser;

    for (int f = 0; f < nF; ++f) {
      const long long aoff = p.frames.off[t.frame0 + f];
      const int n = p.frames.cnt[t.frame0 + f];
      const float4 *atoms = p.frames.atoms + aoff;
      const float *inv = (MODE == NRL_MODE_VOXELIZE) ? t.inv_norm + aoff - p.frames.pbase[t.frame0 + f] : nullptr;   // addressed by payload

      // ---- pass 1: each warp counts the survivors of its contiguous atom segment.  Four
      // independent coalesced float4 loads are in flight per lane (the scan is latency bound).
      const int seg = ((n + NRL_NW - 1) / NRL_NW + 31) & ~31;
      const int a_begin = min(warp * seg, n), a_end = min(a_begin + seg, n);
      int mycnt = 0;
      for (int i0 = a_begin; i0 < a_end; i0 += 32 * NRL_SCAN_ILP) {
        float4 a[NRL_SCAN_ILP];
#pragma unroll
        for (int u = 0; u < NRL_SCAN_ILP; ++u) {
          const int i = i0 + u * 32 + lane;
          a[u] = i < a_end ? __ldg(atoms + i) : make_float4(NRL_FAR, NRL_FAR, NRL_FAR, 0.f);
        }
#pragma unroll
        for (int u = 0; u < NRL_SCAN_ILP; ++u)
          mycnt += __popc(__ballot_sync(0xffffffffu, nrl_near_box_nan(a[u], box)));
      }
      if (lane == 0) s_wcnt[warp] = mycnt;
      __syncthreads();
      int base = 0, total = 0;
#pragma unroll
      for (int w = 0; w < NRL_NW; ++w) {
        const int cw = s_wcnt[w];
        if (w < warp) base += cw;
        total += cw;
      }
      const bool cand_in_smem = total + 4 <= p.cand_cap;  // p.cand_cap counts candidates, whatever their record size; room for the sentinel
      float4 *cand = cand_in_smem ? s_cand : g_cand;

      // ---- pass 2: order-preserving compaction, weight resolved through the payload
      for (int i0 = a_begin; i0 < a_end; i0 += 32 * NRL_SCAN_ILP) {
        float4 a[NRL_SCAN_ILP];
        bool keep[NRL_SCAN_ILP];
#pragma unroll
        for (int u = 0; u < NRL_SCAN_ILP; ++u) {
          const int i = i0 + u * 32 + lane;
          a[u] = i < a_end ? __ldg(atoms + i) : make_float4(NRL_FAR, NRL_FAR, NRL_FAR, 0.f);
          keep[u] = nrl_near_box_nan(a[u], box);
        }
        float w[NRL_SCAN_ILP][NWT];
#pragma unroll
        for (int u = 0; u < NRL_SCAN_ILP; ++u) {
#pragma unroll
          for (int q = 0; q < NWT; ++q) w[u][q] = 0.0f;
          if (keep[u]) {
            const int payload = __float_as_int(a[u].w);
            if constexpr (NrlModeTraits<MODE>::kNeedsWeight) {
#pragma unroll
              for (int q = 0; q < NWT; ++q) if (q < n_w) w[u][q] = __ldg(wtabs[q] + payload);
            }
            if constexpr (MODE == NRL_MODE_VOXELIZE) {
              const float iv = __ldg(inv + payload);
#pragma unroll
              for (int q = 0; q < NWT; ++q) w[u][q] = w[u][q] * iv;
            }
            if constexpr (MODE == NRL_OBS_DISTINCT_COUNT) w[u][0] = w[u][0] + 0.0f;   // -0.0 == +0.0 in the reference's float compare
          }
        }
#pragma unroll
        for (int u = 0; u < NRL_SCAN_ILP; ++u) {
          const unsigned m = __ballot_sync(0xffffffffu, keep[u]);
          if (keep[u]) {
            a[u].w = w[u][0];
            float4 *slot = cand + (size_t)(base + __popc(m & ((1u << lane) - 1u))) * NrlCandStride<MODE, NWT>::value;
            slot[0] = a[u];
            if constexpr (MODE == NRL_MODE_VOXELIZE) {
              float e[NRL_TZ];
#pragma unroll
              for (int k = 0; k < NRL_TZ; ++k) { const float dz = a[u].z - (float)(z0 + k) * s; e[k] = nrl_ex2_approx(dz * dz * t.k2); }
              slot[1] = make_float4(e[0], e[1], e[2], e[3]);
              slot[2] = make_float4(e[4], e[5], e[6], e[7]);
              if constexpr (NWT > 1) slot[3] = make_float4(w[u][0], w[u][1 % NWT], w[u][2 % NWT], w[u][3 % NWT]);
            }
          }
          base += __popc(m);
        }
      }
      if (cand_in_smem && tid == 0) {
        // far-away sentinel record behind the list: pads the warps' sub-lists to multiples of four
#pragma unroll
        for (int q = 0; q < NrlCandStride<MODE, NWT>::value; ++q)
          s_cand[total * NrlCandStride<MODE, NWT>::value + q] =
              q == 0 ? make_float4(NRL_FAR, NRL_FAR, NRL_FAR, 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();

      // ---- sweep
      float val[NWT][NRL_VPT];
      if (cand_in_smem) {
        // second-level culling: each warp keeps (in order) only the candidates that can reach its block
        int nsub = 0;
        for (int c0 = 0; c0 < total; c0 += 32) {
          const int ci = c0 + lane;
          const bool k = ci < total && nrl_near_box_nan(s_cand[ci * NrlCandStride<MODE, NWT>::value], wbox);
          const unsigned m = __ballot_sync(0xffffffffu, k);
          if (k) my_idx[nsub + __popc(m & ((1u << lane) - 1u))] = (unsigned short)ci;
          nsub += __popc(m);
        }
        while (nsub & 3) { if (lane == 0) my_idx[nsub] = (unsigned short)total; ++nsub; }
        __syncwarp();
        nrl_sweep<MODE, NWT, true>(s_cand, my_idx, nsub, t, c2, px, py, pz, val);
      } else {
        nrl_sweep<MODE, NWT, false>(static_cast<const float4 *>(g_cand), nullptr, total, t, c2, px, py, pz, val);
      }

      if (t.ch_agg[0] == 0) {
        // static channels: the single frame is the result
        if (ix < nx && iy < ny) {
          for (int ch = 0; ch < t.n_ch; ++ch) {
            float *o = p.out + t.ch_out[ch] + ((long long)ix * ny + iy) * nz;
            const int slot = NWT > 1 ? t.ch_w[ch] : 0;
#pragma unroll
            for (int q = 0; q < NWT; ++q)
              if (q == slot) {
#pragma unroll
                for (int k = 0; k < NRL_VPT; ++k)
                  if (iz + k < nz) o[iz + k] = val[q][k];
              }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < NWT; ++q)
          if (q < n_w) ser[((size_t)q * nF + f) * NRL_NT + tid] = make_float4(val[q][0], val[q][1], val[q][2], val[q][3]);
      }
      __syncthreads();   // candidate list is rebuilt by the next frame
    }

    if (t.ch_agg[0] != 0) {
      // K3 fused: every thread reduces the series of its own four voxels (it wrote them), once per
      // channel.  The median sorts its series in place, so channels that need it go last.
      for (int pass = 0; pass < 2; ++pass)
        for (int ch = 0; ch < t.n_ch; ++ch) {
          const int agg = t.ch_agg[ch];
          if ((agg == 3) != (pass == 1)) continue;
          const float4 r = nrl_reduce4(ser + (size_t)t.ch_w[ch] * nF * NRL_NT + tid, NRL_NT, nF, agg);
          if (ix < nx && iy < ny) {
            float *o = p.out + t.ch_out[ch] + ((long long)ix * ny + iy) * nz;
            if (iz + 0 < nz) o[iz + 0] = r.x;
            if (iz + 1 < nz) o[iz + 1] = r.y;
            if (iz + 2 < nz) o[iz + 2] = r.z;
            if (iz + 3 < nz) o[iz + 3] = r.w;
          }
        }
    }
  }
}
