// nearl_b200 -- standalone K3 (`aggregate`) and `summation`.
//
// Every form reads the frame-major [F, G] series out of HBM exactly once:
//   nrl_aggregate_stream_kernel  mean / max / min / drift (recurrences over the frames): nothing staged
//   nrl_aggregate_reg_kernel     std / var / entropy over <= 64 frames: the series lives in registers
//   nrl_aggregate_pipe_kernel    std / var / entropy over longer series: persistent CTAs, ring of [F, W] tiles in
//                                shared memory, one 2-D tiled TMA copy per tile (UTMALDG) on an mbarrier
//   nrl_aggregate_kernel         one-shot staged tile (one cp.async.bulk per row, UBLKCP): the median, ragged G
// Replaces gridwise_aggregation_global + aggregate_host (/root/reference/src/gpuutils.cu:48-105),
// which copies every series into a 2 KB per-thread local-memory array first.
#pragma once
#include <cuda.h>   // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no link against libcuda)
#include "common.cuh"
#include "reducers.cuh"

__device__ __forceinline__ uint32_t nrl_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void nrl_mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nrl_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nrl_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nrl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nrl_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   nrl_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(nrl_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void nrl_mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = nrl_smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}


// One voxel's series out of a staged [n, W] tile (column v), reduced by one thread: standard deviation / variance
// (same arithmetic as nrl_reduce4, gpuutils.cuh:35-74), median (the reference's in-place exchange sort,
// median_device, gpuutils.cuh:227-245 -- the tile column is overwritten) and information entropy
// (information_entropy_histogram_device, gpuutils.cuh:140-173).
#define NRL_AGG_BLOCK 16
#define NRL_MEDIAN_SELECT_MIN 24
#define NRL_AGG_LEGACY_ENTROPY 107
__device__ __forceinline__ float nrl_agg_voxel(float *fb, const int W, const int v, const int n, const int agg) {
  const float fn = (float)n;
  if (agg == 2 || agg == 4) {
    // the additions are a dependent chain in the reference's order; the shared-memory reads are not: fetch a block
    // of them first, so that a thread has NRL_AGG_BLOCK loads in flight instead of one per addition
    const float *col = fb + v;
    float sum = 0.0f;
    int i = 0;
    for (; i + NRL_AGG_BLOCK <= n; i += NRL_AGG_BLOCK) {
      float x[NRL_AGG_BLOCK];
#pragma unroll
      for (int u = 0; u < NRL_AGG_BLOCK; ++u) x[u] = col[(i + u) * W];
#pragma unroll
      for (int u = 0; u < NRL_AGG_BLOCK; ++u) sum = sum + x[u];
    }
    for (; i < n; ++i) sum = sum + col[i * W];
    const float mean = nrl_div_approx(sum, fn);
    float acc = 0.0f;
    i = 0;
    for (; i + NRL_AGG_BLOCK <= n; i += NRL_AGG_BLOCK) {
      float d[NRL_AGG_BLOCK];
#pragma unroll
      for (int u = 0; u < NRL_AGG_BLOCK; ++u) d[u] = col[(i + u) * W] - mean;
#pragma unroll
      for (int u = 0; u < NRL_AGG_BLOCK; ++u) acc = fmaf(d[u], d[u], acc);
    }
    for (; i < n; ++i) { const float d = col[i * W] - mean; acc = fmaf(d, d, acc); }
    return agg == 2 ? nrl_sqrt_approx(nrl_div_approx(acc, fn)) : nrl_div_approx(acc, fn);
  }
  if (agg == 3) {
    // The median is an order statistic: any exact selection returns the value the reference's exchange sort
    // (median_device, gpuutils.cuh:227-245) leaves in the middle.  Series longer than NRL_MEDIAN_SELECT_MIN without a NaN
    // are selected by radix (4-bit digits of the order-preserving integer image of the floats, most significant first,
    // counted in packed 4-bit counters like the entropy histogram): ~60-100 operations per sample instead of n / 2
    // compare-exchanges.  A NaN compares false with everything, so where it ends up depends on the sort itself: those series
    // keep the reference's sort, as do short ones.
    unsigned *key = reinterpret_cast<unsigned *>(fb) + v;
    bool select = n >= NRL_MEDIAN_SELECT_MIN;
    if (select) {
      for (int i = 0; i < n; ++i) { const float x = fb[(size_t)i * W + v]; if (x != x) select = false; }
    }
    if (select) {
      for (int i = 0; i < n; ++i) {            // float -> unsigned with the same order (-0.0 sorts before +0.0: equal anyway)
        const unsigned u = key[(size_t)i * W];
        key[(size_t)i * W] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
      }
      const unsigned long long lanes = 0x000F000F000F000Full;
      unsigned sel[2];
      int n_sel = 0;
      const int k0 = (n & 1) ? n / 2 : n / 2 - 1;
      unsigned prefix = 0u;
      int kk = k0, group = n;                   // rank wanted inside the candidate group, size of the group
      for (int shift = 28; shift >= 0 && group > 1; shift -= 4) {
        const unsigned himask = shift == 28 ? 0u : ~((1u << (shift + 4)) - 1u);
        unsigned long long h0 = 0ull, h1 = 0ull, h2 = 0ull, h3 = 0ull;
        for (int i0 = 0; i0 < n; i0 += 15) {
          unsigned long long acc = 0ull;
          const int i1 = min(i0 + 15, n);
          for (int i = i0; i < i1; ++i) {
            const unsigned u = key[(size_t)i * W];
            if (((u ^ prefix) & himask) == 0u) acc += 1ull << (4 * ((u >> shift) & 15u));
          }
          h0 += acc & lanes; h1 += (acc >> 4) & lanes; h2 += (acc >> 8) & lanes; h3 += (acc >> 12) & lanes;
        }
        const unsigned long long h[4] = {h0, h1, h2, h3};
        int below = 0, d = 0, cnt = 0;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
          const int cb = (int)((h[b & 3] >> ((b >> 2) * 16)) & 0xffffull);
          if (cnt == 0 && kk < below + cb) { d = b; cnt = cb; } else if (cnt == 0) below += cb;
        }
        kk -= below;
        group = cnt;
        prefix |= (unsigned)d << shift;
        if (shift == 0) break;
        if (group == 1) {                       // one candidate left: fetch it
          const unsigned hm = ~((1u << shift) - 1u);
          for (int i = 0; i < n; ++i) { const unsigned u = key[(size_t)i * W]; if (((u ^ prefix) & hm) == 0u) prefix = u; }
        }
      }
      sel[n_sel++] = prefix;
      if ((n & 1) == 0) {
        // the next order statistic: the same value when the group of equals reaches past rank kk, else the smallest larger key
        unsigned next = 0xffffffffu;
        int equal = 0;
        for (int i = 0; i < n; ++i) {
          const unsigned u = key[(size_t)i * W];
          equal += u == prefix;
          if (u > prefix && u < next) next = u;
        }
        int smaller = 0;
        for (int i = 0; i < n; ++i) smaller += key[(size_t)i * W] < prefix;
        sel[n_sel++] = (k0 + 1 < smaller + equal) ? prefix : next;
      }
      float val[2];
      for (int q = 0; q < n_sel; ++q) {
        const unsigned u = sel[q];
        val[q] = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
      }
      return (n & 1) ? val[0] : nrl_div_approx(val[0] + val[1], 2.0f);
    }
    for (int i = 0; i < n; ++i) {
      float a = fb[(size_t)i * W + v];
      for (int j = i + 1; j < n; ++j) {
        const float b = fb[(size_t)j * W + v];
        if (a > b) { fb[(size_t)j * W + v] = a; a = b; }
      }
      fb[(size_t)i * W + v] = a;
    }
    return (n & 1) ? fb[(size_t)(n / 2) * W + v]
                   : nrl_div_approx(fb[(size_t)(n / 2 - 1) * W + v] + fb[(size_t)(n / 2) * W + v], 2.0f);
  }
  // information entropy: the histogram is ALU work.  Sixteen 4-bit counters in one 64-bit word take the increments
  // (one shifted add per sample); every 15 samples they are spilled into 16-bit counters (bin b -> word b & 3, lane b >> 2,
  // so that a spill is four masked adds).  NRL_AGG_LEGACY_ENTROPY (agg = 107, tests only) keeps the former form -- four
  // predicated 64-bit adds per sample -- to check the counts bit for bit.
  const bool legacy = agg == NRL_AGG_LEGACY_ENTROPY;
  float r = 0.0f;
  if (n > 1) {
    const float *col = fb + v;
    // (minimum and maximum are exact whatever the order: sixteen loads in flight, two chains each)
    float mn = col[0], mx = col[0];
    {
      float mn2 = mn, mx2 = mx;
      int i = 1;
      for (; i + 16 <= n; i += 16) {
        float x[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) x[u] = col[(i + u) * W];
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
          if (x[u] < mn) mn = x[u];
          if (x[u] > mx) mx = x[u];
          if (x[u + 1] < mn2) mn2 = x[u + 1];
          if (x[u + 1] > mx2) mx2 = x[u + 1];
        }
      }
      for (; i < n; ++i) { const float x = col[i * W]; if (x < mn) mn = x; if (x > mx) mx = x; }
      if (mn2 < mn) mn = mn2;
      if (mx2 > mx) mx = mx2;
    }
    const float range = mx - mn;
    unsigned long long h0 = 0ull, h1 = 0ull, h2 = 0ull, h3 = 0ull;      // 16 counters of 16 bits
    if (legacy) {
#pragma unroll 4
      for (int i = 0; i < n; ++i) {
        const float sc = nrl_div_approx(col[i * W] - mn, range) * 16.0f;
        int bin;
        asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(bin) : "f"(sc));
        bin = max(0, min(bin, NRL_ENTROPY_BINS - 1));
        const unsigned long long inc = 1ull << ((bin >> 2) * 16);
        const int word = bin & 3;
        h0 += (word == 0) ? inc : 0ull; h1 += (word == 1) ? inc : 0ull;
        h2 += (word == 2) ? inc : 0ull; h3 += (word == 3) ? inc : 0ull;
      }
    } else {
      const unsigned long long lanes = 0x000F000F000F000Full;
      for (int i0 = 0; i0 < n; i0 += 15) {
        // fifteen samples per spill; the full groups are unrolled completely: fifteen independent load -> bin chains in
        // flight per thread (a CTA of the pipelined kernel is alone on its SM: eight warps have to hide the latency of the
        // shared-memory loads and of the float -> int conversions themselves), three partial sums of the shifted ones
        unsigned long long acc = 0ull;
        if (i0 + 15 <= n) {
          unsigned long long part[3] = {0ull, 0ull, 0ull};
#pragma unroll
          for (int u = 0; u < 15; ++u) {
            const float sc = nrl_div_approx(col[(i0 + u) * W] - mn, range) * 16.0f;
            int bin;
            asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(bin) : "f"(sc));
            bin = max(0, min(bin, NRL_ENTROPY_BINS - 1));
            part[u % 3] += 1ull << (4 * bin);
          }
          acc = part[0] + part[1] + part[2];
        } else {
          for (int i = i0; i < n; ++i) {
            const float sc = nrl_div_approx(col[i * W] - mn, range) * 16.0f;
            int bin;
            asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(bin) : "f"(sc));
            bin = max(0, min(bin, NRL_ENTROPY_BINS - 1));
            acc += 1ull << (4 * bin);
          }
        }
        h0 += acc & lanes; h1 += (acc >> 4) & lanes; h2 += (acc >> 8) & lanes; h3 += (acc >> 12) & lanes;
      }
    }
    if (range != 0.0f) {
      const unsigned long long h[4] = {h0, h1, h2, h3};
#pragma unroll
      for (int b = 0; b < NRL_ENTROPY_BINS; ++b) {
        const int cnt = (int)((h[b & 3] >> ((b >> 2) * 16)) & 0xffffull);
        if (cnt > 0) {
          const float pr = nrl_div_approx((float)cnt, fn);
          r = fmaf(-pr, nrl_lg2_approx(pr), r);
        }
      }
    }
  }
  return r;
}

#define NRL_AGG_NT 128

// W: voxels per CTA (multiple of 4, <= 512).  bulk != 0 requires G % 4 == 0 and a 16-byte
// aligned series pointer (checked by the host).
__global__ void __launch_bounds__(NRL_AGG_NT) nrl_aggregate_kernel(const float *__restrict__ series, const int n,
                                                                   const long long G, const int agg, const int W,
                                                                   const int bulk, float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char nrl_agg_smem[];
  __shared__ __align__(8) uint64_t s_bar;
  float4 *buf = reinterpret_cast<float4 *>(nrl_agg_smem);
  const int tid = threadIdx.x;
  const long long g0 = (long long)blockIdx.x * W;
  const int w = (int)min((long long)W, G - g0);
  const int W4 = W >> 2;
  if (bulk) {
    if (tid == 0) nrl_mbar_init(&s_bar, 1);
    __syncthreads();
    if (tid == 0) nrl_mbar_expect_tx(&s_bar, (uint32_t)n * (uint32_t)w * 4u);
    if (tid < 32)
      for (int i = tid; i < n; i += 32)
        nrl_bulk_g2s(buf + (size_t)i * W4, series + (long long)i * G + g0, (uint32_t)w * 4u, &s_bar);
    nrl_mbar_wait(&s_bar, 0);
  } else {
    float *fb = reinterpret_cast<float *>(buf);
    for (int i = 0; i < n; ++i)
      for (int v = tid; v < W; v += NRL_AGG_NT) fb[(size_t)i * W + v] = v < w ? series[(long long)i * G + g0 + v] : 0.0f;
    __syncthreads();
  }
  if (agg == 2 || agg == 4 || agg == 7 || agg == NRL_AGG_LEGACY_ENTROPY || (agg == 3 && W <= NRL_AGG_NT)) {
    // one voxel per thread keeps all 128 threads busy (a float4 column per thread leaves three warps of a narrow
    // tile idle); the median does so only for narrow tiles, i.e. long series
    float *fb = reinterpret_cast<float *>(buf);
    for (int v = tid; v < w; v += NRL_AGG_NT) out[g0 + v] = nrl_agg_voxel(fb, W, v, n, agg);
    return;
  }
  for (int t4 = tid; t4 * 4 < w; t4 += NRL_AGG_NT) {
    const float4 r = nrl_reduce4(buf + t4, W4, n, agg);
    float *o = out + g0 + 4 * t4;
    if (4 * t4 + 3 < w && bulk) {
      *reinterpret_cast<float4 *>(o) = r;
    } else {
      if (4 * t4 + 0 < w) o[0] = r.x;
      if (4 * t4 + 1 < w) o[1] = r.y;
      if (4 * t4 + 2 < w) o[2] = r.z;
      if (4 * t4 + 3 < w) o[3] = r.w;
    }
  }
}

// Pipelined form for the two-pass reducers over long series (standard deviation, variance, median, entropy):
// a persistent CTA walks tiles of W voxels through a ring of `stages` shared-memory buffers; while the threads
// reduce tile k (one voxel per thread) the copies of tiles k+1 .. k+stages-1 are in flight, so the HBM stream
// never waits for a reduction.  A tile is ONE 2-D tiled TMA copy ([rows, W] box of the [n, G] series through a
// tensor map; SASS UTMALDG) -- measured on B200: an SM retires one 1-D bulk copy per ~46 clocks whatever its
// size, so F row copies of W*4 bytes cap a tile at W*4/46 B/clk per SM (512-byte rows: 66 % of the HBM peak).
// Needs G % 4 == 0 and a 16-byte aligned series (the host checks).  Out-of-range voxels / rows of the last
// box are zero-filled by the copy and never read.
#define NRL_AGG_PIPE_NT 256
#define NRL_AGG_PIPE_MAX_STAGES 8

__device__ __forceinline__ void nrl_tma_2d_g2s(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   nrl_smem_u32(dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(nrl_smem_u32(bar))
               : "memory");
}

// rows_per_box * n_boxes >= n rows are staged per tile (n_boxes = 1 up to 256 frames)
__global__ void __launch_bounds__(NRL_AGG_PIPE_NT, 1) nrl_aggregate_pipe_kernel(const __grid_constant__ CUtensorMap map, const int n,
                                                                                const long long G, const int agg, const int W,
                                                                                const int stages, const long long n_tiles,
                                                                                const int rows_per_box, const int n_boxes,
                                                                                float *__restrict__ out) {
  extern __shared__ __align__(1024) unsigned char nrl_agg_ring[];
  __shared__ __align__(8) uint64_t s_bar[NRL_AGG_PIPE_MAX_STAGES];
  // the destination of a tiled copy must be 128-byte aligned: round the dynamic window up (the host adds the slack)
  float *ring = reinterpret_cast<float *>(nrl_agg_ring + ((128u - (nrl_smem_u32(nrl_agg_ring) & 127u)) & 127u));
  const int tid = threadIdx.x;
  const size_t tile_floats = (size_t)rows_per_box * n_boxes * W;
  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
    for (int s = 0; s < stages; ++s) nrl_mbar_init(&s_bar[s], 1);
  }
  __syncthreads();
  // tile j of this CTA is global tile blockIdx.x + j * gridDim.x and lives in stage j % stages
  auto issue = [&](const long long j) {
    const long long tile = (long long)blockIdx.x + j * (long long)gridDim.x;
    if (tile >= n_tiles || tid != 0) return;
    const int s = (int)(j % stages);
    float *dst = ring + (size_t)s * tile_floats;
    nrl_mbar_expect_tx(&s_bar[s], (uint32_t)(tile_floats * sizeof(float)));
    for (int b = 0; b < n_boxes; ++b)
      nrl_tma_2d_g2s(dst + (size_t)b * rows_per_box * W, &map, (int)(tile * W), b * rows_per_box, &s_bar[s]);
  };
  for (int j = 0; j < stages - 1; ++j) issue(j);
  for (long long j = 0;; ++j) {
    const long long tile = (long long)blockIdx.x + j * (long long)gridDim.x;
    if (tile >= n_tiles) break;
    issue(j + stages - 1);          // its stage was drained in iteration j - 1 (barrier below)
    const int s = (int)(j % stages);
    nrl_mbar_wait(&s_bar[s], (uint32_t)((j / stages) & 1));
    const long long g0 = tile * W;
    const int w = (int)min((long long)W, G - g0);
    float *fb = ring + (size_t)s * tile_floats;
    for (int v = tid; v < w; v += NRL_AGG_PIPE_NT) out[g0 + v] = nrl_agg_voxel(fb, W, v, n, agg);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the median wrote the tile; the next copy overwrites it
    __syncthreads();
  }
}

// Streaming form for the reducers that are recurrences over the frames (mean, max, min, drift): every thread
// owns VEC consecutive voxels and walks the frames with independent 16-byte loads (eight in flight), nothing is
// staged.  The arithmetic is the one of nrl_reduce4 (reference order: sum_device / max_device / min_device /
// slope_device, /root/reference/src/gpuutils.cuh:11-50,190-205).
#define NRL_AGG_STREAM_NT 256

template <int VEC>
__global__ void __launch_bounds__(NRL_AGG_STREAM_NT) nrl_aggregate_stream_kernel(const float *__restrict__ series, const int n,
                                                                                  const long long G, const int agg,
                                                                                  float *__restrict__ out) {
  const long long g = ((long long)blockIdx.x * NRL_AGG_STREAM_NT + threadIdx.x) * VEC;
  if (g >= G) return;
  const float fn = (float)n;
  float acc[VEC], sxy[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) { acc[q] = 0.0f; sxy[q] = 0.0f; }
  float sx = 0.0f, sx2 = 0.0f;
  const float *col = series + g;
  for (int i0 = 0; i0 < n; i0 += 8) {
    float x[8][VEC];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (i0 + u < n) {
        if constexpr (VEC == 4) {
          const float4 v = __ldcs(reinterpret_cast<const float4 *>(col + (long long)(i0 + u) * G));
          x[u][0] = v.x; x[u][1] = v.y; x[u][2] = v.z; x[u][3] = v.w;
        } else {
          x[u][0] = __ldcs(col + (long long)(i0 + u) * G);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u;
      if (i >= n) break;
#pragma unroll
      for (int q = 0; q < VEC; ++q) {
        if (agg == 1) acc[q] = acc[q] + x[u][q];
        else if (agg == 5) { if (i == 0 || x[u][q] > acc[q]) acc[q] = x[u][q]; }
        else if (agg == 6) { if (i == 0 || x[u][q] < acc[q]) acc[q] = x[u][q]; }
        else { acc[q] = acc[q] + x[u][q]; sxy[q] = fmaf(x[u][q], (float)i, sxy[q]); }
      }
      if (agg == 8) { sx = sx + (float)i; sx2 = sx2 + (float)(i * i); }
    }
  }
  float r[VEC];
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    if (agg == 1) r[q] = nrl_div_approx(acc[q], fn);
    else if (agg == 8) {
      const float denom = fmaf(fn, sx2, -(sx * sx));
      r[q] = (n > 1 && denom != 0.0f) ? nrl_div_approx(fmaf(fn, sxy[q], -(acc[q] * sx)), denom) : 0.0f;
    } else r[q] = acc[q];
  }
  if constexpr (VEC == 4) *reinterpret_cast<float4 *>(out + g) = make_float4(r[0], r[1], r[2], r[3]);
  else out[g] = r[0];
}

// Register form for the two-pass reducers (standard deviation, variance, entropy) over short series (n <= NMAX):
// a thread owns ONE voxel and holds its whole series in registers -- all n loads are in flight at once, the
// second pass never leaves the register file.  Arithmetic as in nrl_reduce4 (gpuutils.cuh:35-74,140-173).
template <int NMAX>
__global__ void __launch_bounds__(256) nrl_aggregate_reg_kernel(const float *__restrict__ series, const int n, const long long G,
                                                                const int agg, float *__restrict__ out) {
  const long long g = (long long)blockIdx.x * 256 + threadIdx.x;
  if (g >= G) return;
  float x[NMAX];
  const float *col = series + g;
#pragma unroll
  for (int i = 0; i < NMAX; ++i) x[i] = i < n ? __ldcs(col + (long long)i * G) : 0.0f;
  const float fn = (float)n;
  float r = 0.0f;
  if (agg == 2 || agg == 4) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < NMAX; ++i) if (i < n) s = s + x[i];
    const float mean = nrl_div_approx(s, fn);
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < NMAX; ++i) if (i < n) { const float d = x[i] - mean; acc = fmaf(d, d, acc); }
    r = agg == 2 ? nrl_sqrt_approx(nrl_div_approx(acc, fn)) : nrl_div_approx(acc, fn);
  } else if (n > 1) {    // information entropy: 16-bin histogram between the series' extremes
    float mn = x[0], mx = x[0];
#pragma unroll
    for (int i = 1; i < NMAX; ++i) if (i < n) { if (x[i] < mn) mn = x[i]; if (x[i] > mx) mx = x[i]; }
    const float range = mx - mn;
    unsigned long long h0 = 0ull, h1 = 0ull, h2 = 0ull, h3 = 0ull;      // 16 counters of 16 bits
#pragma unroll
    for (int i = 0; i < NMAX; ++i)
      if (i < n) {
        const float sc = nrl_div_approx(x[i] - mn, range) * 16.0f;
        int bin;
        asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(bin) : "f"(sc));
        bin = max(0, min(bin, NRL_ENTROPY_BINS - 1));
        const unsigned long long inc = 1ull << ((bin & 3) * 16);
        const int word = bin >> 2;
        h0 += (word == 0) ? inc : 0ull; h1 += (word == 1) ? inc : 0ull;
        h2 += (word == 2) ? inc : 0ull; h3 += (word == 3) ? inc : 0ull;
      }
    if (range != 0.0f) {
      const unsigned long long h[4] = {h0, h1, h2, h3};
#pragma unroll
      for (int b = 0; b < NRL_ENTROPY_BINS; ++b) {
        const int cnt = (int)((h[b >> 2] >> ((b & 3) * 16)) & 0xffffull);
        if (cnt > 0) {
          const float pr = nrl_div_approx((float)cnt, fn);
          r = fmaf(-pr, nrl_lg2_approx(pr), r);
        }
      }
    }
  }
  out[g] = r;
}

// summation: the reference's 256-wide shared-memory tree per block, then a sequential sum of
// the block partials (/root/reference/src/gpuutils.cu:9-31,107-134) -- kept on the device.
__global__ void __launch_bounds__(256) nrl_block_sum_kernel(const float *__restrict__ in, float *__restrict__ partial,
                                                            const long long n) {
  __shared__ float sm[256];
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  sm[threadIdx.x] = idx < n ? in[idx] : 0.0f;
  __syncthreads();
  for (int stride = 128; stride > 0; stride >>= 1) {
    if ((int)threadIdx.x < stride) sm[threadIdx.x] = sm[threadIdx.x] + sm[threadIdx.x + stride];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sm[0];
}
__global__ void nrl_final_sum_kernel(const float *__restrict__ partial, const int nblocks, float *__restrict__ out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float s = 0.0f;
    for (int i = 0; i < nblocks; ++i) s = s + partial[i];
    out[0] = s;
  }
}

__global__ void nrl_buff_dim_kernel(const float cutoff, const float spacing, int *out) {
  if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = nrl_buff_dim_device(cutoff, spacing);
}
