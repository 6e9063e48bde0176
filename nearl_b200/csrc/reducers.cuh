// nearl_b200 -- K3: the eight per-voxel time-series reducers, four voxels at a time.
//
// Restates /root/reference/src/gpuutils.cuh:11-245 as dispatched by
// gridwise_aggregation_global (/root/reference/src/gpuutils.cu:48-77), with the operation
// order and the approximate division / square root / log2 of the reference binary
// (PTX of the reference built with -use_fast_math).  The series of four neighbouring
// voxels is a strided column of float4 (`stride` float4 between consecutive frames) in
// shared memory (fused use) or global scratch.
#pragma once
#include "common.cuh"

#define NRL4(...) { { const int q = 0; __VA_ARGS__; } { const int q = 1; __VA_ARGS__; } { const int q = 2; __VA_ARGS__; } { const int q = 3; __VA_ARGS__; } }

__device__ __forceinline__ void nrl_unpack4(const float4 v, float (&x)[4]) { x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w; }

__device__ __noinline__ float4 nrl_reduce4(float4 *col, const int stride, const int n, const int agg) {
  float out[4] = {0.f, 0.f, 0.f, 0.f};
  const float fn = (float)n;
  if (agg == 1 || agg == 2 || agg == 4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int i = 0; i < n; ++i) {                      // sum_device, gpuutils.cuh:35-42
      float x[4]; nrl_unpack4(col[(size_t)i * stride], x);
      NRL4(s[q] = s[q] + x[q]);
    }
    float mean[4];
    NRL4(mean[q] = nrl_div_approx(s[q], fn));          // mean_device :45-50
    if (agg == 1) {
      NRL4(out[q] = mean[q]);
    } else {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int i = 0; i < n; ++i) {                    // :53-74
        float x[4]; nrl_unpack4(col[(size_t)i * stride], x);
        NRL4({ const float d = x[q] - mean[q]; acc[q] = fmaf(d, d, acc[q]); });
      }
      if (agg == 2) { NRL4(out[q] = nrl_sqrt_approx(nrl_div_approx(acc[q], fn))); }
      else          { NRL4(out[q] = nrl_div_approx(acc[q], fn)); }
    }
  } else if (agg == 3) {
    // median_device :227-245 -- the reference's in-place exchange sort, a[i] kept in registers
    for (int i = 0; i < n; ++i) {
      float a[4]; nrl_unpack4(col[(size_t)i * stride], a);
      for (int j = i + 1; j < n; ++j) {
        float b[4]; nrl_unpack4(col[(size_t)j * stride], b);
        bool any = false;
        NRL4(if (a[q] > b[q]) { const float tmp = a[q]; a[q] = b[q]; b[q] = tmp; any = true; });
        if (any) col[(size_t)j * stride] = make_float4(b[0], b[1], b[2], b[3]);
      }
      col[(size_t)i * stride] = make_float4(a[0], a[1], a[2], a[3]);
    }
    if ((n & 1) == 0) {
      float lo[4], hi[4];
      nrl_unpack4(col[(size_t)(n / 2 - 1) * stride], lo);
      nrl_unpack4(col[(size_t)(n / 2) * stride], hi);
      NRL4(out[q] = nrl_div_approx(lo[q] + hi[q], 2.0f));
    } else {
      nrl_unpack4(col[(size_t)(n / 2) * stride], out);
    }
  } else if (agg == 5 || agg == 6) {
    float m[4]; nrl_unpack4(col[0], m);
#pragma unroll 8
    for (int i = 1; i < n; ++i) {                      // max_device / min_device :11-32
      float x[4]; nrl_unpack4(col[(size_t)i * stride], x);
      if (agg == 5) { NRL4(if (x[q] > m[q]) m[q] = x[q]); }
      else          { NRL4(if (x[q] < m[q]) m[q] = x[q]); }
    }
    NRL4(out[q] = m[q]);
  } else if (agg == 7) {
    // information_entropy_histogram_device :140-173; 16 bin counters of 16 bits packed in 4 x u64
    if (n > 1) {
      float mn[4], mx[4];
      nrl_unpack4(col[0], mn); nrl_unpack4(col[0], mx);
      for (int i = 1; i < n; ++i) {
        float x[4]; nrl_unpack4(col[(size_t)i * stride], x);
        NRL4({ if (x[q] < mn[q]) mn[q] = x[q]; if (x[q] > mx[q]) mx[q] = x[q]; });
      }
      float range[4];
      NRL4(range[q] = mx[q] - mn[q]);
      unsigned long long h[4][4];
      NRL4({ h[q][0] = 0ull; h[q][1] = 0ull; h[q][2] = 0ull; h[q][3] = 0ull; });
      for (int i = 0; i < n; ++i) {
        float x[4]; nrl_unpack4(col[(size_t)i * stride], x);
        NRL4({
          const float nrm = nrl_div_approx(x[q] - mn[q], range[q]);
          int bin;
          const float sc = nrm * 16.0f;
          asm("cvt.rzi.ftz.s32.f32 %0, %1;" : "=r"(bin) : "f"(sc));
          bin = max(0, min(bin, NRL_ENTROPY_BINS - 1));
          const unsigned long long inc = 1ull << ((bin & 3) * 16);
          const int word = bin >> 2;
          h[q][0] += (word == 0) ? inc : 0ull;
          h[q][1] += (word == 1) ? inc : 0ull;
          h[q][2] += (word == 2) ? inc : 0ull;
          h[q][3] += (word == 3) ? inc : 0ull;
        });
      }
      NRL4({
        float e = 0.0f;
        if (range[q] != 0.0f) {
          for (int b = 0; b < NRL_ENTROPY_BINS; ++b) {
            const int cnt = (int)((h[q][b >> 2] >> ((b & 3) * 16)) & 0xffffull);
            if (cnt > 0) {
              const float pr = nrl_div_approx((float)cnt, fn);
              e = fmaf(-pr, nrl_lg2_approx(pr), e);      // ptxas fuses the reference's mul+sub into FFMA
            }
          }
        }
        out[q] = e;
      });
    }
  } else if (agg == 8) {
    // slope_device :190-205; op order of the reference SASS: one product, one fused
    // multiply-subtract (ptxas contracts the unrounded mul/sub pair), then rcp * numerator
    if (n > 1) {
      float sx = 0.0f, sx2 = 0.0f;
      float sy[4] = {0.f, 0.f, 0.f, 0.f}, sxy[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < n; ++i) {
        float x[4]; nrl_unpack4(col[(size_t)i * stride], x);
        const float fi = (float)i;
        sx = sx + fi;
        NRL4({ sy[q] = sy[q] + x[q]; sxy[q] = fmaf(x[q], fi, sxy[q]); });
        sx2 = sx2 + (float)(i * i);
      }
      const float denom = fmaf(fn, sx2, -(sx * sx));
      if (denom != 0.0f) { NRL4(out[q] = nrl_div_approx(fmaf(fn, sxy[q], -(sy[q] * sx)), denom)); }
    }
  }
  return make_float4(out[0], out[1], out[2], out[3]);
}
