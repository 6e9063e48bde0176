/*
 * nearl_oracle.c -- CPU restatement of Nearl's featurisation hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py may load
 * it.  The product path (nearl_b200/) never calls into this file.
 *
 * Every function restates, in plain C, what the reference computes on the GPU, in the
 * reference's own operation order.  Citations are relative to /root/reference/.
 *
 *   A1  orc_frame_voxelize      src/voxelize.cu:38-130  (frame_interp_global)
 *   A2  orc_density_flow        src/voxelize.cu:350-415 (trajectory_voxelization_host)
 *   A3  orc_observe_frame       src/marching_observers.cu:30-429
 *       orc_marching_observer   src/marching_observers.cu:435-516
 *   A4  orc_aggregate           src/gpuutils.cu:48-77 + src/gpuutils.cuh:11-245
 *       orc_summation           src/gpuutils.cu:9-31,107-134
 *
 * Parity pins (see tests/test_oracle_pins.py): the reference's own known-answer tests
 *   pytests/test_actions.py:18-31  (mass conservation),
 *   pytests/test_benchmark.py:19-29,
 *   pytests/test_features.py:86-126 (eight reducers vs numpy),
 * and, on the GPU box, the reference's kernels themselves compiled from
 * /root/reference/src into oracle/_ref/ (oracle/build_ref.sh).
 *
 * Documented deviations from the reference *binary* (all below the 1e-5 relative
 * tolerance, none changes a discrete decision except where stated):
 *   - the reference is built with -use_fast_math: div.approx / sqrt.approx / lg2.approx
 *     and flush-to-zero.  Here IEEE division, sqrtf and log2f are used.
 *   - buff_dim = (int)((cutoff+spacing)/spacing) is evaluated by the reference with
 *     div.approx on the device (src/voxelize.cu:50).  Here it is IEEE; callers that
 *     know the device value pass it in (buff_dim >= 0) so that the oracle follows the
 *     hardware.
 *   - the reference deposits with float atomicAdd in a nondeterministic order
 *     (src/voxelize.cu:127).  Here per-voxel sums are accumulated in double in atom
 *     order and rounded once: the limit every legal order approximates.
 *   - entropy bins ((int)(normalized*16)) can flip by one bin on 1-ulp differences of
 *     the approximate division; the reference's own test tolerates <10 / 32768 voxels.
 *
 * The squared-distance expressions reproduce the FMA contraction nvcc emits for the
 * reference (read from the PTX of the reference built with its own flags):
 *   voxeliser : d2 = fma(dz,dz, fma(dx,dx, dy*dy))   with d = atom - gridpoint
 *   observers : d2 = fma(dz,dz, fma(dy,dy, dx*dx))   with d = observer - atom
 * so that in/out-of-cutoff decisions are bit-identical.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define ORC_BLOCK 256                     /* src/constants.h:8  BLOCK_SIZE */
#define ORC_PLACEHOLDER 99999.0f          /* src/constants.h:9-10 */
#define ORC_MAX_FRAMES 512                /* src/constants.h:11 */
#define ORC_DISTINCT_LIMIT 1000           /* src/constants.h:12 */
#define ORC_ENTROPY_BINS 16               /* src/constants.h:34 */

#ifdef __cplusplus
extern "C" {
#endif

/* src/voxelize.cu:50 -- IEEE evaluation of the expression */
int orc_buff_dim(float cutoff, float spacing) {
  volatile float num = cutoff + spacing;
  volatile float q = num / spacing;
  return (int)q;
}

/* src/gpuutils.cuh:310-317 gaussian_map_device<float>(distance, 0, sigma):
 * float argument, double exp, divided by the FLOAT product sigma*sqrtf(2*pi),
 * result rounded to float by the return type. */
static inline float orc_gaussian(float distance, float sigma) {
  if (sigma == 0.0f) return 0.0f;
  volatile float q = distance / sigma;
  double qd = (double)q;
  double e = exp((-0.5 * qd) * qd);
  volatile float twopi = (float)(2 * M_PI);
  volatile float denom = sigma * sqrtf(twopi);
  return (float)(e / (double)denom);
}

/* voxeliser squared distance, see header comment */
static inline float orc_d2_vox(float cx, float cy, float cz, float gx, float gy, float gz) {
  float dx = cx - gx, dy = cy - gy, dz = cz - gz;
  volatile float t = dy * dy;
  float u = fmaf(dx, dx, t);
  return fmaf(dz, dz, u);
}

/* observer squared distance (src/gpuutils.cuh:338-345), see header comment */
static inline float orc_d2_obs(float ox, float oy, float oz, float ax, float ay, float az) {
  float dx = ox - ax, dy = oy - ay, dz = oz - az;
  volatile float t = dx * dx;
  float u = fmaf(dy, dy, t);
  return fmaf(dz, dz, u);
}

static inline int orc_imax(int a, int b) { return a > b ? a : b; }
static inline int orc_imin(int a, int b) { return a < b ? a : b; }

/*
 * A1.  The reference decodes a flat index with dims[0] twice and dims[1] once
 * (src/voxelize.cu:80-82,113-115):  x = gid/d0/d1, y = gid/d0%d1, z = gid%d0.
 * That is a lattice of extents (NX,NY,NZ) = (d2,d1,d0) stored as (x*NY+y)*NZ+z; for
 * cubic grids it is the ordinary C layout.  The same extents, each widened by
 * 2*buff_dim, give the buffered lattice of the normalisation pass.
 */
void orc_frame_voxelize(const float *coords, const float *weights, int n, const int *dims,
                        float spacing, float cutoff, float sigma, int buff_dim, float *out) {
  const int NX = dims[2], NY = dims[1], NZ = dims[0];
  const long G = (long)NX * NY * NZ;
  const int b = buff_dim >= 0 ? buff_dim : orc_buff_dim(cutoff, spacing);
  const int BX = NX + 2 * b, BY = NY + 2 * b, BZ = NZ + 2 * b;
  const float c2 = cutoff * cutoff;
  double *acc = (double *)calloc((size_t)(G > 0 ? G : 1), sizeof(double));
  const int reach = (int)ceilf(cutoff / spacing) + 2;

  for (int a = 0; a < n; ++a) {
    const float cx = coords[3 * a], cy = coords[3 * a + 1], cz = coords[3 * a + 2];
    const float w = weights[a];
    if (cx == ORC_PLACEHOLDER && cy == ORC_PLACEHOLDER && cz == ORC_PLACEHOLDER) continue; /* :58-62 */
    if (w == 0.0f) continue;                                                               /* :63 */
    if (!(cx == cx) || !(cy == cy) || !(cz == cz)) continue; /* NaN never passes d2 < c2 */
    if (fabsf(cx) > 1e9f || fabsf(cy) > 1e9f || fabsf(cz) > 1e9f) continue;

    /* pass 1 (:79-108): 256 strided partial sums, then the shared-memory tree */
    float part[ORC_BLOCK];
    for (int t = 0; t < ORC_BLOCK; ++t) part[t] = 0.0f;
    const int ax = (int)floorf(cx / spacing), ay = (int)floorf(cy / spacing), az = (int)floorf(cz / spacing);
    const int x0 = orc_imax(ax - reach, -b), x1 = orc_imin(ax + reach, NX + b - 1);
    const int y0 = orc_imax(ay - reach, -b), y1 = orc_imin(ay + reach, NY + b - 1);
    const int z0 = orc_imax(az - reach, -b), z1 = orc_imin(az + reach, NZ + b - 1);
    for (int x = x0; x <= x1; ++x)
      for (int y = y0; y <= y1; ++y)
        for (int z = z0; z <= z1; ++z) {
          const float gx = (float)x * spacing, gy = (float)y * spacing, gz = (float)z * spacing;
          const float d2 = orc_d2_vox(cx, cy, cz, gx, gy, gz);
          if (d2 < c2) {                                                     /* :95 strict */
            const long gid = ((long)(x + b) * BY + (y + b)) * BZ + (z + b);
            const int t = (int)(gid % ORC_BLOCK);
            volatile float s = part[t] + orc_gaussian(sqrtf(d2), sigma);
            part[t] = s;
          }
        }
    (void)BX;
    for (int stride = ORC_BLOCK / 2; stride > 0; stride >>= 1)               /* :102-107 */
      for (int t = 0; t < stride; ++t) {
        volatile float s = part[t] + part[t + stride];
        part[t] = s;
      }
    const float total = part[0];
    if (total == 0.0f) continue;                                             /* :109 */
    volatile float inv = w / total;                                          /* :111 */

    /* pass 2 (:112-129): deposit on in-box lattice points */
    const int ix0 = orc_imax(x0, 0), ix1 = orc_imin(x1, NX - 1);
    const int iy0 = orc_imax(y0, 0), iy1 = orc_imin(y1, NY - 1);
    const int iz0 = orc_imax(z0, 0), iz1 = orc_imin(z1, NZ - 1);
    for (int x = ix0; x <= ix1; ++x)
      for (int y = iy0; y <= iy1; ++y)
        for (int z = iz0; z <= iz1; ++z) {
          const float gx = (float)x * spacing, gy = (float)y * spacing, gz = (float)z * spacing;
          const float d2 = orc_d2_vox(cx, cy, cz, gx, gy, gz);
          if (d2 < c2) {                                                     /* :125 strict */
            volatile float dep = orc_gaussian(sqrtf(d2), sigma) * inv;       /* :127 */
            acc[((long)x * NY + y) * NZ + z] += (double)dep;
          }
        }
  }
  for (long g = 0; g < G; ++g) out[g] = (float)acc[g];
  free(acc);
}

/* ------------------------------------------------------------------ A3 observers */

typedef struct { const float *xyz; const float *w; int n; float c; float c2; } orc_frame_t;

/* :30-47 */
static float orc_existence(const float *o, const orc_frame_t *f) {
  for (int j = 0; j < f->n; ++j) {
    float d2 = orc_d2_obs(o[0], o[1], o[2], f->xyz[3 * j], f->xyz[3 * j + 1], f->xyz[3 * j + 2]);
    if (d2 > f->c2) continue;
    return 1.0f;
  }
  return 0.0f;
}
/* :67-84 */
static float orc_direct_count(const float *o, const orc_frame_t *f) {
  float r = 0.0f;
  for (int j = 0; j < f->n; ++j) {
    float d2 = orc_d2_obs(o[0], o[1], o[2], f->xyz[3 * j], f->xyz[3 * j + 1], f->xyz[3 * j + 2]);
    if (d2 > f->c2) continue;
    r += 1.0f;
  }
  return r;
}
/* :110-142 */
static float orc_distinct_count(const float *o, const orc_frame_t *f) {
  float seen[ORC_DISTINCT_LIMIT];
  for (int i = 0; i < ORC_DISTINCT_LIMIT; ++i) seen[i] = ORC_PLACEHOLDER;
  float cnt = 0.0f;
  for (int j = 0; j < f->n; ++j) {
    float d2 = orc_d2_obs(o[0], o[1], o[2], f->xyz[3 * j], f->xyz[3 * j + 1], f->xyz[3 * j + 2]);
    if (d2 > f->c2) continue;
    const float v = f->w[j];
    for (int k = 0; k < ORC_DISTINCT_LIMIT; ++k) {
      if (v == seen[k]) break;
      else if (seen[k] == ORC_PLACEHOLDER) { seen[k] = v; cnt += 1; break; }
    }
  }
  return cnt;
}
/* :151-174 */
static float orc_mean_distance(const float *o, const orc_frame_t *f) {
  volatile float r = 0.0f, ws = 0.0f;
  int count = 0;
  for (int j = 0; j < f->n; ++j) {
    float d2 = orc_d2_obs(o[0], o[1], o[2], f->xyz[3 * j], f->xyz[3 * j + 1], f->xyz[3 * j + 2]);
    if (d2 > f->c2) continue;
    r = fmaf(sqrtf(d2), f->w[j], r);
    ws = ws + f->w[j];
    count += 1;
  }
  if (count > 0) return r / ws;
  return 0.0f;
}
/* :180-191 */
static float orc_cumulative_weight(const float *o, const orc_frame_t *f) {
  volatile float r = 0.0f;
  for (int j = 0; j < f->n; ++j) {
    float d2 = orc_d2_obs(o[0], o[1], o[2], f->xyz[3 * j], f->xyz[3 * j + 1], f->xyz[3 * j + 2]);
    if (d2 > f->c2) continue;
    r = r + f->w[j];
  }
  return r;
}
/* :197-202  `float volume = (4.0/3.0)*M_PI*cutoff^3`: evaluated in double, stored as a float; float / float (the
 * reference binary divides with div.approx under -use_fast_math: <= 2 ulp from this IEEE quotient) */
static float orc_density(const float *o, const orc_frame_t *f) {
  float ws = orc_cumulative_weight(o, f);
  float vol = (float)((4.0 / 3.0) * M_PI * (double)f->c * (double)f->c * (double)f->c);
  return ws / vol;
}
/* :219-249 */
static float orc_dispersion(const float *o, const orc_frame_t *f) {
  volatile float ws = 0.0f, r = 0.0f;
  for (int j = 0; j < f->n; ++j) {
    const float *pj = f->xyz + 3 * j;
    float d2j = orc_d2_obs(o[0], o[1], o[2], pj[0], pj[1], pj[2]);
    if (d2j > f->c2) continue;
    for (int k = j + 1; k < f->n; ++k) {
      const float *pk = f->xyz + 3 * k;
      float d2k = orc_d2_obs(o[0], o[1], o[2], pk[0], pk[1], pk[2]);
      if (d2k > f->c2) continue;
      float d2 = orc_d2_obs(pj[0], pj[1], pj[2], pk[0], pk[1], pk[2]);
      volatile float ww = f->w[j] * f->w[k];
      r = fmaf(ww, sqrtf(d2), r);
      ws = ws + ww;
    }
  }
  if (ws == 0.0f) return 0.0f;
  return r / ws;
}
/* :270-299 */
static float orc_eccentricity(const float *o, const orc_frame_t *f) {
  volatile float cx = 0.0f, cy = 0.0f, cz = 0.0f, ws = 0.0f;
  int count = 0;
  for (int j = 0; j < f->n; ++j) {
    const float *p = f->xyz + 3 * j;
    float d2 = orc_d2_obs(o[0], o[1], o[2], p[0], p[1], p[2]);
    if (d2 > f->c2) continue;
    count += 1;
    cx = fmaf(f->w[j], p[0], cx);
    cy = fmaf(f->w[j], p[1], cy);
    cz = fmaf(f->w[j], p[2], cz);
    ws = ws + f->w[j];
  }
  if (count > 0 && ws != 0.0f) {
    /* the reference binary: one reciprocal, each quotient fused with the subtraction (FFMA rcp, -com, o) */
    float rcp = 1.0f / ws;
    float dx = fmaf(rcp, -cx, o[0]), dy = fmaf(rcp, -cy, o[1]), dz = fmaf(rcp, -cz, o[2]);
    return sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
  }
  return f->c;
}
/* :313-355 */
static float orc_radius_of_gyration(const float *o, const orc_frame_t *f) {
  volatile float cx = 0.0f, cy = 0.0f, cz = 0.0f, ws = 0.0f;
  int count = 0;
  for (int j = 0; j < f->n; ++j) {
    const float *p = f->xyz + 3 * j;
    float d2 = orc_d2_obs(o[0], o[1], o[2], p[0], p[1], p[2]);
    if (d2 > f->c2) continue;
    cx = fmaf(f->w[j], p[0], cx);
    cy = fmaf(f->w[j], p[1], cy);
    cz = fmaf(f->w[j], p[2], cz);
    ws = ws + f->w[j];
    count += 1;
  }
  if (count <= 1) return 0.0f;
  if (ws == 0.0f) return 0.0f;
  const float com0 = cx / ws, com1 = cy / ws, com2 = cz / ws;
  volatile float r = 0.0f;
  for (int j = 0; j < f->n; ++j) {
    const float *p = f->xyz + 3 * j;
    float d2 = orc_d2_obs(o[0], o[1], o[2], p[0], p[1], p[2]);
    if (d2 > f->c2) continue;
    d2 = orc_d2_obs(com0, com1, com2, p[0], p[1], p[2]);
    r = fmaf(f->w[j], d2, r);
  }
  return sqrtf(r / ws);
}

static float orc_observe_one(const float *o, const orc_frame_t *f, int type_obs) {
  switch (type_obs) {                                  /* :364-399 */
    case 1: return orc_existence(o, f);
    case 2: return orc_direct_count(o, f);
    case 3: return orc_distinct_count(o, f);
    case 11: return orc_mean_distance(o, f);
    case 12: return orc_cumulative_weight(o, f);
    case 13: return orc_density(o, f);
    case 14: return orc_dispersion(o, f);
    case 15: return orc_eccentricity(o, f);
    case 16: return orc_radius_of_gyration(o, f);
    default: return 0.0f;
  }
}

/* src/marching_observers.cu:405-429, one frame -> [G] */
void orc_observe_frame(const float *coords, const float *weights, int n, const int *dims,
                       float spacing, float cutoff, int type_obs, float *out) {
  const unsigned d0 = (unsigned)dims[0], d1 = (unsigned)dims[1], d2n = (unsigned)dims[2];
  const long G = (long)d0 * d1 * d2n;
  orc_frame_t fr = { coords, weights, n, cutoff, cutoff * cutoff };
#pragma omp parallel for schedule(dynamic, 64)
  for (long g = 0; g < G; ++g) {
    const unsigned index = (unsigned)g;
    float o[3];
    o[0] = (float)(index / (d0 * d1)) * spacing;           /* :416-420 */
    o[1] = (float)((index / d0) % d1) * spacing;
    o[2] = (float)(index % d0) * spacing;
    out[g] = orc_observe_one(o, &fr, type_obs);
  }
}

/* ------------------------------------------------------------------ A4 reducers */

static float orc_reduce(float *a, int n, int type_agg) {
  switch (type_agg) {
    case 1: {                                              /* gpuutils.cuh:35-50 */
      volatile float s = 0.0f;
      for (int i = 0; i < n; ++i) s = s + a[i];
      return s / (float)n;
    }
    case 2: case 4: {                                      /* :53-74 */
      volatile float s = 0.0f;
      for (int i = 0; i < n; ++i) s = s + a[i];
      const float mean = s / (float)n;
      volatile float q = 0.0f;
      for (int i = 0; i < n; ++i) { float d = a[i] - mean; q = fmaf(d, d, q); }
      if (type_agg == 2) return sqrtf(q / (float)n);
      return q / (float)n;
    }
    case 3: {                                              /* :227-245 exchange sort */
      for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
          if (a[i] > a[j]) { float t = a[i]; a[i] = a[j]; a[j] = t; }
      if (n % 2 == 0) return (a[n / 2 - 1] + a[n / 2]) / 2;
      return a[n / 2];
    }
    case 5: { float m = a[0]; for (int i = 1; i < n; ++i) if (a[i] > m) m = a[i]; return m; }
    case 6: { float m = a[0]; for (int i = 1; i < n; ++i) if (a[i] < m) m = a[i]; return m; }
    case 7: {                                              /* :140-173 */
      if (n <= 1) return 0.0f;
      float mn = a[0], mx = a[0];
      for (int i = 1; i < n; ++i) { if (a[i] < mn) mn = a[i]; if (a[i] > mx) mx = a[i]; }
      const float range = mx - mn;
      if (range == 0) return 0.0f;
      int hist[ORC_ENTROPY_BINS];
      memset(hist, 0, sizeof hist);
      for (int i = 0; i < n; ++i) {
        volatile float nrm = (a[i] - mn) / range;
        volatile float sc = nrm * (float)ORC_ENTROPY_BINS;
        int bin = (int)sc;
        if (bin >= ORC_ENTROPY_BINS) bin = ORC_ENTROPY_BINS - 1;
        if (bin < 0) bin = 0;  /* NaN guard only; unreachable for finite series */
        hist[bin] += 1;
      }
      volatile float e = 0.0f;
      for (int i = 0; i < ORC_ENTROPY_BINS; ++i)
        if (hist[i] > 0) {
          volatile float p = (float)hist[i] / (float)n;
          e = fmaf(-p, log2f(p), e);          /* FFMA in the reference SASS */
        }
      return e;
    }
    case 8: {                                              /* :190-205, op order from the SASS */
      if (n <= 1) return 0.0f;
      volatile float sx = 0, sy = 0, sxy = 0, sx2 = 0;
      for (int i = 0; i < n; ++i) {
        sx = sx + (float)i;
        sy = sy + a[i];
        sxy = fmaf(a[i], (float)i, sxy);
        sx2 = sx2 + (float)(i * i);
      }
      volatile float t2 = sx * sx;
      const float denom = fmaf((float)n, sx2, -t2);   /* FMUL + FFMA in the reference SASS */
      if (denom == 0) return 0.0f;
      volatile float u2 = sy * sx;
      return fmaf((float)n, sxy, -u2) / denom;
    }
    default: return 0.0f;                                  /* gpuutils.cu:73-76 */
  }
}

/* src/gpuutils.cu:48-77,80-105: series is frame-major [F,G]; frames beyond 512 ignored */
void orc_aggregate(const float *series, int F, long G, int type_agg, float *out) {
  const int n = F > ORC_MAX_FRAMES ? ORC_MAX_FRAMES : F;
#pragma omp parallel for schedule(static)
  for (long g = 0; g < G; ++g) {
    float tmp[ORC_MAX_FRAMES];
    for (int i = 0; i < n; ++i) tmp[i] = series[(long)i * G + g];
    out[g] = n > 0 ? orc_reduce(tmp, n, type_agg) : 0.0f;
  }
}

/* src/marching_observers.cu:435-516 (the trailing sum-reduction is dead code, :496-504) */
void orc_marching_observer(const float *coords, const float *weights, int F, int N, const int *dims,
                           float spacing, float cutoff, int type_obs, int type_agg, float *out) {
  const long G = (long)dims[0] * dims[1] * dims[2];
  float *traj = (float *)malloc(sizeof(float) * (size_t)(F > 0 ? F : 1) * (size_t)G);
  for (int f = 0; f < F; ++f)
    orc_observe_frame(coords + (long)f * N * 3, weights + (long)f * N, N, dims, spacing, cutoff, type_obs,
                      traj + (long)f * G);
  orc_aggregate(traj, F, G, type_agg, out);
  free(traj);
}

/* writes the [F,G] per-frame observations too (used to pin integer modes bit-exactly) */
void orc_marching_observer_frames(const float *coords, const float *weights, int F, int N, const int *dims,
                                  float spacing, float cutoff, int type_obs, float *frames_out) {
  const long G = (long)dims[0] * dims[1] * dims[2];
  for (int f = 0; f < F; ++f)
    orc_observe_frame(coords + (long)f * N * 3, weights + (long)f * N, N, dims, spacing, cutoff, type_obs,
                      frames_out + (long)f * G);
}

/* src/voxelize.cu:350-415 */
void orc_density_flow(const float *coords, const float *weights, int F, int N, const int *dims,
                      float spacing, float cutoff, float sigma, int type_agg, int buff_dim, float *out) {
  const long G = (long)dims[0] * dims[1] * dims[2];
  float *traj = (float *)malloc(sizeof(float) * (size_t)(F > 0 ? F : 1) * (size_t)G);
#pragma omp parallel for schedule(dynamic, 1)
  for (int f = 0; f < F; ++f)
    orc_frame_voxelize(coords + (long)f * N * 3, weights + (long)f * N, N, dims, spacing, cutoff, sigma,
                       buff_dim, traj + (long)f * G);
  orc_aggregate(traj, F, G, type_agg, out);
  free(traj);
}

/* src/gpuutils.cu:9-31,107-134: 256-wide block tree sums, then a sequential host sum */
float orc_summation(const float *arr, long n) {
  volatile float total = 0.0f;
  for (long b0 = 0; b0 < n; b0 += ORC_BLOCK) {
    float sm[ORC_BLOCK];
    for (int t = 0; t < ORC_BLOCK; ++t) sm[t] = (b0 + t < n) ? arr[b0 + t] : 0.0f;
    for (int stride = ORC_BLOCK / 2; stride > 0; stride >>= 1)
      for (int t = 0; t < stride; ++t) { volatile float s = sm[t] + sm[t + stride]; sm[t] = s; }
    total = total + sm[0];
  }
  return total;
}

#ifdef __cplusplus
}
#endif
