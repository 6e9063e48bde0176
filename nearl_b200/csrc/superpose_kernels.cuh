// nearl_b200 -- rigid-body superposition of trajectory frames on the device.
//
// The step in front of the hot path: MisatoTraj(..., superpose=True) fits every frame onto the
// first one over the C-alpha atoms (pt.superpose(traj, mask="@CA"),
// /root/reference/nearl/io/traj.py:398-404; the arithmetic itself lives in cpptraj, which is not
// part of the reference tree).  Restated as the textbook least-squares fit: centre both atom sets,
// Horn's quaternion form of the optimal rotation (largest eigenvector of a symmetric 4x4 built
// from the 3x3 correlation matrix, Jacobi sweeps in double), x' = R (x - c_frame) + c_ref.
// Two kernels.  nrl_superpose_kernel, one CTA per frame: block reductions over the FITTED atoms for the centroid and
// the nine correlation sums, one thread solves the eigenproblem and leaves the frame's transform (rotation, centroid)
// in global memory.  nrl_superpose_apply_kernel rewrites the frames in place: the HBM-bound half, 24 bytes per atom and
// frame (read + write of three floats), a grid over (chunks of atoms, frames) with four atoms -- three 16-byte words --
// per thread and step, so that its bandwidth does not depend on the number of frames (one CTA per frame moved
// 1.8 TB/s on 200 frames x 200 k atoms).
#pragma once
#include "common.cuh"

#define NRL_SUP_NT 512

__device__ __forceinline__ double nrl_block_sum_double(double v, double *s_red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < NRL_SUP_NT / 32; ++w) t += s_red[w];     // same order in every thread: deterministic
  return t;
}

// reference set: centroid + centred coordinates of the fitted atoms of frame `ref`
template <typename T>
__global__ void __launch_bounds__(NRL_SUP_NT) nrl_superpose_ref_kernel(const T *__restrict__ xyz, const int n_atoms,
                                                                        const int *__restrict__ idx, const int n_fit,
                                                                        const int ref, double *__restrict__ ref_set) {
  __shared__ double s_red[NRL_SUP_NT / 32];
  const T *f = xyz + (long long)ref * n_atoms * 3;
  double v[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < n_fit; i += NRL_SUP_NT) {
    const long long a = idx[i];
    v[0] += (double)f[3 * a]; v[1] += (double)f[3 * a + 1]; v[2] += (double)f[3 * a + 2];
  }
  double c[3];
  for (int k = 0; k < 3; ++k) c[k] = nrl_block_sum_double(v[k], s_red) / (double)n_fit;
  if (threadIdx.x < 3) ref_set[threadIdx.x] = c[threadIdx.x];
  double sq = 0.0;
  for (int i = threadIdx.x; i < n_fit; i += NRL_SUP_NT)
    for (int k = 0; k < 3; ++k) { const double r = (double)f[3 * (long long)idx[i] + k] - c[k]; ref_set[3 + 3 * i + k] = r; sq += r * r; }
  sq = nrl_block_sum_double(sq, s_red);
  if (threadIdx.x == 0) ref_set[3 + 3 * (long long)n_fit] = sq;     // sum |r|^2, for the RMSD of every frame
}

// largest-eigenvalue eigenvector of the symmetric 4x4 `a` (cyclic Jacobi), returned as a unit quaternion
__device__ void nrl_jacobi4_top(double a[4][4], double q[4]) {
  double v[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    double offd = 0.0;
    for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) offd += a[i][j] * a[i][j];
    if (offd < 1e-30) break;
    for (int p = 0; p < 3; ++p)
      for (int r = p + 1; r < 4; ++r) {
        if (fabs(a[p][r]) < 1e-300) continue;
        const double theta = (a[r][r] - a[p][p]) / (2.0 * a[p][r]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
        for (int k = 0; k < 4; ++k) { const double akp = a[k][p], akr = a[k][r]; a[k][p] = cs * akp - sn * akr; a[k][r] = sn * akp + cs * akr; }
        for (int k = 0; k < 4; ++k) { const double apk = a[p][k], ark = a[r][k]; a[p][k] = cs * apk - sn * ark; a[r][k] = sn * apk + cs * ark; }
        for (int k = 0; k < 4; ++k) { const double vkp = v[k][p], vkr = v[k][r]; v[k][p] = cs * vkp - sn * vkr; v[k][r] = sn * vkp + cs * vkr; }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i) if (a[i][i] > a[best][best]) best = i;
  double nrm = 0.0;
  for (int k = 0; k < 4; ++k) { q[k] = v[k][best]; nrm += q[k] * q[k]; }
  nrm = 1.0 / sqrt(nrm);
  for (int k = 0; k < 4; ++k) q[k] *= nrm;
}

template <typename T>
__global__ void __launch_bounds__(NRL_SUP_NT) nrl_superpose_kernel(T *__restrict__ xyz, const int n_atoms,
                                                                    const int *__restrict__ idx, const int n_fit,
                                                                    const double *__restrict__ ref_set, const int ref,
                                                                    float *__restrict__ rmsd_out, double *__restrict__ xform) {
  __shared__ double s_part[NRL_SUP_NT / 32][13];
  __shared__ double s_rot[9];
  const int frame = blockIdx.x;
  T *f = xyz + (long long)frame * n_atoms * 3;
  // ONE pass over the fitted atoms (a strided gather: every atom is fetched once): sum x, sum x r^T, sum |x|^2.  The
  // reference set is centred (sum r = 0), so the correlation of the CENTRED frame, sum (x - c) r^T, is sum x r^T itself,
  // and sum |x - c|^2 = sum |x|^2 - n |c|^2; all sums in double.
  double acc[13];
#pragma unroll
  for (int j = 0; j < 13; ++j) acc[j] = 0.0;
  for (int i = threadIdx.x; i < n_fit; i += NRL_SUP_NT) {
    const long long a = idx[i];
    const double x[3] = {(double)f[3 * a], (double)f[3 * a + 1], (double)f[3 * a + 2]};
    const double *r = ref_set + 3 + 3 * (long long)i;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      acc[j] += x[j];
#pragma unroll
      for (int k = 0; k < 3; ++k) acc[3 + 3 * j + k] += x[j] * r[k];
    }
    acc[12] += x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
  }
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 13; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
      if (lane == 0) s_part[warp][j] = acc[j];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t13[13];
    for (int j = 0; j < 13; ++j) {
      double t = 0.0;
      for (int w = 0; w < NRL_SUP_NT / 32; ++w) t += s_part[w][j];     // fixed order: deterministic
      t13[j] = t;
    }
    const double c[3] = {t13[0] / (double)n_fit, t13[1] / (double)n_fit, t13[2] / (double)n_fit};
    const double *m = t13 + 3;
    const double g = t13[12] - (double)n_fit * (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) + ref_set[3 + 3 * (long long)n_fit];
    const double Sxx = m[0], Sxy = m[1], Sxz = m[2], Syx = m[3], Syy = m[4], Syz = m[5], Szx = m[6], Szy = m[7], Szz = m[8];
    double a[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                      {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                      {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                      {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
    double q[4];
    nrl_jacobi4_top(a, q);
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    // rotation that takes the frame onto the reference
    s_rot[0] = w * w + x * x - y * y - z * z; s_rot[1] = 2 * (x * y - w * z);           s_rot[2] = 2 * (x * z + w * y);
    s_rot[3] = 2 * (x * y + w * z);           s_rot[4] = w * w - x * x + y * y - z * z; s_rot[5] = 2 * (y * z - w * x);
    s_rot[6] = 2 * (x * z - w * y);           s_rot[7] = 2 * (y * z + w * x);           s_rot[8] = w * w - x * x - y * y + z * z;
    if (rmsd_out) {
      // residual of the best fit: (G - 2 * lambda_max) / n, lambda_max = q^T N q
      double lam = 0.0;
      const double n4[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                               {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                               {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                               {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) lam += q[i] * n4[i][j] * q[j];
      rmsd_out[frame] = (float)sqrt(fmax(0.0, (g - 2.0 * lam) / (double)n_fit));
    }
    // the frame's transform for nrl_superpose_apply_kernel: rotation, frame centroid
    double *t = xform + 12 * (long long)frame;
    for (int j = 0; j < 9; ++j) t[j] = s_rot[j];
    t[9] = c[0]; t[10] = c[1]; t[11] = c[2];
  }
}

// x' = R (x - c_frame) + c_ref for every atom of every frame but the reference one (which stays as it is, bit for bit);
// the same operations in the same order, in double, as when the fit kernel did this itself
#define NRL_SUPA_NT 256
template <typename T> struct NrlVec16 { typedef float4 type; static constexpr int n = 4; };
template <> struct NrlVec16<double> { typedef double2 type; static constexpr int n = 2; };

template <typename T>
__global__ void __launch_bounds__(NRL_SUPA_NT) nrl_superpose_apply_kernel(T *__restrict__ xyz, const int n_atoms, const double *__restrict__ xform,
                                                                          const double *__restrict__ ref_set, const int ref) {
  const int frame = blockIdx.y;
  if (frame == ref) return;
  T *f = xyz + (long long)frame * n_atoms * 3;
  const double *t = xform + 12 * (long long)frame;
  double R[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) R[j] = t[j];
  const double c0 = t[9], c1 = t[10], c2 = t[11];
  const double cr0 = ref_set[0], cr1 = ref_set[1], cr2 = ref_set[2];
  auto move = [&](T &px, T &py, T &pz) {
    const double x = (double)px - c0, y = (double)py - c1, z = (double)pz - c2;
    px = (T)(R[0] * x + R[1] * y + R[2] * z + cr0);
    py = (T)(R[3] * x + R[4] * y + R[5] * z + cr1);
    pz = (T)(R[6] * x + R[7] * y + R[8] * z + cr2);
  };
  typedef typename NrlVec16<T>::type V;
  constexpr int words = 12 * (int)sizeof(T) / 16;          // 16-byte words of four atoms: 3 (float32), 6 (float64)
  const int n4 = ((reinterpret_cast<uintptr_t>(f) & 15) == 0) ? n_atoms / 4 : 0;   // groups of four atoms read as whole words
  for (int q = blockIdx.x * NRL_SUPA_NT + threadIdx.x; q < n4; q += gridDim.x * NRL_SUPA_NT) {
    V *w = reinterpret_cast<V *>(f + 12 * (long long)q);
    union { V v[words]; T c[12]; } u;
#pragma unroll
    for (int k = 0; k < words; ++k) u.v[k] = w[k];
#pragma unroll
    for (int a = 0; a < 4; ++a) move(u.c[3 * a], u.c[3 * a + 1], u.c[3 * a + 2]);
#pragma unroll
    for (int k = 0; k < words; ++k) w[k] = u.v[k];
  }
  for (int i = 4 * n4 + blockIdx.x * NRL_SUPA_NT + threadIdx.x; i < n_atoms; i += gridDim.x * NRL_SUPA_NT)
    move(f[3 * i], f[3 * i + 1], f[3 * i + 2]);
}
