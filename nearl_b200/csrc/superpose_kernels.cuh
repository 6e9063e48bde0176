// nearl_b200 -- rigid-body superposition of trajectory frames on the device.
//
// The step in front of the hot path: MisatoTraj(..., superpose=True) fits every frame onto the
// first one over the C-alpha atoms (pt.superpose(traj, mask="@CA"),
// /root/reference/nearl/io/traj.py:398-404; the arithmetic itself lives in cpptraj, which is not
// part of the reference tree).  Restated as the textbook least-squares fit: centre both atom sets,
// Horn's quaternion form of the optimal rotation (largest eigenvector of a symmetric 4x4 built
// from the 3x3 correlation matrix, Jacobi sweeps in double), x' = R (x - c_frame) + c_ref.
// One CTA per frame: block reductions for the centroid and the nine correlation sums, one thread
// solves the eigenproblem, every thread rewrites its share of the frame in place.  The kernel is
// HBM bound: 24 bytes per atom per frame (read + write of three floats).
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
  double c[3];
  for (int k = 0; k < 3; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n_fit; i += NRL_SUP_NT) v += (double)f[3 * idx[i] + k];
    c[k] = nrl_block_sum_double(v, s_red) / (double)n_fit;
  }
  if (threadIdx.x < 3) ref_set[threadIdx.x] = c[threadIdx.x];
  for (int i = threadIdx.x; i < n_fit; i += NRL_SUP_NT)
    for (int k = 0; k < 3; ++k) ref_set[3 + 3 * i + k] = (double)f[3 * idx[i] + k] - c[k];
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
                                                                    float *__restrict__ rmsd_out) {
  __shared__ double s_red[NRL_SUP_NT / 32];
  __shared__ double s_rot[9];
  const int frame = blockIdx.x;
  T *f = xyz + (long long)frame * n_atoms * 3;
  double c[3];
  for (int k = 0; k < 3; ++k) {
    double v = 0.0;
    for (int i = threadIdx.x; i < n_fit; i += NRL_SUP_NT) v += (double)f[3 * idx[i] + k];
    c[k] = nrl_block_sum_double(v, s_red) / (double)n_fit;
  }
  // correlation matrix M[j][k] = sum (x_frame - c)_j (x_ref - c_ref)_k, and the two squared norms for the RMSD
  double m[9], g = 0.0;
  for (int j = 0; j < 9; ++j) m[j] = 0.0;
  for (int i = threadIdx.x; i < n_fit; i += NRL_SUP_NT) {
    const double x[3] = {(double)f[3 * idx[i]] - c[0], (double)f[3 * idx[i] + 1] - c[1], (double)f[3 * idx[i] + 2] - c[2]};
    const double *r = ref_set + 3 + 3 * i;
    for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) m[3 * j + k] += x[j] * r[k];
    g += x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  }
  for (int j = 0; j < 9; ++j) m[j] = nrl_block_sum_double(m[j], s_red);
  g = nrl_block_sum_double(g, s_red);
  if (threadIdx.x == 0) {
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
  }
  __syncthreads();
  if (frame == ref) return;                       // the reference frame stays as it is, bit for bit
  double R[9];
  for (int j = 0; j < 9; ++j) R[j] = s_rot[j];
  const double cr0 = ref_set[0], cr1 = ref_set[1], cr2 = ref_set[2];
  for (int i = threadIdx.x; i < n_atoms; i += NRL_SUP_NT) {
    const double x = (double)f[3 * i] - c[0], y = (double)f[3 * i + 1] - c[1], z = (double)f[3 * i + 2] - c[2];
    f[3 * i] = (T)(R[0] * x + R[1] * y + R[2] * z + cr0);
    f[3 * i + 1] = (T)(R[3] * x + R[4] * y + R[5] * z + cr1);
    f[3 * i + 2] = (T)(R[6] * x + R[7] * y + R[8] * z + cr2);
  }
}
