// Kernels of the per-frame feeders: the steps either side of the hot path that the
// reference runs as chains of eager ATen ops once per scan.
//   region_sdf_kernel   LocalPointCloudMap.region_specific_sdf_estimation (model/local_point_cloud_map.py:98-152)
//                       + estimate_plane (:155-201): 7-cell probe of the raw-point voxel hash, 4 nearest stored
//                       points, least-squares plane through them, point-to-plane distance (or nearest-point
//                       distance when the fit is rejected) -- the region-specific SDF label of CLID-SLAM
#pragma once
#include "common.cuh"

namespace clid {

#ifdef CLID_PLAIN_KERNELS
struct RegionSdfParams {
  const float* points;          // [n,3] world frame
  const int64_t* table;         // [buffer_size] voxel hash of the local point-cloud map, -1 empty
  const float* map_points;      // [m,3]
  const int64_t* neighbor_idx;  // [kc,3] cell offsets
  int64_t n, buffer_size, m;
  int64_t primes[3];
  int32_t kc;
  float resolution, max_valid_range;
  float eta_threshold, dist_threshold;
  float* sdf_abs;               // [n]
  uint8_t* surface_mask;        // [n]
};

// eigen-decomposition of a symmetric 3x3 matrix by cyclic Jacobi rotations (double: the matrix is A^T A of the
// centred neighbours, whose small eigenvalue decides whether the neighbourhood is flat)
__device__ __forceinline__ void jacobi3(double (&a)[3][3], double (&v)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 8; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off < 1e-30) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      if (fabs(a[p][q]) < 1e-300) continue;
      const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // A <- A J
        const double akp = a[k][p], akq = a[k][q];
        a[k][p] = c * akp - s * akq;
        a[k][q] = s * akp + c * akq;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {  // A <- J^T A
        const double apk = a[p][k], aqk = a[q][k];
        a[p][k] = c * apk - s * aqk;
        a[q][k] = s * apk + c * aqk;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double vkp = v[k][p], vkq = v[k][q];
        v[k][p] = c * vkp - s * vkq;
        v[k][q] = s * vkp + c * vkq;
      }
    }
  }
}

__global__ void __launch_bounds__(128) region_sdf_kernel(const RegionSdfParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float px = p.points[3 * i], py = p.points[3 * i + 1], pz = p.points[3 * i + 2];
    const int64_t cx = cell_of(px, p.resolution), cy = cell_of(py, p.resolution), cz = cell_of(pz, p.resolution);
    // four nearest of the probed cells; empty cells count with max_valid_range (torch.where(idx == -1, far, dist))
    float d4[4] = {3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f};
    float q4[4][3] = {};
    for (int c = 0; c < p.kc; ++c) {
      const int64_t gx = cx + p.neighbor_idx[3 * c], gy = cy + p.neighbor_idx[3 * c + 1], gz = cz + p.neighbor_idx[3 * c + 2];
      const int64_t slot = floor_mod(gx * p.primes[0] + gy * p.primes[1] + gz * p.primes[2], p.buffer_size);
      const int64_t id = p.table[slot];
      float d = p.max_valid_range, qx = 0.f, qy = 0.f, qz = 0.f;
      if (id >= 0) {
        qx = p.map_points[3 * id]; qy = p.map_points[3 * id + 1]; qz = p.map_points[3 * id + 2];
        const float dx = qx - px, dy = qy - py, dz = qz - pz;
        d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));  // torch.norm
      }
      if (d < d4[3]) {  // insertion into the ascending list; ties keep the earlier cell
        int pos = 3;
        while (pos > 0 && d < d4[pos - 1]) { d4[pos] = d4[pos - 1]; q4[pos][0] = q4[pos - 1][0]; q4[pos][1] = q4[pos - 1][1]; q4[pos][2] = q4[pos - 1][2]; --pos; }
        d4[pos] = d; q4[pos][0] = qx; q4[pos][1] = qy; q4[pos][2] = qz;
      }
    }
    const float far = p.max_valid_range;
    float out = d4[0];
    if (d4[3] < far) {  // four stored neighbours: fit a plane (estimate_plane)
      const double mx = ((double)q4[0][0] + q4[1][0] + q4[2][0] + q4[3][0]) * 0.25;
      const double my = ((double)q4[0][1] + q4[1][1] + q4[2][1] + q4[3][1]) * 0.25;
      const double mz = ((double)q4[0][2] + q4[1][2] + q4[2][2] + q4[3][2]) * 0.25;
      double a[3][3] = {}, v[3][3];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double x = q4[k][0] - mx, y = q4[k][1] - my, z = q4[k][2] - mz;
        a[0][0] += x * x; a[0][1] += x * y; a[0][2] += x * z; a[1][1] += y * y; a[1][2] += y * z; a[2][2] += z * z;
      }
      a[1][0] = a[0][1]; a[2][0] = a[0][2]; a[2][1] = a[1][2];
      jacobi3(a, v);
      // singular values of the centred 4 x 3 matrix = sqrt of the eigenvalues; smallest -> normal, middle -> eta
      double lam[3] = {a[0][0], a[1][1], a[2][2]};
      int lo = 0;
      if (lam[1] < lam[lo]) lo = 1;
      if (lam[2] < lam[lo]) lo = 2;
      int hi = 0;
      if (lam[1] > lam[hi]) hi = 1;
      if (lam[2] > lam[hi]) hi = 2;
      const int mid = 3 - lo - hi >= 0 && lo != hi ? 3 - lo - hi : (lo + 1) % 3;
      const double s_min = sqrt(fmax(lam[lo], 0.0)), s_mid = sqrt(fmax(lam[mid], 0.0));
      const bool flat = s_min / (s_mid + 1e-6) <= (double)p.eta_threshold;
      if (flat) {
        const double nx = v[0][lo], ny = v[1][lo], nz = v[2][lo];
        const double pc = -(nx * mx + ny * my + nz * mz);
        double worst = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) worst = fmax(worst, fabs(nx * q4[k][0] + ny * q4[k][1] + nz * q4[k][2] + pc));
        if (worst <= (double)p.dist_threshold) out = (float)fabs(nx * px + ny * py + nz * pz + pc);
      }
    }
    p.sdf_abs[i] = out;
    p.surface_mask[i] = d4[0] < far ? 1 : 0;
  }
}
#endif  // CLID_PLAIN_KERNELS

}  // namespace clid
