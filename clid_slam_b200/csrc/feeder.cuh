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
// ---------------------------------------------------------------------------------------------------------
// brick index build (ClidBricks): which points own their voxel's hash slot and pass the
// per-point predicates, their bounding box, their (brick, cell) sort keys, and after the sort the headers /
// records / neighbourhood lines.  Replaces ~20 eager torch ops and three host synchronisations per frame
// (ops/bricks.py) by four launches around one library sort and ONE 32-byte read-back (the bounding box sizes
// the header array).
// ---------------------------------------------------------------------------------------------------------
struct BrickKeyParams {
  const float* points;           // [n,3] candidate points (local window or the whole map)
  const int64_t* gids;           // [n] global id of every candidate, or NULL (= its index)
  const int64_t* table;          // voxel hash
  int64_t buffer_size;
  int64_t primes[3];
  const int32_t* ts_create;      // [n_global] or NULL (no time filter)
  const float* travel_dist;
  int32_t cur_ts;
  float diff_travel_dist_local;
  float resolution;
  int64_t n;
  int32_t* cells;                // [n,3] out: voxel cell of every candidate
  uint8_t* keep;                 // [n]   out
  int32_t* bbox;                 // [7]   in/out: min xyz (init INT_MAX), max xyz (init INT_MIN), kept count
};

__global__ void __launch_bounds__(256) brick_keep_kernel(const BrickKeyParams p) {
  int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN}, cnt = 0;
  float td_cur = 0.f;
  if (p.ts_create) td_cur = p.travel_dist[p.cur_ts];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int cx = cell_of(p.points[3 * i], p.resolution), cy = cell_of(p.points[3 * i + 1], p.resolution),
              cz = cell_of(p.points[3 * i + 2], p.resolution);
    const int64_t gid = p.gids ? p.gids[i] : i;
    const int64_t slot = floor_mod((int64_t)cx * p.primes[0] + (int64_t)cy * p.primes[1] + (int64_t)cz * p.primes[2], p.buffer_size);
    bool keep = p.table[slot] == gid;  // the point owns its voxel's slot (orphans are unreachable through the table)
    if (keep && p.ts_create) keep = fabsf(td_cur - p.travel_dist[p.ts_create[gid]]) < p.diff_travel_dist_local;
    p.cells[3 * i] = cx; p.cells[3 * i + 1] = cy; p.cells[3 * i + 2] = cz;
    p.keep[i] = keep ? 1 : 0;
    if (keep) {
      lo[0] = min(lo[0], cx); lo[1] = min(lo[1], cy); lo[2] = min(lo[2], cz);
      hi[0] = max(hi[0], cx); hi[1] = max(hi[1], cy); hi[2] = max(hi[2], cz);
      ++cnt;
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
    hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
  }
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) { atomicMin(p.bbox + a, lo[a]); atomicMax(p.bbox + 3 + a, hi[a]); }
    atomicAdd(p.bbox + 6, cnt);
  }
}

// sort key (brick * 64 + cell bit) of every kept candidate, INT64_MAX for the others
__global__ void __launch_bounds__(256) brick_key_kernel(const int32_t* __restrict__ cells, const uint8_t* __restrict__ keep, int64_t n,
                                                        int lox, int loy, int loz, int d0, int d1, int64_t* __restrict__ keys) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t key = INT64_MAX;
    if (keep[i]) {
      const int rx = cells[3 * i] - lox, ry = cells[3 * i + 1] - loy, rz = cells[3 * i + 2] - loz;
      const int64_t brick = (rx >> 2) + (int64_t)d0 * ((ry >> 2) + (int64_t)d1 * (rz >> 2));
      key = brick * 64 + ((rx & 3) + 4 * (ry & 3) + 16 * (rz & 3));
    }
    keys[i] = key;
  }
}

// records in sorted order + headers {mask, base = first record, count}; headers pre-set to {0, INT_MAX, 0}
__global__ void __launch_bounds__(256) brick_scatter_kernel(const int64_t* __restrict__ sorted_keys, const int64_t* __restrict__ order,
                                                            int64_t n_kept, const float* __restrict__ points,
                                                            const int32_t* __restrict__ rows, float4* __restrict__ records,
                                                            ClidBrickHeader* __restrict__ headers) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_kept; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t key = sorted_keys[j], src = order[j];
    const int64_t brick = key >> 6;
    const int bit = (int)(key & 63);
    const int row = rows ? rows[src] : (int)src;
    records[j] = make_float4(points[3 * src], points[3 * src + 1], points[3 * src + 2], __int_as_float(row));
    ClidBrickHeader* h = headers + brick;
    atomicOr(reinterpret_cast<unsigned long long*>(&h->mask), 1ull << bit);
    atomicMin(&h->base, (int)j);
    atomicAdd(&h->count, 1);
  }
}

// empty bricks keep base INT_MAX from the initialisation: normalise to 0; then the 128-byte neighbourhood lines
__global__ void __launch_bounds__(256) brick_hood_kernel(ClidBrickHeader* __restrict__ headers, int d0, int d1, int d2,
                                                         uint32_t* __restrict__ hood) {
  const int64_t nb = (int64_t)d0 * d1 * d2;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nb * 8; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = t >> 3;
    const int s = (int)(t & 7);
    const int x = (int)(b % d0), y = (int)((b / d0) % d1), z = (int)(b / ((int64_t)d0 * d1));
    const int nx = x + (s & 1), ny = y + ((s >> 1) & 1), nz = z + (s >> 2);
    uint32_t lo = 0u, hi = 0u, base = 0u;
    if (nx < d0 && ny < d1 && nz < d2) {
      const ClidBrickHeader h = headers[((int64_t)nz * d1 + ny) * d0 + nx];
      lo = (uint32_t)h.mask; hi = (uint32_t)(h.mask >> 32);
      base = h.count > 0 ? (uint32_t)h.base : 0u;
    }
    if (hood) {
      uint32_t* line = hood + b * 32;
      line[2 * s] = lo; line[2 * s + 1] = hi; line[16 + s] = base;
      if (s == 0) {
#pragma unroll
        for (int k = 24; k < 32; ++k) line[k] = 0u;
      }
    }
  }
}

__global__ void __launch_bounds__(256) brick_header_init_kernel(ClidBrickHeader* __restrict__ headers, int64_t nb) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) headers[b].base = INT_MAX;
}

__global__ void __launch_bounds__(256) brick_header_fix_kernel(ClidBrickHeader* __restrict__ headers, int64_t nb) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x)
    if (headers[b].count == 0) headers[b].base = 0;
}
#endif  // CLID_PLAIN_KERNELS

}  // namespace clid
