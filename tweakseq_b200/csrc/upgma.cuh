// upgma.cuh -- guide tree from the distance matrix: UPGMA, sm_100a.
//
// The next consumer of the distance matrix: what clustalo builds from --distmat-in before its
// progressive alignment.  Spec (build-defined, restated by oracle/gotoh_oracle.c:tsq_oracle_upgma):
//   * clusters live in slots 0..n-1; every step merges the active pair (a < b) with the smallest
//     distance, ties broken by smallest a, then smallest b; the merged cluster keeps slot a;
//   * d(a,k) <- (|a|*d(a,k) + |b|*d(b,k)) / (|a|+|b|), four separately rounded fp64 operations;
//   * node ids: leaves 0..n-1, the node created by step t is n+t; its height is d(a,b)/2.
//
// Device layout: a full symmetric n x n fp64 matrix (rows contiguous) plus, per row i, the minimum
// over active j > i and its argument.  One persistent CTA runs the n-1 sequential steps; each step
// is a block-wide lexicographic arg-min over the row minima, an O(n) row update, and a re-scan of
// the few rows whose cached minimum pointed at a or b.  The work per step is O(n) + re-scans, the
// whole tree O(n^2) instead of the naive O(n^3).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/tsq_b200.h"

#ifdef __CUDACC__
#define TSQ_UPGMA_HD __host__ __device__
#else
#define TSQ_UPGMA_HD
#endif

namespace tsq {

struct UpgmaParams {
  const double* dist;   // packed upper triangle, submitted order
  double* D;            // n x n workspace
  double* rowmin;       // [n]
  uint32_t* rowarg;     // [n]
  uint32_t* active;     // [n] 1 = slot in use
  uint32_t* csize;      // [n] leaves under the slot
  uint32_t* node;       // [n] node id currently held by the slot
  double* nheight;      // [n] height of that node
  tsq_merge* merges;    // [n-1] output
  uint32_t* rescan;     // [n+1] rows to re-scan in the current step
  uint32_t n;
};

#ifdef TSQ_DEVICE_IMPL  // kernel definitions: only tsq_device.cu instantiates them

__device__ __forceinline__ bool upgma_less(double va, uint32_t ia, uint32_t ja, double vb, uint32_t ib, uint32_t jb) {
  if (va != vb) return va < vb;
  if (ia != ib) return ia < ib;
  return ja < jb;
}

// fill the square matrix and the initial row minima; one CTA per row
__global__ void __launch_bounds__(256) upgma_init_kernel(const __grid_constant__ UpgmaParams p) {
  const unsigned long long n = p.n;
  __shared__ double sv[256];
  __shared__ uint32_t sj[256];
  for (unsigned long long i = blockIdx.x; i < n; i += gridDim.x) {
    double best = __longlong_as_double(0x7ff0000000000000ll);  // +inf
    uint32_t bj = 0xffffffffu;
    for (unsigned long long j = threadIdx.x; j < n; j += blockDim.x) {
      double v;
      if (j == i) v = __longlong_as_double(0x7ff0000000000000ll);
      else {
        const unsigned long long a = i < j ? i : j, b = i < j ? j : i;
        v = p.dist[a * n - a * (a + 1) / 2 + (b - a - 1)];
      }
      p.D[i * n + j] = v;
      if (j > i && (v < best || (v == best && (uint32_t)j < bj))) { best = v; bj = (uint32_t)j; }
    }
    sv[threadIdx.x] = best;
    sj[threadIdx.x] = bj;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) {
        const double v2 = sv[threadIdx.x + s];
        const uint32_t j2 = sj[threadIdx.x + s];
        if (v2 < sv[threadIdx.x] || (v2 == sv[threadIdx.x] && j2 < sj[threadIdx.x])) { sv[threadIdx.x] = v2; sj[threadIdx.x] = j2; }
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      p.rowmin[i] = sv[0];
      p.rowarg[i] = sj[0];
      p.active[i] = 1;
      p.csize[i] = 1;
      p.node[i] = (uint32_t)i;
      p.nheight[i] = 0.0;
    }
    __syncthreads();
  }
}

constexpr int UPGMA_THREADS = 1024;

// block-wide lexicographic arg-min of (v, i, j); result valid in every thread
__device__ __forceinline__ void upgma_block_argmin(double& v, uint32_t& i, uint32_t& j, double* sv, uint32_t* si,
                                                   uint32_t* sj) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
    const uint32_t i2 = __shfl_xor_sync(0xffffffffu, i, o);
    const uint32_t j2 = __shfl_xor_sync(0xffffffffu, j, o);
    if (upgma_less(v2, i2, j2, v, i, j)) { v = v2; i = i2; j = j2; }
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int nwarps = (int)(blockDim.x >> 5);
  __syncthreads();
  if (l == 0) { sv[w] = v; si[w] = i; sj[w] = j; }
  __syncthreads();
  if (w == 0) {
    if (l < nwarps) { v = sv[l]; i = si[l]; j = sj[l]; }
    else { v = __longlong_as_double(0x7ff0000000000000ll); i = 0xffffffffu; j = 0xffffffffu; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double v2 = __shfl_xor_sync(0xffffffffu, v, o);
      const uint32_t i2 = __shfl_xor_sync(0xffffffffu, i, o);
      const uint32_t j2 = __shfl_xor_sync(0xffffffffu, j, o);
      if (upgma_less(v2, i2, j2, v, i, j)) { v = v2; i = i2; j = j2; }
    }
    if (l == 0) { sv[0] = v; si[0] = i; sj[0] = j; }
  }
  __syncthreads();
  v = sv[0]; i = si[0]; j = sj[0];
}

// SMEM: row minima, their arguments and the active flags live in shared memory (n <= UPGMA_SMEM_N);
// every step then touches global memory only for the two matrix rows it combines and for re-scans.
constexpr uint32_t UPGMA_SMEM_N = 12288;   // 12288 * (8 + 4 + 1) B = 156 KB
// ... and up to this n the cluster sizes and node ids too (8 B more per slot): they are read right behind the arg-min of
// every step, two L2 round trips on the critical path of a 6 us step otherwise
constexpr uint32_t UPGMA_SMEM_N2 = 8192;
TSQ_UPGMA_HD inline size_t upgma_smem_bytes(uint32_t n) {
  size_t b = (size_t)n * (sizeof(double) + sizeof(uint32_t) + 1) + 16;
  if (n <= UPGMA_SMEM_N2) b = ((b + 15) & ~(size_t)15) + (size_t)n * 8;
  return b;
}

template <bool SMEM>
__global__ void __launch_bounds__(UPGMA_THREADS, 1) upgma_merge_kernel(const __grid_constant__ UpgmaParams p) {
  const uint32_t n = p.n;
  const double INF = __longlong_as_double(0x7ff0000000000000ll);
  __shared__ double sv[32];
  __shared__ uint32_t si[32], sj[32];
  __shared__ uint32_t nrescan;
  __shared__ double pv[32][32];       // re-scan partials: [row in batch][warp]
  __shared__ uint32_t pj[32][32];
  extern __shared__ double upgma_dyn[];
  double* rowmin = SMEM ? upgma_dyn : p.rowmin;
  uint32_t* rowarg = SMEM ? reinterpret_cast<uint32_t*>(upgma_dyn + n) : p.rowarg;
  uint8_t* act8 = reinterpret_cast<uint8_t*>(rowarg + n);   // SMEM only
  const bool small = SMEM && n <= UPGMA_SMEM_N2;            // cluster sizes and node ids in shared memory as well
  uint32_t* const csize = small ? reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(upgma_dyn) + (((size_t)n * 13 + 16 + 15) & ~(size_t)15)) : p.csize;
  uint32_t* const node = small ? csize + n : p.node;
  uint32_t* const rescan = p.rescan;
  const uint32_t tid = threadIdx.x;
  const uint32_t NT = blockDim.x;   // UPGMA_THREADS unless a tuning run asks for fewer (upgma_launch)
  if (SMEM) {
    for (uint32_t i = tid; i < n; i += NT) {
      rowmin[i] = p.rowmin[i];
      rowarg[i] = p.rowarg[i];
      act8[i] = 1;
      if (small) {
        csize[i] = 1;
        node[i] = i;
      }
    }
    __syncthreads();
  }
  auto is_active = [&](uint32_t i) -> bool { return SMEM ? act8[i] != 0 : p.active[i] != 0; };

  for (uint32_t step = 0; step + 1 < n; ++step) {
    // ---- 1. the closest active pair --------------------------------------------------------------
    double v = INF;
    uint32_t a = 0xffffffffu, b = 0xffffffffu;
    for (uint32_t i = tid; i < n; i += NT) {
      if (is_active(i)) {
        const double rv = rowmin[i];
        const uint32_t rj = rowarg[i];
        if (rj != 0xffffffffu && upgma_less(rv, i, rj, v, a, b)) { v = rv; a = i; b = rj; }
      }
    }
    upgma_block_argmin(v, a, b, sv, si, sj);
    // ---- 2. record the merge ----------------------------------------------------------------------
    const uint32_t sa = csize[a], sb = csize[b];
    if (tid == 0) {
      tsq_merge mg;
      mg.left = node[a];
      mg.right = node[b];
      mg.height = __dmul_rn(v, 0.5);
      p.merges[step] = mg;
      nrescan = 0;
    }
    __syncthreads();
    // ---- 3. new distances of slot a; slot b retires ------------------------------------------------
    const double da = (double)sa, db = (double)sb, dsum = (double)(sa + sb);
    double* Da = p.D + (size_t)a * n;
    const double* Db = p.D + (size_t)b * n;
    for (uint32_t k = tid; k < n; k += NT) {
      if (k == a || k == b || !is_active(k)) continue;
      const double nd = __ddiv_rn(__dadd_rn(__dmul_rn(da, Da[k]), __dmul_rn(db, Db[k])), dsum);
      Da[k] = nd;
      p.D[(size_t)k * n + a] = nd;
      if (k < a) {
        // entry (k,a) of row k changed, entry (k,b) disappears
        const uint32_t rk = rowarg[k];
        if (rk == a || rk == b) {
          const uint32_t pos = atomicAdd(&nrescan, 1u);
          rescan[pos] = k;
        } else if (nd < rowmin[k] || (nd == rowmin[k] && a < rk)) {
          rowmin[k] = nd;
          rowarg[k] = a;
        }
      } else if (k < b) {
        // a < k < b: row k loses entry (k,b); entry (a,k) belongs to row a (re-scanned below)
        if (rowarg[k] == b) {
          const uint32_t pos = atomicAdd(&nrescan, 1u);
          rescan[pos] = k;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (SMEM) act8[b] = 0; else p.active[b] = 0;
      csize[a] = sa + sb;
      node[a] = n + step;
      rescan[nrescan] = a;   // row a always
      nrescan = nrescan + 1;
    }
    __syncthreads();
    // ---- 4. re-scan the rows whose cached minimum is stale -----------------------------------------
    // Batched: every warp reduces its slice of each stale row with shuffles and parks the partial
    // result in shared memory; after ONE barrier, warp r finishes row r.  (A block-wide reduction per
    // row cost three barriers each and made the barriers the kernel's top stall.)
    const uint32_t nr = nrescan;
    const uint32_t nwarps = NT >> 5;   // a batch: as many rows as the CTA has warps
    for (uint32_t r0 = 0; r0 < nr; r0 += nwarps) {
      const uint32_t nb = nr - r0 < nwarps ? nr - r0 : nwarps;
#pragma unroll 4
      for (uint32_t r = 0; r < nb; ++r) {
        const uint32_t row = rescan[r0 + r];
        const double* Dr = p.D + (size_t)row * n;
        double bv = INF;
        uint32_t bj = 0xffffffffu;
        for (uint32_t j = row + 1 + tid; j < n; j += NT) {
          if (is_active(j)) {
            const double x = Dr[j];
            if (x < bv || (x == bv && j < bj)) { bv = x; bj = j; }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
          const uint32_t j2 = __shfl_xor_sync(0xffffffffu, bj, o);
          if (v2 < bv || (v2 == bv && j2 < bj)) { bv = v2; bj = j2; }
        }
        if ((tid & 31) == 0) { pv[r][tid >> 5] = bv; pj[r][tid >> 5] = bj; }
      }
      __syncthreads();
      if ((tid >> 5) < nb) {
        const uint32_t r = tid >> 5;
        double bv = (tid & 31) < nwarps ? pv[r][tid & 31] : INF;
        uint32_t bj = (tid & 31) < nwarps ? pj[r][tid & 31] : 0xffffffffu;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double v2 = __shfl_xor_sync(0xffffffffu, bv, o);
          const uint32_t j2 = __shfl_xor_sync(0xffffffffu, bj, o);
          if (v2 < bv || (v2 == bv && j2 < bj)) { bv = v2; bj = j2; }
        }
        if ((tid & 31) == 0) {
          const uint32_t row = rescan[r0 + r];
          rowmin[row] = bv;
          rowarg[row] = bj;
        }
      }
      __syncthreads();
    }
  }
}

#endif  // TSQ_DEVICE_IMPL

}  // namespace tsq
