// tsq_device.cu -- kernel instantiations and launchers (sm_100a only).
#define TSQ_DEVICE_IMPL 1
#include "tsq_device.h"

#include <cstdio>
#include <cstdlib>

namespace tsq {

// ---- packed 16-bit inter-task kernel variants -------------------------------------------
// (K, threads per CTA, CTAs per SM).  Register budget = 65536 / (tpb * ctas_sm).
#define TSQ_G16_VARIANTS(X) \
  X(30, 128, 4)             \
  X(32, 128, 4)             \
  X(36, 128, 3)             \
  X(40, 128, 3)             \
  X(44, 128, 3)             \
  X(48, 128, 3)             \
  X(50, 128, 3)             \
  X(52, 128, 3)             \
  X(56, 128, 3)

const int kStripWidths[kNumStripWidths] = {30, 32, 36, 40, 44, 48, 50, 52, 56};

bool g16_variant(int K, uint32_t nsym, G16Launch* out) {
#define X(KK, TT, MM)                                                                     \
  if (K == KK) {                                                                          \
    if (out) {                                                                            \
      out->K = KK;                                                                        \
      out->stride = G16Cfg<KK>::STRIDE;                                                   \
      out->tpb = TT;                                                                      \
      out->ctas_sm = MM;                                                                  \
      const size_t sbsz = ((size_t)(nsym + 1) * nsym + 31) & ~(size_t)31;                 \
      out->smem = (sbsz + (size_t)(TT / 32) * nsym * G16Cfg<KK>::STRIDE) * sizeof(uint32_t); \
    }                                                                                     \
    return true;                                                                          \
  }
  TSQ_G16_VARIANTS(X)
#undef X
  return false;
}

cudaError_t g16_launch(int K, int grid, const G16Params& p, cudaStream_t stream, const cudaAccessPolicyWindow* scratch_window) {
  G16Launch v;
  if (!g16_variant(K, p.nsym, &v)) return cudaErrorInvalidValue;
  // the strip-boundary scratch column gets an L2 access policy for THIS launch only: the share of it that fits the
  // persisting carve-out stays resident between the strip that writes a row and the strip that reads it back, the
  // rest streams (evict-first) instead of thrashing the whole cache (tsq_api.cpp: enqueue_gotoh16)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)v.tpb);
  cfg.dynamicSmemBytes = v.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = 0;
  if (scratch_window && scratch_window->num_bytes > 0) {
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow = *scratch_window;
    cfg.numAttrs = 1;
  }
#define X(KK, TT, MM)                                                                          \
  if (K == KK) {                                                                               \
    /* -ge' == +1 in both halves (ge = 1, delta = 2: the protein and nucleotide defaults) gets */ \
    /* the immediate-operand specialisation; every other gap model the generic kernel.        */ \
    auto kern = p.negge2 == 0x00010001u ? gotoh16_kernel<KK, TT, MM, 0x00010001u>              \
                                        : gotoh16_kernel<KK, TT, MM, 0u>;                      \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         (int)v.smem);                                         \
    if (e != cudaSuccess) return e;                                                            \
    return cudaLaunchKernelEx(&cfg, kern, p);                                                  \
  }
  TSQ_G16_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

cudaError_t g32_launch(int K, int grid, const G32Params& p, cudaStream_t stream) {
  G16Launch v;
  if (!g16_variant(K, p.nsym, &v)) return cudaErrorInvalidValue;
#define X(KK, TT, MM)                                                                          \
  if (K == KK) {                                                                               \
    auto kern = gotoh32_kernel<KK, TT, MM>;                                                    \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                         (int)v.smem);                                         \
    if (e != cudaSuccess) return e;                                                            \
    kern<<<grid, TT, v.smem, stream>>>(p);                                                     \
    return cudaGetLastError();                                                                 \
  }
  TSQ_G16_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

// ---- 32-bit wavefront kernel ----------------------------------------------------------------
// nucleotides (5 symbols): 24 columns per lane, 768 per pass, 15 KB of profile per warp, 12 warps/SM;
// proteins (23 symbols): 8 columns per lane, 256 per pass, 23.5 KB of profile per warp, 8 warps/SM.
#define TSQ_W32_VARIANTS(X) \
  X(32, 128, 2)             \
  X(24, 128, 3)             \
  X(16, 128, 4)             \
  X(16, 128, 3)             \
  X(8, 128, 4)              \
  X(8, 128, 2)

bool w32_variant(uint32_t nsym, W32Launch* out) {
  if (nsym == 0 || nsym > 24) return false;
  W32Launch v;
  // r01 sweep on 150 x 10-30 kb (profiles/): KW,CTAs = 24,3 -> 3728 GCUPS; 32,2 -> 3425; 16,4 -> 3267
  v.KW = nsym <= 8 ? 24 : 8;
  v.tpb = 128;
  v.ctas_sm = nsym <= 8 ? 3 : 2;
  if (const char* e = getenv("TSQ_FORCE_KW")) {  // developer override for tuning runs: "KW,ctas"
    int kw = 0, ct = 0;
    if (sscanf(e, "%d,%d", &kw, &ct) == 2) {
#define X(KK, TT, MM) if (kw == KK && ct == MM) { v.KW = KK; v.ctas_sm = MM; }
      TSQ_W32_VARIANTS(X)
#undef X
    }
  }
  const size_t sbsz = ((size_t)(nsym + 1) * nsym + 31) & ~(size_t)31;
  v.smem = (sbsz + (size_t)(v.tpb / 32) * nsym * 32 * v.KW) * sizeof(int32_t);
  if (out) *out = v;
  return true;
}

cudaError_t w32_launch(int grid, const W32Params& p, cudaStream_t stream) {
  W32Launch v;
  if (!w32_variant(p.nsym, &v)) return cudaErrorInvalidValue;
#define X(KK, TT, MM)                                                                        \
  if (v.KW == KK && v.ctas_sm == MM) {                                                       \
    auto kern = wave32_kernel<KK, TT, MM>;                                                   \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         (int)v.smem);                                       \
    if (e != cudaSuccess) return e;                                                          \
    kern<<<grid, TT, v.smem, stream>>>(p);                                                   \
    return cudaGetLastError();                                                               \
  }
  TSQ_W32_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

// ---- packed wavefront kernel -------------------------------------------------------------------
#define TSQ_W16_VARIANTS(X) \
  X(32, 128, 2)             \
  X(24, 128, 3)             \
  X(16, 128, 3)             \
  X(8, 128, 2)

// Columns per lane and CTAs per SM.  Small alphabets (nucleotides): 24 columns per lane, 3 CTAs (12 warps) per SM.
// r02, four-row body, A/B inside one box on the FULL configs[3] (62 375 tasks): 24,3 -> 7 730 GCUPS, 32,2 -> 7 417
// (on 120 genomes 32,2 looked better, 6 864 vs 6 366 -- a wave-quantisation accident: 3 570 tasks are exactly three
// waves of 1 184 warps but 2.01 of 1 776).  A gap model whose window of 24 columns is too wide gets 16 columns
// before the job falls back to the 32-bit kernel.  Proteins: 8 columns (23.5 KB of profile per warp).
static bool w16_pick(uint32_t nsym, long long lipschitz, W32Launch* out) {
  if (nsym == 0 || nsym > 24) return false;
  W32Launch v;
  v.tpb = 128;
  if (nsym <= 8) {
    v.KW = 24;
    v.ctas_sm = 3;
    if (lipschitz > 0 && (32ll * 24 + 4 * 31 + 4 * 16 + 16) * lipschitz > kW16WindowMax) v.KW = 16;
  } else {
    v.KW = 8;
    v.ctas_sm = 2;
  }
  if (const char* e = getenv("TSQ_FORCE_KW16")) {  // developer override for tuning runs: "KW,ctas"
    int kw = 0, ct = 0;
    if (sscanf(e, "%d,%d", &kw, &ct) == 2) {
#define X(KK, TT, MM) if (kw == KK && ct == MM) { v.KW = KK; v.ctas_sm = MM; }
      TSQ_W16_VARIANTS(X)
#undef X
    }
  }
  const size_t sbsz = ((size_t)(nsym + 1) * nsym + 31) & ~(size_t)31;
  v.smem = (sbsz + (size_t)(v.tpb / 32) * nsym * 32 * v.KW) * sizeof(uint32_t) +
           (size_t)(v.tpb / 32) * (2 * 1024 + 16);  // + two 1 KB TMA subject tiles and 2 mbarriers per warp
  if (out) *out = v;
  return true;
}

bool w16_variant(uint32_t nsym, long long lipschitz, W32Launch* out) { return w16_pick(nsym, lipschitz, out); }

long long w16_window(uint32_t nsym, long long lipschitz) {
  W32Launch v;
  if (!w16_pick(nsym, lipschitz, &v)) return 1ll << 40;
  return (32ll * v.KW + 4 * 31 + 4 * 16 + 16) * lipschitz;  // 4 rows per step, RB = 16 steps between re-centrings
}

cudaError_t w16_launch(int grid, const W16Params& p, long long lipschitz, cudaStream_t stream) {
  W32Launch v;
  if (!w16_variant(p.nsym, lipschitz, &v)) return cudaErrorInvalidValue;
#define X(KK, TT, MM)                                                                        \
  if (v.KW == KK && v.ctas_sm == MM) {                                                       \
    auto kern = p.negge2 == 0x00010001u ? wave16_kernel<KK, TT, MM, 0x00010001u>             \
                                        : wave16_kernel<KK, TT, MM, 0u>;                     \
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                         (int)v.smem);                                       \
    if (e != cudaSuccess) return e;                                                          \
    kern<<<grid, TT, v.smem, stream>>>(p);                                                   \
    return cudaGetLastError();                                                               \
  }
  TSQ_W16_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

// ---- interleaved subject database ---------------------------------------------------------------
__global__ void __launch_bounds__(256) subject_db_kernel(const uint8_t* __restrict__ lin, const uint32_t* __restrict__ loff,
                                                         const uint32_t* __restrict__ lens, const uint32_t* __restrict__ goff,
                                                         uint32_t* __restrict__ dbw, uint32_t n, uint32_t lo, uint32_t hi,
                                                         uint32_t scale) {
  const uint32_t g = blockIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t i = g * 32 + lane;                       // this thread's sequence
  const uint32_t words = goff[g + 1] - goff[g];
  const uint32_t rows2 = words / 32;
  const bool in = i < n && i >= lo && i < hi;
  const uint32_t l = in ? lens[i] : 0u;
  const uint32_t odd = l & 1u;
  const uint8_t* sq = lin + (in ? loff[i] : 0u);
  uint32_t* out = dbw + goff[g] + lane;
  for (uint32_t r2 = threadIdx.x >> 5; r2 < rows2; r2 += blockDim.x >> 5) {
    // slots 2*r2 and 2*r2+1 of the right-aligned stream hold residues slot - odd
    const int32_t ra = (int32_t)(2 * r2) - (int32_t)odd, rb = ra + 1;
    uint32_t w = 0;
    if (ra >= 0 && (uint32_t)ra < l) w = (uint32_t)sq[ra] * scale;
    if (rb >= 0 && (uint32_t)rb < l) w |= ((uint32_t)sq[rb] * scale) << 16;
    out[(size_t)r2 * 32] = w;
  }
}

cudaError_t subject_db_launch(const uint8_t* lin, const uint32_t* loff, const uint32_t* lens, const uint32_t* goff,
                              uint32_t* dbw, uint32_t n, uint32_t lo, uint32_t hi, uint32_t scale, cudaStream_t stream) {
  const uint32_t ngroups = (n + 31) / 32;
  if (ngroups == 0) return cudaSuccess;
  subject_db_kernel<<<ngroups, 256, 0, stream>>>(lin, loff, lens, goff, dbw, n, lo, hi, scale);
  return cudaGetLastError();
}

// ---- consensus annotation (Consensus.cpp:80-161) ---------------------------------------
// One CTA per tile of 32 columns.  Pass 1: the 8 warps stream the rows (a warp reads 32 adjacent
// columns of one row: coalesced) into per-column class histograms and first-occurrence rows in shared
// memory.  Pass 2: one thread per column scores every class present against the histogram.
// All weights are 1.0 in the reference, so its double-precision sums are exact integers and integer
// arithmetic here reproduces them bit for bit; ties go to the smallest row, as its strict '>' does.
constexpr int CONS_CLASSES = 24;   // 0..22 = BLOSUM62 rows, 23 = "not a residue" (99 in the reference)

__global__ void __launch_bounds__(256) consensus_kernel(const uint8_t* __restrict__ aln, uint32_t nrows, uint32_t ncols,
                                                        double plurality, const int8_t* __restrict__ blosum,
                                                        const uint8_t* __restrict__ letter_map, uint8_t* __restrict__ out) {
  __shared__ uint32_t hist[32][CONS_CLASSES + 1];
  __shared__ uint32_t first[32][CONS_CLASSES + 1];
  __shared__ int8_t sB[23 * 23];
  __shared__ uint8_t smap[26];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 23 * 23; i += blockDim.x) sB[i] = blosum[i];
  if (threadIdx.x < 26) smap[threadIdx.x] = letter_map[threadIdx.x];
  for (uint32_t i = threadIdx.x; i < 32 * (CONS_CLASSES + 1); i += blockDim.x) {
    (&hist[0][0])[i] = 0;
    (&first[0][0])[i] = 0xffffffffu;
  }
  __syncthreads();
  const uint32_t col = blockIdx.x * 32 + lane;
  if (col < ncols) {
    for (uint32_t r = warp; r < nrows; r += blockDim.x >> 5) {
      const int idx = (int)aln[(size_t)r * ncols + col] - 65;           // Consensus.cpp:99-105
      const uint32_t cls = (idx < 0 || idx > 25) ? 23u : (uint32_t)smap[idx];
      atomicAdd(&hist[lane][cls], 1u);
      atomicMin(&first[lane][cls], r);
    }
  }
  __syncthreads();
  if (warp == 0 && col < ncols) {
    long long best = 0, best_matches = 0;
    uint32_t best_row = 0xffffffffu;
    for (int a = 0; a < CONS_CLASSES; a++) {
      if (hist[lane][a] == 0) continue;
      long long score = 0, matches = 0;
      for (int b = 0; b < CONS_CLASSES; b++) {
        const long long cnt = (long long)hist[lane][b] - (a == b ? 1 : 0);   // every other row
        if (cnt <= 0) continue;
        if (a == 23 && b == 23) { score += cnt; matches += cnt; }             // :118-122
        else if (a == 23 || b == 23) score += -4 * cnt;                       // :123-124
        else {
          const int t = sB[a * 23 + b];                                       // :125-130
          score += t * cnt;
          if (t > 0) matches += cnt;
        }
      }
      const uint32_t fr = first[lane][a];
      // first row wins ties (strict '>' at :138), and row 0 seeds the maximum (:133-137)
      if (best_row == 0xffffffffu || score > best || (score == best && fr < best_row)) {
        best = score; best_matches = matches; best_row = fr;
      }
    }
    uint8_t ch = '?';
    if (best_row != 0xffffffffu && (double)best_matches >= plurality) ch = aln[(size_t)best_row * ncols + col];
    out[col] = ch;
  }
}

cudaError_t consensus_launch(const uint8_t* aln, uint32_t nrows, uint32_t ncols, double plurality, const int8_t* blosum,
                             const uint8_t* letter_map, uint8_t* out, cudaStream_t stream) {
  if (ncols == 0) return cudaSuccess;
  consensus_kernel<<<(ncols + 31) / 32, 256, 0, stream>>>(aln, nrows, ncols, plurality, blosum, letter_map, out);
  return cudaGetLastError();
}

// ---- UPGMA guide tree -----------------------------------------------------------------------------
cudaError_t upgma_launch(const UpgmaParams& p, cudaStream_t stream) {
  if (p.n < 2) return cudaSuccess;
  unsigned int grid = p.n < 148u * 16u ? p.n : 148u * 16u;
  upgma_init_kernel<<<grid, 256, 0, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  // One matrix entry per thread and step at n = 1 000: fewer, fatter threads were measured and are slower (r02,
  // profiles/upgma_threads_r02.txt: 5.8 ms at 1 024 threads, 6.4 at 512, 8.6 at 256 -- the step is the latency of its
  // row re-scans, which more threads hide better).  TSQ_UPGMA_THREADS: the tuning override that sweep used.
  unsigned int threads = (unsigned)UPGMA_THREADS;
  if (const char* ev = getenv("TSQ_UPGMA_THREADS")) {
    const int t = atoi(ev);
    if (t >= 32 && t <= UPGMA_THREADS && (t & 31) == 0) threads = (unsigned)t;
  }
  if (p.n <= UPGMA_SMEM_N) {
    const size_t smem = upgma_smem_bytes(p.n);
    e = cudaFuncSetAttribute(upgma_merge_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    upgma_merge_kernel<true><<<1, threads, smem, stream>>>(p);
  } else {
    upgma_merge_kernel<false><<<1, threads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

// ---- pairwise traceback ------------------------------------------------------------------------------
cudaError_t traceback_launch(const TbParams& p, cudaStream_t stream) {
  const uint32_t longest_diag = (p.m < p.n ? p.m : p.n) + 1;
  if (longest_diag <= 2u * TB_THREADS) {   // a CTA covers the diagonal in <= 2 sweeps: the CTA barrier is cheaper
    traceback_kernel<1><<<1, TB_THREADS, 0, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(TB_CLUSTER);
  cfg.blockDim = dim3(TB_THREADS);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = TB_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, traceback_kernel<TB_CLUSTER>, p);
}

// ---- progressive alignment (msa.cuh) ---------------------------------------------------------
cudaError_t msa_leaf_launch(const MsaLeaf* d_leaves, uint32_t n, uint32_t nsym, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  msa_leaf_kernel<<<n < 65535u ? n : 65535u, 128, 0, stream>>>(d_leaves, n, nsym);
  return cudaGetLastError();
}

cudaError_t msa_merge_launch(const MsaTask* d_tasks, uint32_t count, uint32_t threads, uint32_t smem_bytes, const MsaConst& k,
                             cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  if (threads < 32 || threads > 512 || (threads & 31u) || smem_bytes > 227u * 1024u) return cudaErrorInvalidValue;
  // one thread per 4 x 4 tile of the longest anti-diagonal of tiles: 512 threads cover 2 048 columns at once (longer
  // ones loop), with 128 registers each -- the tile's 16 column scores and its edges live in registers, no spills
  auto kern = msa_merge_kernel<512>;
  if (smem_bytes > 48u * 1024u) {   // above the default limit the kernel must opt in (per device, cheap to repeat)
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  kern<<<count, threads, smem_bytes, stream>>>(d_tasks, k, smem_bytes);
  return cudaGetLastError();
}

cudaError_t msa_rows_launch(const MsaRows& p, cudaStream_t stream) {
  if (p.n == 0) return cudaSuccess;
  msa_rows_kernel<<<p.n < 65535u ? p.n : 65535u, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// ---- finalize: empties, un-sort, fp64 distances -----------------------------------------
// One CTA per sorted row i; threads stride over j > i.  Distances follow the oracle's
// tsq_oracle_distance(): two separately rounded IEEE operations (div, sub), no contraction.
// The logarithm of the Kimura correction is part of the spec (oracle/gotoh_oracle.c: tsq_oracle_ln): the same
// sequence of separately rounded IEEE double operations, so that the distances equal the oracle's bit for bit.
__device__ __forceinline__ double ln_spec(double x) {
  long long bits = __double_as_longlong(x);
  int e = (int)((bits >> 52) & 0x7ff) - 1022;                 // x = m 2^e, m in [1/2, 1)
  double m = __longlong_as_double((bits & 0x000fffffffffffffll) | 0x3fe0000000000000ll);
  if (m < 0.70710678118654752440) {
    m = __dmul_rn(m, 2.0);
    e -= 1;
  }
  const double z = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0));
  const double w = __dmul_rn(z, z);
  double p = 1.0 / 23.0;
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 21.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 19.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 17.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 15.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 13.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 11.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 9.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 7.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 5.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0 / 3.0);
  p = __dadd_rn(__dmul_rn(p, w), 1.0);
  const double lnm = __dmul_rn(__dmul_rn(2.0, z), p);
  return __dadd_rn(__dmul_rn((double)e, 0.693147180559945309417), lnm);
}

__global__ void __launch_bounds__(256) finalize_kernel(const __grid_constant__ FinalizeParams p) {
  const unsigned long long n = p.n;
  for (unsigned long long i = (unsigned long long)p.row_begin + blockIdx.x; i < p.row_end && i + 1 < n; i += gridDim.x) {
    const uint32_t li = p.lens[i];
    const uint32_t oi = p.perm_identity ? (uint32_t)i : p.perm[i];
    const int32_t si = p.self[i];
    const unsigned long long rowbase = tri_index(i, i + 1, n);
    for (unsigned long long j = i + 1 + threadIdx.x; j < n; j += blockDim.x) {
      const unsigned long long sidx = rowbase + (j - i - 1);
      int32_t s;
      if (li == 0) {
        const uint32_t lj = p.lens[j];
        s = lj == 0 ? 0 : -(p.go + (int32_t)lj * p.ge);
      } else {
        s = p.sorted[sidx];
      }
      int32_t nid = 0;
      if (p.idshift && li != 0) {   // key = score * 2^idshift + identities, identities < 2^idshift
        nid = s & (int32_t)((1u << p.idshift) - 1u);
        s >>= p.idshift;            // arithmetic shift = floor division: exact for negative scores
      }
      unsigned long long oidx = sidx;
      if (!p.perm_identity) {
        const uint32_t oj = p.perm[j];
        const unsigned long long a = oi < oj ? oi : oj, b = oi < oj ? oj : oi;
        oidx = tri_index(a, b, n);
      }
      if (p.out_scores != p.sorted || li == 0 || p.idshift) p.out_scores[oidx] = s;
      if (p.out_nid) p.out_nid[oidx] = nid;
      if (p.out_dist) {
        const int32_t sj = p.self[j];
        const int32_t mn = si < sj ? si : sj;
        double d = 1.0;
        if (p.idshift) {            // ClustalW pairwise distance: 1 - identities / shorter length
          const uint32_t lj = p.lens[j];
          const uint32_t ml = li < lj ? li : lj;
          if (ml > 0) d = __dsub_rn(1.0, __ddiv_rn((double)nid, (double)ml));
          if (p.kimura) {           // -ln(1 - D - D^2/5), operation for operation as tsq_oracle_kimura
            if (ml > 0 && d < 0.75) {
              const double u = __ddiv_rn(__dmul_rn(d, d), 5.0);
              d = __dsub_rn(0.0, ln_spec(__dsub_rn(__dsub_rn(1.0, d), u)));
            } else {
              atomicAdd(p.kimura_oob, 1);   // the uncorrected value stays in place; the host turns the count into an error
            }
          }
        } else if (mn > 0) {
          d = __dsub_rn(1.0, __ddiv_rn((double)s, (double)mn));
        }
        p.out_dist[oidx] = d;
      }
    }
  }
}

cudaError_t finalize_launch(const FinalizeParams& p, cudaStream_t stream) {
  if (p.n < 2 || p.row_end <= p.row_begin) return cudaSuccess;
  unsigned int grid = p.row_end - p.row_begin;
  if (grid > 148u * 64u) grid = 148u * 64u;
  finalize_kernel<<<grid, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// ---- int32 -> int16 scores -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) narrow_scores_kernel(const int32_t* __restrict__ src, int16_t* __restrict__ dst,
                                                            unsigned long long count) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    dst[i] = (int16_t)src[i];
}

cudaError_t narrow_scores_launch(const int32_t* src, int16_t* dst, unsigned long long count, cudaStream_t stream) {
  if (count == 0) return cudaSuccess;
  unsigned long long blocks = (count + 255) / 256;
  if (blocks > 148ull * 32ull) blocks = 148ull * 32ull;
  narrow_scores_kernel<<<(unsigned int)blocks, 256, 0, stream>>>(src, dst, count);
  return cudaGetLastError();
}

// ---- DPX issue-rate probe -----------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1) dpx_probe_kernel(uint32_t* out, uint32_t k1, uint32_t k2,
                                                            long long* cyc) {
  constexpr int NCH = 8, ITER = 2048;
  uint32_t a[NCH], b[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) {
    a[i] = threadIdx.x * 7 + i + k1;
    b[i] = threadIdx.x * 3 + i * 5 + k2;
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      a[i] = __viaddmax_u16x2(a[i], k1, b[i]);
      b[i] = __vimax3_u16x2(b[i], k2, a[i]);
    }
  }
  const long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) r ^= a[i] ^ b[i];
  if (r == 0x12345678u) out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

cudaError_t dpx_probe(int sms, double* ops_per_clk_per_sm, double* sm_mhz, cudaStream_t stream) {
  uint32_t* d_out = nullptr;
  long long* d_cyc = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&d_out, 1024 * sizeof(uint32_t))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_cyc, sizeof(long long) * sms)) != cudaSuccess) {
    cudaFree(d_out);
    return e;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  dpx_probe_kernel<<<sms, 1024, 0, stream>>>(d_out, 3, 5, d_cyc);  // warm-up
  cudaEventRecord(e0, stream);
  dpx_probe_kernel<<<sms, 1024, 0, stream>>>(d_out, 3, 5, d_cyc);
  cudaEventRecord(e1, stream);
  e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) {
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long* h = new long long[sms];
    e = cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    long long mx = 0;
    for (int i = 0; i < sms; i++) {
      avg += (double)h[i];
      if (h[i] > mx) mx = h[i];
    }
    avg /= sms;
    delete[] h;
    const double ops = 1024.0 * 8 * 2048 * 2;
    if (ops_per_clk_per_sm) *ops_per_clk_per_sm = ops / avg;
    if (sm_mhz) *sm_mhz = ms > 0 ? (double)mx / (ms * 1e3) : 0.0;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  cudaFree(d_cyc);
  return e;
}

// ---- issue-rate and instruction-mix probes (the denominators of bench.py's two-pipe roofline) -----------
// MODE 0: the issue ceiling -- independent 32-bit adds and xors (full-rate integer work on both pipes).
// MODE 1: the inner loop's own mix per packed cell, dependency-free: 1 LDS.32 (conflict-free), t = h + s,
//         h = VIMNMX3.U16x2(t, E, F), hg = h - goe, E/F = VIADDMNMX.U16x2(., imm, hg) -- gotoh16.cuh's cell.
template <int MODE>
__global__ void __launch_bounds__(512, 1) mix_probe_kernel(uint32_t* out, uint32_t k1, uint32_t k2, long long* cyc) {
  constexpr int NCH = 8, ITER = 4096;
  __shared__ uint32_t sm[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = i & 15;
  __syncthreads();
  uint32_t a[NCH], b[NCH], h[NCH];
#pragma unroll
  for (int i = 0; i < NCH; i++) {
    a[i] = 0x40004000u + threadIdx.x * 7 + i;
    b[i] = 0x40004000u + threadIdx.x * 3 + i * 5;
    h[i] = 0x40004000u + i;
  }
  const uint32_t* sp = sm + (threadIdx.x & 31);
  uint32_t goe_r;
  asm volatile("mov.u32 %0, %1;" : "=r"(goe_r) : "r"(k1));
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; it++) {
    const uint32_t* row = sp + (it & 7) * 64;
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      if (MODE == 0) {
        asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(k1));
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(k2));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(h[i]) : "r"(k2));
      } else {
        const uint32_t s = row[i * 33];
        const uint32_t t = h[i] + s;
        const uint32_t hh = __vimax3_u16x2(t, a[i], b[i]);
        const uint32_t hg = hh - goe_r;
        a[i] = __viaddmax_u16x2(a[i], 0x00010001u, hg);
        b[i] = __viaddmax_u16x2(b[i], 0x00010001u, hg);
        h[i] = hh;
      }
    }
  }
  const long long t1 = clock64();
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < NCH; i++) r ^= a[i] ^ b[i] ^ h[i];
  if (r == 0x12345678u) out[threadIdx.x] = r;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

cudaError_t mix_probe(int sms, double* issue_per_clk_per_sm, double* mix_cells_per_clk_per_sm, cudaStream_t stream) {
  uint32_t* d_out = nullptr;
  long long* d_cyc = nullptr;
  cudaError_t e;
  if ((e = cudaMalloc(&d_out, 1024 * sizeof(uint32_t))) != cudaSuccess) return e;
  if ((e = cudaMalloc(&d_cyc, sizeof(long long) * sms)) != cudaSuccess) {
    cudaFree(d_out);
    return e;
  }
  long long* h = new long long[sms];
  auto avg_cycles = [&](int mode) -> double {
    for (int rep = 0; rep < 2; rep++) {   // the first launch warms up
      if (mode == 0) mix_probe_kernel<0><<<sms, 512, 0, stream>>>(d_out, 0x000a000a, 0x00010001, d_cyc);
      else mix_probe_kernel<1><<<sms, 512, 0, stream>>>(d_out, 0x000a000a, 0x00010001, d_cyc);
    }
    if ((e = cudaStreamSynchronize(stream)) != cudaSuccess) return 0.0;
    if ((e = cudaMemcpy(h, d_cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost)) != cudaSuccess) return 0.0;
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)h[i];
    return avg / sms;
  };
  const double c0 = avg_cycles(0);
  const double c1 = e == cudaSuccess ? avg_cycles(1) : 0.0;
  if (e == cudaSuccess && c0 > 0 && c1 > 0) {
    if (issue_per_clk_per_sm) *issue_per_clk_per_sm = 512.0 * 8 * 4096 * 3 / c0;
    if (mix_cells_per_clk_per_sm) *mix_cells_per_clk_per_sm = 512.0 * 8 * 4096 / c1;
  }
  delete[] h;
  cudaFree(d_out);
  cudaFree(d_cyc);
  return e;
}

}  // namespace tsq
