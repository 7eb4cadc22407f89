// tsq_device.h -- host-callable launchers of the sm_100a kernels (internal, C++).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gotoh16.cuh"
#include "gotoh32.cuh"
#include "wave32.cuh"
#include "wave16.cuh"
#include "upgma.cuh"
#include "traceback.cuh"
#include "msa.cuh"

namespace tsq {

// Strip widths the packed 16-bit kernel is instantiated for.
constexpr int kNumStripWidths = 9;
extern const int kStripWidths[kNumStripWidths];

struct G16Launch {
  int K;         // strip width
  int stride;    // profile row stride in words (== 1 mod 32): the database pre-scales letters by it
  int tpb;       // threads per CTA
  int ctas_sm;   // resident CTAs per SM the variant is compiled for
  size_t smem;   // dynamic shared memory per CTA for a given nsym
};

// Static description of the variant with strip width K (nullptr-safe: returns false).
bool g16_variant(int K, uint32_t nsym, G16Launch* out);
// Launch the packed 16-bit kernel: grid CTAs of the variant's size on `stream`; scratch_window (may be null): L2
// access policy of the strip-boundary scratch for this launch.
cudaError_t g16_launch(int K, int grid, const G16Params& p, cudaStream_t stream, const cudaAccessPolicyWindow* scratch_window);

// 32-bit inter-task kernel (gotoh32.cuh): same strip widths / launch shapes as the packed one.
cudaError_t g32_launch(int K, int grid, const G32Params& p, cudaStream_t stream);

// 32-bit wavefront kernel: columns per lane (KW) is chosen from the alphabet size so that the
// per-warp profile fits shared memory; *warps_per_cta / *ctas_sm describe the launch shape.
struct W32Launch {
  int KW, tpb, ctas_sm;
  size_t smem;
};
bool w32_variant(uint32_t nsym, W32Launch* out);
cudaError_t w32_launch(int grid, const W32Params& p, cudaStream_t stream);

// Packed wavefront kernel (wave16.cuh).  w16_window() is the span D (in score units) of the cells a
// warp holds at one time for a per-step Lipschitz bound L; the kernel is exact while D <= 30000.
// The variant depends on the per-step Lipschitz bound too: the widest column block whose window still fits.
constexpr long long kW16WindowMax = 30000;
bool w16_variant(uint32_t nsym, long long lipschitz, W32Launch* out);
long long w16_window(uint32_t nsym, long long lipschitz);
cudaError_t w16_launch(int grid, const W16Params& p, long long lipschitz, cudaStream_t stream);

// UPGMA guide tree: init (one CTA per row) + one persistent CTA for the n-1 merges.
cudaError_t upgma_launch(const UpgmaParams& p, cudaStream_t stream);

// One pair with its path (traceback.cuh): a single CTA sweeps the anti-diagonals.
cudaError_t traceback_launch(const TbParams& p, cudaStream_t stream);

// Progressive alignment along the guide tree (msa.cuh): leaf profiles, one launch per batch of
// independent merges (one CTA each, `threads` per CTA), final rows.
cudaError_t msa_leaf_launch(const MsaLeaf* d_leaves, uint32_t n, uint32_t nsym, cudaStream_t stream);
cudaError_t msa_merge_launch(const MsaTask* d_tasks, uint32_t count, uint32_t threads, uint32_t smem_bytes, const MsaConst& k,
                             cudaStream_t stream);
cudaError_t msa_rows_launch(const MsaRows& p, cudaStream_t stream);

// Builds the 32-way interleaved subject database of the packed kernel from the linear residues:
// per residue the 16-bit byte offset of its profile row, two rows per word, right-aligned to an
// even row count (gotoh16.cuh).  One CTA per group of 32 sequences; every word of the group is written.
cudaError_t subject_db_launch(const uint8_t* lin, const uint32_t* loff, const uint32_t* lens, const uint32_t* goff,
                              uint32_t* dbw, uint32_t n, uint32_t lo, uint32_t hi, uint32_t scale, cudaStream_t stream);

// Consensus annotation: aln = nrows x ncols characters, row-major; out = ncols chars.
cudaError_t consensus_launch(const uint8_t* aln, uint32_t nrows, uint32_t ncols, double plurality,
                             const int8_t* blosum /*23x23*/, const uint8_t* letter_map /*26*/, uint8_t* out,
                             cudaStream_t stream);

// All three packed-triangle pointers are indexed by ABSOLUTE packed index: a context that holds only its
// slab passes addresses moved back by the slab's first index.  out_scores / out_nid / out_dist may point into
// ANOTHER device's memory (peer access over NVLink): the children of a multi-device context un-sort their
// rows straight into device 0's matrix -- finalize and gather in one kernel, no staging copy.
struct FinalizeParams {
  const int32_t* sorted;      // packed triangle, sorted order
  const uint32_t* lens;       // sorted lengths
  const uint32_t* perm;       // sorted index -> original index
  const int32_t* self;        // self scores, sorted order
  int32_t* out_scores;        // packed triangle, original order (may alias sorted if perm_identity)
  double* out_dist;           // packed triangle, original order, or nullptr
  uint32_t n;
  uint32_t row_begin, row_end;   // sorted rows this launch covers (a rank's own rows; 0, n for everything)
  int32_t go, ge;
  uint32_t perm_identity;        // perm is the identity and there are no empty sequences
  uint32_t idshift;           // identity mode: sorted[] holds score * 2^idshift + identities; 0 = off
  int32_t* out_nid;           // identity mode: identities, packed triangle, original order
  uint32_t kimura;            // identity mode: distances Kimura-corrected, -ln(1 - D - D^2/5) for D < 0.75
  int* kimura_oob;            // counts the pairs with D >= 0.75 (the formula does not apply: the host reports an error)
};
cudaError_t finalize_launch(const FinalizeParams& p, cudaStream_t stream);

// int32 scores -> int16 (TSQ_FLAG_SCORES_I16; the host has checked that every score fits)
cudaError_t narrow_scores_launch(const int32_t* src, int16_t* dst, unsigned long long count, cudaStream_t stream);

// DPX issue-rate probe: thread-level VIADDMNMX.U16x2 + VIMNMX3.U16x2 results / clk / SM.
cudaError_t dpx_probe(int device_sms, double* ops_per_clk_per_sm, double* sm_mhz, cudaStream_t stream);

// Issue ceiling (independent full-rate integer instructions, thread-level per clk per SM) and the packed cells
// per clk per SM the inner loop's own instruction mix reaches in isolation (dependency-free).
cudaError_t mix_probe(int device_sms, double* issue_per_clk_per_sm, double* mix_cells_per_clk_per_sm, cudaStream_t stream);

}  // namespace tsq
