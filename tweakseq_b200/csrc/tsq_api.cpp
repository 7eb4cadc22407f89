// tsq_api.cpp -- the C ABI of libtsqb200.so (include/tsq_b200.h): context, host-side
// encode / length sort / packing, work partitioning, kernel orchestration, result access.
//
// Replaces the process boundary tweakseq/UI/SeqEditMainWin.cpp:1654-1660 for the pairwise
// distance stage; matrix and alphabet from tweakseq/Core/Annotations/Consensus.cpp:34-69.
// No CPU compute path exists here: every score comes out of the sm_100a kernels.
#include <algorithm>
#include <charconv>
#include <chrono>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <fstream>
#include <mutex>
#include <numeric>
#include <string>
#include <condition_variable>
#include <functional>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cuda.h>   // types of the one driver entry point resolved at run time (no link dependency)
#include <cuda_runtime.h>

#include "../../include/tsq_b200.h"
#include "tsq_device.h"
#include "encode_simd.h"
#include "msa_host.h"

namespace {

// Rows/columns A R N D C Q E G H I L K M F P S T W Y V B Z X
// (values and order: tweakseq/Core/Annotations/Consensus.cpp:34-59).
const int8_t kBlosum62[23 * 23] = {
     4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0,-2,-1, 0,
    -1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3,-1, 0,-1,
    -2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3, 3, 0,-1,
    -2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3, 4, 1,-1,
     0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1,-3,-3,-2,
    -1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2, 0, 3,-1,
    -1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,
     0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3,-1,-2,-1,
    -2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3, 0, 0,-1,
    -1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3,-3,-3,-1,
    -1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1,-4,-3,-1,
    -1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2, 0, 1,-1,
    -1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1,-3,-1,-1,
    -2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1,-3,-3,-1,
    -1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2,-2,-1,-2,
     1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2, 0, 0, 0,
     0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0,-1,-1, 0,
    -3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3,-4,-3,-2,
    -2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1,-3,-2,-1,
     0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4,-3,-2,-1,
    -2,-1, 3, 4,-3, 0, 1,-1, 0,-3,-4, 0,-3,-3,-2, 0,-1,-4,-3,-3, 4, 1,-1,
    -1, 0, 0, 1,-3, 3, 4,-2, 0,-3,-3, 1,-1,-3,-1, 0,-1,-3,-2,-2, 1, 4,-1,
     0,-1,-1,-1,-2,-1,-1,-1,-1,-1,-1,-1,-1,-1,-2, 0, 0,-2,-1,-1,-1,-1,-1,
};
// 'A'..'Z' -> matrix row (Consensus.cpp:61-69; J, O, U, X -> X)
const uint8_t kProteinIndex[26] = {0,  20, 4,  3,  6,  13, 7,  8,  9,  22, 11, 10, 12,
                                   2,  22, 14, 5,  1,  15, 16, 22, 19, 17, 22, 18, 21};
// A C G T N (the reference has no nucleotide matrix)
const int8_t kDna[5 * 5] = {5, -4, -4, -4, -2, -4, 5, -4, -4, -2, -4, -4, 5,
                            -4, -2, -4, -4, -4, 5, -2, -2, -2, -2, -2, -1};

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// Process-wide caches of released device and page-locked host blocks.  The plugin call (tsq_run_fasta) builds a
// context per alignment job; cudaMalloc / cudaHostAlloc / cudaFree of its dozen buffers cost more than the kernels
// of a 100-sequence job, so released blocks are kept (up to a bound) and handed to the next context that asks for
// a similar size on the same device.  Never freed at exit: the CUDA runtime may be gone by then.
class BlockCache {
 public:
  static BlockCache& get() {
    static BlockCache* inst = new BlockCache();   // intentionally leaked
    return *inst;
  }
  // device < 0: page-locked host memory (portable: any device may copy into it)
  cudaError_t take(int device, size_t bytes, void** out) {
    {
      std::lock_guard<std::mutex> g(m_);
      size_t best = free_.size();
      for (size_t i = 0; i < free_.size(); i++)
        if (free_[i].device == device && free_[i].bytes >= bytes && free_[i].bytes <= 2 * bytes + 4096 &&
            (best == free_.size() || free_[i].bytes < free_[best].bytes))
          best = i;
      if (best != free_.size()) {
        *out = free_[best].p;
        live_[*out] = free_[best].bytes;
        cached_ -= free_[best].bytes;
        free_.erase(free_.begin() + (long)best);
        return cudaSuccess;
      }
    }
    cudaError_t e = device < 0 ? cudaHostAlloc(out, bytes, cudaHostAllocPortable) : cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {   // give the cache back and try once more
      cudaGetLastError();
      trim(device);
      e = device < 0 ? cudaHostAlloc(out, bytes, cudaHostAllocPortable) : cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) {
      std::lock_guard<std::mutex> g(m_);
      live_[*out] = bytes;
    }
    return e;
  }
  void give(int device, void* p) {
    if (!p) return;
    size_t bytes = 0;
    {
      std::lock_guard<std::mutex> g(m_);
      auto it = live_.find(p);
      if (it != live_.end()) {
        bytes = it->second;
        live_.erase(it);
      }
      if (bytes > 0 && bytes <= kMaxBlock && cached_ + bytes <= kMaxCached && free_.size() < 64) {
        free_.push_back({p, bytes, device});
        cached_ += bytes;
        return;
      }
    }
    if (device < 0) cudaFreeHost(p);
    else cudaFree(p);
  }

 private:
  struct Entry {
    void* p;
    size_t bytes;
    int device;
  };
  static constexpr size_t kMaxBlock = (size_t)256 << 20, kMaxCached = (size_t)1 << 30;
  void trim(int device) {
    std::vector<Entry> drop;
    {
      std::lock_guard<std::mutex> g(m_);
      for (size_t i = free_.size(); i-- > 0;)
        if (free_[i].device == device) {
          drop.push_back(free_[i]);
          cached_ -= free_[i].bytes;
          free_.erase(free_.begin() + (long)i);
        }
    }
    for (const Entry& e : drop) {
      if (e.device < 0) cudaFreeHost(e.p);
      else cudaFree(e.p);
    }
  }
  std::mutex m_;
  std::vector<Entry> free_;
  std::unordered_map<void*, size_t> live_;
  size_t cached_ = 0;
};

int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) cudaGetLastError();
  return d;
}

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  int dev = 0;     // the device the block lives on
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    release();
    dev = current_device();
    cudaError_t e = BlockCache::get().take(dev, std::max<size_t>(n, 1) * sizeof(T), (void**)&p);
    if (e == cudaSuccess) cap = n;
    else p = nullptr;
    return e;
  }
  // scratch that kernels read a few slack rows beyond what they wrote (prefetch): zero it once
  cudaError_t reserve_zeroed(size_t n) {
    if (n <= cap) return cudaSuccess;
    cudaError_t e = reserve(n);
    if (e == cudaSuccess) e = cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T));
    return e;
  }
  void release() {
    if (p) {
      // what cudaFree did implicitly: nothing in flight on the block's device may still use it when the cache
      // hands it to the next taker
      const int cur = current_device();
      if (cur != dev) cudaSetDevice(dev);
      if (cudaDeviceSynchronize() != cudaSuccess) cudaGetLastError();
      if (cur != dev) cudaSetDevice(cur);
      BlockCache::get().give(dev, p);
    }
    p = nullptr;
    cap = 0;
  }
};

template <typename T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    release();
    // portable: every device of a multi-device context copies its slab straight into this buffer
    cudaError_t e = BlockCache::get().take(-1, std::max<size_t>(n, 1) * sizeof(T), (void**)&p);
    if (e == cudaSuccess) cap = n;
    else p = nullptr;
    return e;
  }
  void release() {
    if (p) BlockCache::get().give(-1, p);
    p = nullptr;
    cap = 0;
  }
};

// A typed pointer into tsq_ctx::d_blob (one allocation, one host -> device copy per upload).
template <typename T>
struct DevView {
  T* p = nullptr;
};

}  // namespace

struct tsq_ctx;

// One persistent host thread per device of a multi-device context: upload (sort, plan, pack, H2D) and compute
// (a dozen runtime calls per device when the results are streamed) run side by side instead of one device after
// the other; a job is handed over with one condition-variable round trip (~10 us), not a thread start.
struct KidPool {
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  std::function<int(tsq_ctx*)> job;
  std::vector<int> rc;
  uint64_t generation = 0;
  size_t pending = 0;
  bool quit = false;

  explicit KidPool(std::vector<tsq_ctx*>& kids) : rc(kids.size(), TSQ_OK) {
    for (size_t r = 0; r < kids.size(); r++)
      threads.emplace_back([this, r, k = kids[r]]() {
        uint64_t seen = 0;
        for (;;) {
          std::function<int(tsq_ctx*)> f;
          {
            std::unique_lock<std::mutex> lk(m);
            cv_go.wait(lk, [&] { return quit || generation != seen; });
            if (quit) return;
            seen = generation;
            f = job;
          }
          const int v = f(k);
          {
            std::lock_guard<std::mutex> lk(m);
            rc[r] = v;
            if (--pending == 0) cv_done.notify_one();
          }
        }
      });
  }
  ~KidPool() {
    {
      std::lock_guard<std::mutex> lk(m);
      quit = true;
    }
    cv_go.notify_all();
    for (auto& t : threads) t.join();
  }
  void run(std::function<int(tsq_ctx*)> f) {
    std::unique_lock<std::mutex> lk(m);
    job = std::move(f);
    pending = threads.size();
    generation++;
    cv_go.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
};

struct tsq_ctx {
  tsq_params prm{};
  int nsym = 23;
  std::vector<int8_t> matrix;  // nsym x nsym
  int go = 11, ge = 1;
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;    // around the score kernels of tsq_compute
  cudaEvent_t tev0 = nullptr, tev1 = nullptr;  // around the guide-tree kernels
  std::string err;

  // packed-16 arithmetic constants
  int smin = 0, smax = 0, delta = 0;
  uint32_t max_len16 = 0;  // longest sequence the packed kernel may see

  // host copies (sorted order unless noted)
  uint32_t n = 0;
  std::vector<std::vector<uint8_t>> enc;  // submitted order
  std::vector<uint32_t> perm;             // sorted -> submitted
  std::vector<uint32_t> lens;             // sorted
  std::vector<uint32_t> loff;             // sorted
  // Pinned staging blob, copied to d_blob by ONE cudaMemcpyAsync: the sorted, concatenated residues
  // (16-byte aligned starts) first, then the small tables (group/sequence offsets, lengths,
  // permutation, self scores, biased score table, task prefix), each 16-byte aligned.
  PinnedBuf<uint8_t> h_blob;
  size_t lin_size = 0;                    // residue part of the blob
  size_t blob_size = 0;                   // bytes in use
  std::vector<int32_t> self_input;        // self scores, submitted order (computed while encoding)
  tsq::EncodeTables enct = {};            // input byte -> symbol and S(x, x) of its symbol (encode_simd.h)
  size_t dbw_size = 0;                    // interleaved subject database: words (built on the device)
  uint32_t db_scale = 0;                  // letter -> profile-row byte offset factor (STRIDE(K) * 4)
  std::vector<uint32_t> goff;             // group offsets
  bool perm_identity = true;             // the length sort left the submitted order unchanged (and no empties)
  uint32_t lo = 0, hi = 0;  // sorted range eligible for the packed 16-bit kernel
  uint32_t row_a = 0, row_b = 0;  // this partition's sorted rows
  uint64_t part_begin = 0, part_end = 0;
  std::vector<uint32_t> first_row;        // first sorted row of every rank of the partition (world + 1 entries)
  std::vector<unsigned long long> task_prefix;
  std::vector<uint2> pairs32;             // tasks of the 32-bit wavefront kernel (sorted indices)
  std::vector<uint4> tasks16w;            // tasks of the packed wavefront kernel (i1, i2, j, 0)
  bool use_w16 = false;
  uint32_t inter_max = 0;                 // longest sequence the inter-task kernel of this job takes
  bool use_g32 = false;                   // [lo, hi) runs on the 32-bit inter-task kernel instead of the packed one
  uint32_t idshift = 0;                   // identity mode: keys are score * 2^idshift + identities
  int go_k = 0, ge_k = 0;                 // gap penalties in key units (scaled by 2^idshift in identity mode)
  std::vector<int32_t> smat_k;            // (nsym+1) x nsym score table in key units
  uint32_t q_begin = 0, q_end = 0;
  int K = 0;
  uint64_t cells16 = 0, cells32 = 0, pairs_part = 0;

  // device
  DevBuf<uint8_t> d_blob;
  DevView<uint8_t> d_lin;                                   // views into d_blob
  DevView<uint32_t> d_goff, d_loff, d_lens, d_perm, d_sbias;
  DevView<int32_t> d_self;
  DevView<unsigned long long> d_prefix;
  DevBuf<uint32_t> d_dbw;
  DevBuf<int32_t> d_sorted, d_scores, d_nid;
  DevBuf<double> d_dist;
  DevBuf<unsigned long long> d_counter;
  DevBuf<uint2> d_bnd;
  DevBuf<double> d_treeD, d_treemin, d_treeh;
  DevBuf<uint32_t> d_treeu;               // rowarg, active, csize, node, rescan
  DevBuf<tsq_merge> d_merges;
  std::vector<tsq_merge> merges;
  bool have_tree = false;
  double tree_ms = 0;
  // progressive alignment along that tree (tsq_msa)
  std::vector<uint8_t> msa_rows;          // n x msa_cols characters, submitted order
  std::vector<uint32_t> msa_order;        // leaves left to right
  uint32_t msa_cols = 0;
  bool have_msa = false;
  double msa_ms = 0;
  volatile int* msa_cancel = nullptr;     // set by tsq_run_fasta around its tsq_msa: polled before every launch
  DevBuf<uint2> d_pairs32;
  DevBuf<uint4> d_tasks16w;
  DevBuf<uint2> d_bnd16w;
  DevBuf<int2> d_bnd32;
  DevBuf<int32_t> d_smat;
  // host results
  PinnedBuf<int32_t> h_scores, h_nid;
  PinnedBuf<int16_t> h_scores16;           // TSQ_FLAG_SCORES_I16: the host result instead of h_scores
  DevBuf<int16_t> d_scores16;              // ... and its device source (same extent as the int32 buffer it narrows)
  PinnedBuf<double> h_dist;

  // Cancel flag the kernels poll at every task fetch.  It lives in DEVICE memory and is written by a
  // copy on a side stream while the kernels run (polling a host-mapped flag cost ~0.5 ms per read on
  // this platform: C2 went from 5.3 to 8.1 ms).
  int* d_cancel = nullptr;
  int* h_one = nullptr;      // pinned source of that copy
  cudaStream_t cancel_stream = nullptr;

  bool have_seqs = false, uploaded = false, computed = false, finalized = false, downloaded = false;
  tsq_stats st{};

  // ---- partitioned jobs (part_world > 1, or the children of a multi-device context) -------------------
  // slab_mode: the length sort is the identity (and no identity keys), so this rank's slab of the sorted
  // triangle IS a contiguous piece of the final one: the rank finalizes and downloads it itself
  // (tsq_results_sharded).  full_sorted: d_sorted spans the whole triangle (single rank, or the root of a
  // gather); otherwise it holds [part_begin, part_end) only and kernels get a pointer biased by part_begin.
  bool slab_mode = false;
  bool full_sorted = true;
  // ---- multi-device (tsq_params.n_devices > 1): a LEADER owns one child context per device ------------
  std::vector<tsq_ctx*> kids;      // leader only
  KidPool* pool = nullptr;         // leader only: one persistent host thread per device
  tsq_ctx* leader = nullptr;       // child only
  const std::vector<std::vector<uint8_t>>* encp = nullptr;   // encoded sequences: own `enc`, or the leader's
  const std::vector<int32_t>* selfp = nullptr;                // self scores: own `self_input`, or the leader's
  cudaEvent_t fin_ev = nullptr;    // recorded behind this context's finalize (cross-device stream waits)
  DevBuf<double> d_dist_full;      // child 0 of a leader in slab mode: all slabs of the distance matrix (guide tree)
  // ---- streamed results (tsq_run, tsq_stream_results): consecutive row ranges of the result leave for the host on a
  // side stream while the packed kernel is still running -- it counts finished tasks per range in d_done and the side
  // stream waits on the counters; or, without stream memory operations, behind one launch per range (SURVEY.md
  // section 8e: "slabs overlapped with remaining compute on a side stream")
  bool ev0_set = false;             // this tsq_compute has recorded its kernel-start event (right ahead of its first launch)
  bool stream_out = false;          // asked for
  bool streamed = false;            // the last tsq_compute did it: tsq_download has nothing left to copy
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> chunk_ev;
  unsigned long long* h_starts = nullptr;   // pinned: first task of every chunk (the kernel's cursor is set from it)
  uint64_t streamed_bytes = 0;
  DevBuf<unsigned int> d_done;             // per row range: finished tasks of the running packed-kernel launch
  // ---- caller-owned host result buffers (tsq_set_result_buffers) ---------------------------------------
  int32_t* ext_scores = nullptr;
  double* ext_dist = nullptr;
  uint64_t ext_count = 0;
  std::vector<void*> ext_registered;   // page ranges this context has page-locked
  // partial pages at the ends of a slab in caller memory: copied into this pinned bounce block asynchronously and
  // moved to their place by the host after the synchronize (copy_out / apply_fixups)
  struct Fixup { void* dst; size_t off, bytes; };
  PinnedBuf<uint8_t> bounce;
  size_t bounce_used = 0;
  std::vector<Fixup> fixups;
  // Device-side fault word next to the cancel flag (d_cancel[1]): a kernel that gives up on a TMA barrier
  // sets it; tsq_synchronize turns it into TSQ_ERR_CUDA.
  int* h_fault = nullptr;          // pinned landing place of its copy
};

namespace {

int fail(tsq_ctx* c, int code, const char* fmt, ...) {
  if (c) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    c->err = buf;
  }
  return code;
}

#define TSQ_CUDA(c, call)                                                                  \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      cudaGetLastError(); /* not sticky: the next launch check must not report this one */ \
      return fail((c), e_ == cudaErrorMemoryAllocation ? TSQ_ERR_NOMEM : TSQ_ERR_CUDA,     \
                  "%s failed: %s", #call, cudaGetErrorString(e_));                         \
    }                                                                                      \
  } while (0)

inline uint64_t tri(uint64_t i, uint64_t j, uint64_t n) { return i * n - i * (i + 1) / 2 + (j - i - 1); }
inline uint64_t score_bytes(const tsq_ctx* c) { return (c->prm.flags & TSQ_FLAG_SCORES_I16) ? 2ull : 4ull; }   // per pair, on the way out

// Kernels index the sorted-order score buffer by ABSOLUTE packed index; a rank that holds only its slab
// hands them the buffer's address moved back by part_begin elements (never dereferenced below the slab).
inline int32_t* sorted_base(const tsq_ctx* c) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(c->d_sorted.p);
  return reinterpret_cast<int32_t*>(c->full_sorted ? a : a - (uintptr_t)c->part_begin * sizeof(int32_t));
}
template <typename T>
inline T* biased(T* p, uint64_t first) {
  return reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(p) - (uintptr_t)first * sizeof(T));
}

bool is_gap_or_space(unsigned char c) {
  return c == '-' || c == '.' || c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f';
}

// 256-entry byte -> symbol tables (0xff = dropped: '-', '.', whitespace).  Letter map of
// tweakseq/Core/Annotations/Consensus.cpp:61-69; anything that is not a letter -> X / N.
struct EncodeLut {
  uint8_t prot[256], nuc[256];
  EncodeLut() {
    for (int c = 0; c < 256; c++) {
      int u = (c >= 'a' && c <= 'z') ? c - 'a' + 'A' : c;
      prot[c] = (u >= 'A' && u <= 'Z') ? kProteinIndex[u - 'A'] : (uint8_t)22;
      nuc[c] = u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : (u == 'T' || u == 'U') ? 3 : 4;
      if (is_gap_or_space((unsigned char)c)) prot[c] = nuc[c] = 0xff;
    }
  }
};
const EncodeLut kLut;

// Packed-16 range analysis.  Values live as v + delta*(i+j) + BIAS in
// an unsigned 16-bit half; this returns BIAS and the largest padded length that stays inside
// [0, 65535] for every intermediate of the recurrence.
uint32_t bias_for(const tsq_ctx* c, uint32_t lpad) {
  const int gep = c->ge - c->delta;
  const long long b = 3LL * c->go + 2LL * c->ge + (long long)std::max(0, gep) * 2LL * (lpad + 1) + c->delta + 16;
  return (uint32_t)b;
}
bool fits16(const tsq_ctx* c, uint32_t lpad) {
  const long long hi = (long long)bias_for(c, lpad) + (long long)std::max(c->smax, 0) * lpad +
                       2LL * c->delta * (lpad + 1) + 16;
  return hi <= 65535;
}

// The packed wavefront kernel (wave16.cuh) is exact while the cells a warp holds at one time span
// less than 2^15 score units: window = cells in flight x per-step Lipschitz bound of the skewed DP.
constexpr long long kWave16WindowMax = tsq::kW16WindowMax;
inline long long lipschitz_of(const tsq_ctx* c) { return std::max(std::abs(c->smax), std::abs(c->smin)) + c->go + c->ge + 2 * c->delta; }
bool wave16_ok(const tsq_ctx* c, int nsym, uint32_t flags) {
  if (flags & TSQ_FLAG_NO_WAVE16) return false;
  const long long lip = std::max(std::abs(c->smax), std::abs(c->smin)) + c->go + c->ge + 2 * c->delta;
  return tsq::w16_window((uint32_t)nsym, lip) <= kWave16WindowMax;
}

// Which inter-task kernel takes the short sequences, and up to which length: the packed 16-bit one
// when the score range admits it; the 32-bit one (gotoh32.cuh) in identity mode (keys are wide) or when
// the gap/score parameters leave the packed kernel no usable range.  Longer sequences: wavefront.
constexpr uint32_t kMaxLen32 = 8192;
uint32_t inter_task_limit(uint32_t max_len16, uint32_t flags, bool* use_g32) {
  const bool idmode = (flags & TSQ_FLAG_IDENTITY) != 0;
  const bool force_wave = (flags & TSQ_FLAG_FORCE_S32) != 0;
  const bool g32 = idmode || (max_len16 < 64 && !force_wave);
  if (use_g32) *use_g32 = g32;
  return g32 ? (force_wave ? 0u : kMaxLen32) : max_len16;
}

// Largest sequence length the packed kernel can take (binary search on the range bound).
uint32_t max_len16_of(const tsq_ctx* c) {
  uint32_t a = 0, b = 60000;
  while (b - a > 1) {
    const uint32_t mid = (a + b) / 2;
    if (fits16(c, mid + 64)) a = mid; else b = mid;
  }
  return a;
}

// Contiguous partition of the sorted rows over `world` ranks, balanced by weighted DP cells
// (regime-2 cells weigh w2: 1.5 for the packed wavefront kernel, 2.4 for the 32-bit one --
// measured 9.2 : 6.1 : 3.7 TCUPS).  Rows
// of the packed kernel come in pairs (lo+2q, lo+2q+1), so cuts inside [lo, hi) fall on pair
// boundaries.  first_row has world+1 entries.
void plan_rows(const std::vector<uint32_t>& lens, uint32_t lo, uint32_t hi, int world,
               std::vector<uint32_t>& first_row, double w2 = 1.5) {
  const uint32_t n = (uint32_t)lens.size();
  std::vector<double> rowcost(n, 0.0);
  double s16 = 0, s32 = 0, tot = 0;  // suffix sums of lengths below / at-or-above hi
  for (uint32_t i = n; i-- > 0;) {
    rowcost[i] = (double)lens[i] * (i < hi ? s16 + w2 * s32 : w2 * s32);
    tot += rowcost[i];
    if (i < hi) s16 += lens[i]; else s32 += lens[i];
  }
  first_row.assign((size_t)world + 1, n);
  first_row[0] = 0;
  double acc = 0;
  uint32_t i = 0;
  for (int r = 1; r < world; r++) {
    const double target = tot * r / world;
    while (i < n && acc + rowcost[i] <= target) acc += rowcost[i++];
    if (i > lo && i < hi && ((i - lo) & 1u)) acc += rowcost[i++];
    first_row[(size_t)r] = std::min(i, n);
  }
}

}  // namespace

namespace {

// ---- tsq_upload, step 1: stable length sort, linear residue buffer, self scores, regime bounds ----
int host_sort_and_pack(tsq_ctx* c) {
  const uint32_t n = c->n;
  // ---- stable length sort ---------------------------------------------------------------
  // (keys = length << 32 | input index: an ordinary sort of them is the stable length sort; input that is
  //  already in order -- every fixed-length workload -- is recognised in one pass)
  c->perm.resize(n);
  {
    bool ordered = true;
    size_t prev = 0;
    for (uint32_t i = 0; i < n && ordered; i++) {
      const size_t l = (*c->encp)[i].size();
      ordered = l >= prev;
      prev = l;
    }
    if (ordered) {
      std::iota(c->perm.begin(), c->perm.end(), 0u);
    } else {
      std::vector<uint64_t> keys(n);
      for (uint32_t i = 0; i < n; i++) {
        const size_t l = (*c->encp)[i].size();
        if (l > 0x7fffffffu) return fail(c, TSQ_ERR_RANGE, "sequence too long");
        keys[i] = ((uint64_t)l << 32) | i;
      }
      std::sort(keys.begin(), keys.end());
      for (uint32_t i = 0; i < n; i++) c->perm[i] = (uint32_t)keys[i];
    }
  }
  c->lens.resize(n);
  c->loff.resize(n + 1);
  uint64_t total = 0;
  c->perm_identity = true;
  for (uint32_t i = 0; i < n; i++) {
    const size_t l = (*c->encp)[c->perm[i]].size();
    if (l > 0x7fffffffu) return fail(c, TSQ_ERR_RANGE, "sequence too long");
    c->lens[i] = (uint32_t)l;
    c->loff[i] = (uint32_t)total;
    total += (l + 15) & ~(size_t)15;   // 16-byte aligned starts: TMA bulk copies read tiles from here
    if (c->perm[i] != i) c->perm_identity = false;
  }
  if (total > 0xfffffff0ull) return fail(c, TSQ_ERR_RANGE, "more than 4 Gi residues");
  c->loff[n] = (uint32_t)total;
  c->lin_size = total + 2048;   // slack: the last TMA tile of a subject may run past its end
  {
    // upper bound of the whole staging blob (tables are laid out in device_upload)
    const size_t ngroups = ((size_t)n + 31) / 32;
    const size_t tables = (ngroups + 1) * 4 + ((size_t)n + 1) * 4 * 4 + (size_t)(c->nsym + 1) * c->nsym * 4 +
                          ((size_t)n + 2) * 8 + 8 * 16;
    TSQ_CUDA(c, c->h_blob.reserve(c->lin_size + tables));
  }
  uint8_t* const lin = c->h_blob.p;
  for (uint32_t i = 0; i < n; i++) {
    const std::vector<uint8_t>& e = (*c->encp)[c->perm[i]];
    if (!e.empty()) memcpy(lin + c->loff[i], e.data(), e.size());
    const size_t end = (size_t)c->loff[i] + e.size();
    memset(lin + end, 0, (size_t)c->loff[i + 1] - end);   // the pad up to the next 16-byte aligned start
  }
  memset(lin + total, 0, 2048);
  // ---- regimes --------------------------------------------------------------------------
  uint32_t lo = 0;
  while (lo < n && c->lens[lo] == 0) lo++;
  const bool idmode = (c->prm.flags & TSQ_FLAG_IDENTITY) != 0;
  const uint32_t inter_max = inter_task_limit(c->max_len16, c->prm.flags, &c->use_g32);
  c->inter_max = inter_max;
  uint32_t hi = lo;
  while (hi < n && c->lens[hi] <= inter_max) hi++;
  // key units: identity mode scales scores and penalties by M = 2^idshift > longest sequence
  c->idshift = 0;
  if (idmode) {
    uint32_t sh = 1;
    while ((1u << sh) <= (n ? c->lens[n - 1] : 0u)) sh++;
    c->idshift = sh;
  }
  {
    const int64_t M = 1ll << c->idshift;
    const uint32_t nsym = (uint32_t)c->nsym;
    c->go_k = (int)(c->go * M);
    c->ge_k = (int)(c->ge * M);
    c->smat_k.assign((size_t)(nsym + 1) * nsym, 0);
    for (uint32_t a = 0; a < nsym; a++)
      for (uint32_t b = 0; b < nsym; b++)
        c->smat_k[a * nsym + b] = (int32_t)(c->matrix[a * nsym + b] * M + ((idmode && a == b) ? 1 : 0));
    if (n > 0) {
      // 32-bit range of the keys: H <= max(S) * L from above; from below H >= the all-gap path
      // -(2 go + 2 L ge), E and F one more gap opening below that, t = H + S one min(S) below.
      const int64_t L = c->lens[n - 1];
      const int64_t bound = std::max<int64_t>((int64_t)std::max(c->smax, 0) * L, 3ll * c->go + (2 * L + 2) * c->ge) +
                            std::abs(c->smin) + 2;
      if (bound * M >= (1ll << 31) - (1ll << 20))
        return fail(c, TSQ_ERR_RANGE, "sequence of length %u: %s would not fit 32 bits", c->lens[n - 1],
                    idmode ? "identity-aware keys" : "scores");
      if ((c->prm.flags & TSQ_FLAG_SCORES_I16) && bound > 32767)
        return fail(c, TSQ_ERR_RANGE, "TSQ_FLAG_SCORES_I16: a score of a sequence of length %u could leave int16 (bound %lld)",
                    c->lens[n - 1], (long long)bound);
    }
  }
  if (c->perm_identity && lo > 0) c->perm_identity = false;  // empties are filled in by finalize
  c->lo = lo;
  c->hi = hi;
  // how a partitioned job's results come together (tsq_results_sharded) and what this rank must hold
  c->slab_mode = c->prm.part_world > 1 && c->perm_identity && c->idshift == 0;
  c->full_sorted = c->prm.part_world == 1 || (!c->leader && c->prm.part_rank == 0 && !c->slab_mode);
  return TSQ_OK;
}

// ---- tsq_upload, step 2: this rank's rows, task lists of the three kernels, strip width ------------
int host_plan_work(tsq_ctx* c) {
  const uint32_t n = c->n, lo = c->lo, hi = c->hi;
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  // ---- partition of the sorted rows across ranks (contiguous, balanced by DP cells) ---------
  const int world = c->prm.part_world, rank = c->prm.part_rank;
  c->use_w16 = wave16_ok(c, c->nsym, c->prm.flags) && c->idshift == 0;   // identity keys are too wide to pack
  plan_rows(c->lens, lo, hi, world, c->first_row, c->use_w16 ? 1.5 : 2.4);
  auto boundary = [&](int r) -> uint32_t { return c->first_row[(size_t)r]; };
  c->row_a = boundary(rank);
  c->row_b = boundary(rank + 1);
  c->part_begin = (n >= 2 && c->row_a + 1 < n) ? tri(c->row_a, c->row_a + 1, n) : npairs;
  c->part_end = (n >= 2 && c->row_b + 1 < n) ? tri(c->row_b, c->row_b + 1, n) : npairs;
  if (c->part_begin > c->part_end) c->part_begin = c->part_end;

  // ---- tasks of the inter-task kernel over [lo, hi) ------------------------------------------------
  // packed kernel: one task row per query PAIR (rows lo+2q, lo+2q+1); 32-bit kernel: per query (lo+q)
  const uint32_t per = c->use_g32 ? 1u : 2u;
  const uint32_t nq_all = c->use_g32 ? (hi > lo ? hi - lo - 1 : 0u) : (hi - lo) / 2;
  auto row_to_q = [&](uint32_t row) -> uint32_t {
    if (row <= lo) return 0;
    return std::min(nq_all, (row - lo + per - 1) / per);
  };
  c->q_begin = row_to_q(c->row_a);
  c->q_end = row_to_q(c->row_b);
  if (c->q_end < c->q_begin) c->q_end = c->q_begin;
  const uint32_t nq = c->q_end - c->q_begin;
  c->task_prefix.assign((size_t)nq + 1, 0);
  c->cells16 = 0;
  c->cells32 = 0;
  c->pairs_part = 0;
  {
    std::vector<uint64_t> suffix(n + 1, 0);
    for (uint32_t i = n; i-- > 0;) suffix[i] = suffix[i + 1] + c->lens[i];
    for (uint32_t r = 0; r < nq; r++) {
      const uint32_t q = c->q_end - 1 - r;
      const uint32_t a1 = lo + per * q;
      const uint32_t nsub = hi - a1 - 1;
      c->task_prefix[r + 1] = c->task_prefix[r] + (nsub + 31) / 32;
      if (c->use_g32) {
        c->cells32 += (uint64_t)c->lens[a1] * (suffix[a1 + 1] - suffix[hi]);
        c->pairs_part += (uint64_t)(hi - a1 - 1);
      } else {
        c->cells16 += (uint64_t)c->lens[a1] * (suffix[a1 + 1] - suffix[hi]) +
                      (uint64_t)c->lens[a1 + 1] * (suffix[a1 + 2] - suffix[hi]);
        c->pairs_part += (uint64_t)(hi - a1 - 1) + (hi - a1 - 2);
      }
    }
  }
  // ---- tasks of the 32-bit wavefront kernel: every pair with a sequence beyond the packed range,
  //      restricted to this rank's rows, biggest pairs first -------------------------------------
  c->pairs32.clear();
  c->tasks16w.clear();
  if (hi < n) {
    const uint32_t ra = std::max(c->row_a, lo);
    if (c->use_w16) {
      // (query pair, subject): subject j long, queries i < j of this rank's rows, two at a time
      for (uint32_t j = n; j-- > hi;) {
        const uint32_t top = std::min(c->row_b, j);  // queries in [ra, top)
        uint32_t i = top;
        while (i > ra) {
          if (i - ra >= 2) {
            c->tasks16w.push_back(make_uint4(i - 2, i - 1, j, 0));
            c->cells32 += ((uint64_t)c->lens[i - 2] + c->lens[i - 1]) * c->lens[j];
            c->pairs_part += 2;
            i -= 2;
          } else {
            c->tasks16w.push_back(make_uint4(i - 1, i - 1, j, 0));
            c->cells32 += (uint64_t)c->lens[i - 1] * c->lens[j];
            c->pairs_part += 1;
            i -= 1;
          }
        }
      }
    } else {
      for (uint32_t i = c->row_b; i-- > ra;) {
        for (uint32_t j = n; j-- > std::max(i + 1, hi);) {
          c->pairs32.push_back(make_uint2(i, j));
          c->cells32 += (uint64_t)c->lens[i] * c->lens[j];
        }
      }
      c->pairs_part += c->pairs32.size();
    }
  }
  // ---- strip width: least estimated work over the instantiated variants -----------------------
  // A strip of K columns costs about K + 2.5 cell-times per row (the row's letter fetch, boundary
  // load/store and loop control are worth ~2.5 cells: profiles/ r01 sweep), padding included.
  {
    // (tasks per distinct query length first: a fixed-length workload is one run, and the estimate below costs a
    //  division per run and variant instead of one per query pair and variant)
    std::vector<std::pair<uint32_t, double>> runs;
    for (uint32_t r = 0; r < nq; r++) {
      const uint32_t q = c->q_end - 1 - r;
      const uint32_t l2 = c->lens[c->use_g32 ? lo + q : lo + 2 * q + 1];
      const double nt = (double)(c->task_prefix[r + 1] - c->task_prefix[r]);
      if (!runs.empty() && runs.back().first == l2) runs.back().second += nt;
      else runs.emplace_back(l2, nt);
    }
    double best = 1e300;
    int bestK = tsq::kStripWidths[0];
    for (int v = 0; v < tsq::kNumStripWidths; v++) {
      const int K = tsq::kStripWidths[v];
      double work = 0;
      for (const auto& run : runs) work += (double)((run.first + K - 1) / K) * (K + 2.5) * run.second;
      if (K <= 32) work *= 1.05;  // the 16-warp variants run ~4-5 % slower per cell (r01 sweep)
      if (work < best * 0.999 || (work <= best * 1.001 && K > bestK)) {
        if (work < best) best = work;
        bestK = K;
      }
    }
    c->K = bestK;
    if (const char* fk = getenv("TSQ_FORCE_K")) {  // developer override for tuning runs
      tsq::G16Launch v;
      if (tsq::g16_variant(atoi(fk), (uint32_t)c->nsym, &v)) c->K = atoi(fk);
    }
  }

  return TSQ_OK;
}

// ---- tsq_upload, step 3: layout (group offsets) of the interleaved subject database -------------------
int host_build_subject_db(tsq_ctx* c) {
  const uint32_t n = c->n;
  // ---- 32-way interleaved subject database of the packed kernel ----------------------------------
  // Per residue the 16-bit byte offset of its profile row (letter * STRIDE(K) * 4), two rows per
  // 32-bit word, right-aligned to an even row count (a pad entry leads an odd-length sequence),
  // so one coalesced 128-byte load per warp feeds one row pair of all 32 lanes.
  {
    tsq::G16Launch kv;
    if (!tsq::g16_variant(c->K, (uint32_t)c->nsym, &kv)) return fail(c, TSQ_ERR_INVALID, "no kernel variant K=%d", c->K);
    const uint32_t scale = (uint32_t)kv.stride * 4u;
    const uint32_t ngroups = (n + 31) / 32;
    c->goff.assign(ngroups + 1, 0);
    uint64_t words = 0;
    for (uint32_t g = 0; g < ngroups; g++) {
      c->goff[g] = (uint32_t)words;
      const uint32_t last = std::min(n, (g + 1) * 32) - 1;
      const uint32_t len16 = std::min(c->lens[last], c->inter_max);  // longer ones never enter this kernel
      const uint32_t rows2 = (len16 + 1) / 2 + 3;                    // +3: two-word prefetch slack
      words += (uint64_t)rows2 * 32;
      if (words > 0xfffffff0ull) return fail(c, TSQ_ERR_RANGE, "interleaved database too large");
    }
    c->goff[ngroups] = (uint32_t)words;
    c->dbw_size = words;   // the words themselves are written on the device (subject_db_launch)
    c->db_scale = scale;
  }

  return TSQ_OK;
}

// ---- tsq_upload, step 4: device buffers and host -> device copies --------------------------------------
int device_upload(tsq_ctx* c) {
  const uint32_t n = c->n;
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  const uint32_t nsym = (uint32_t)c->nsym;
  // ---- lay the small tables out behind the residues in the pinned blob -----------------------------
  size_t off = c->lin_size;
  auto carve = [&](size_t bytes) -> size_t {
    off = (off + 15) & ~(size_t)15;
    const size_t at = off;
    off += bytes;
    return at;
  };
  const size_t o_goff = carve(c->goff.size() * 4), o_loff = carve(c->loff.size() * 4), o_lens = carve(((size_t)n + 1) * 4),
               o_perm = carve(((size_t)n + 1) * 4), o_self = carve(((size_t)n + 1) * 4),
               o_sbias = carve((size_t)(nsym + 1) * nsym * 4), o_prefix = carve(c->task_prefix.size() * 8);
  c->blob_size = off;
  if (off > c->h_blob.cap) return fail(c, TSQ_ERR_INVALID, "internal: staging blob bound too small");
  uint8_t* const hb = c->h_blob.p;
  memcpy(hb + o_goff, c->goff.data(), c->goff.size() * 4);
  memcpy(hb + o_loff, c->loff.data(), c->loff.size() * 4);
  if (n) {
    memcpy(hb + o_lens, c->lens.data(), (size_t)n * 4);
    memcpy(hb + o_perm, c->perm.data(), (size_t)n * 4);
    int32_t* self_sorted = reinterpret_cast<int32_t*>(hb + o_self);
    for (uint32_t i = 0; i < n; i++) self_sorted[i] = (*c->selfp)[c->perm[i]];
  }
  {  // biased score table of the packed kernels: S + 2 delta >= 0, one extra all-zero row
    uint32_t* sbias = reinterpret_cast<uint32_t*>(hb + o_sbias);
    memset(sbias, 0, (size_t)(nsym + 1) * nsym * 4);
    for (uint32_t a = 0; a < nsym; a++)
      for (uint32_t b = 0; b < nsym; b++) sbias[a * nsym + b] = (uint32_t)(c->matrix[a * nsym + b] + 2 * c->delta);
  }
  memcpy(hb + o_prefix, c->task_prefix.data(), c->task_prefix.size() * 8);

  // ---- H2D: one copy ------------------------------------------------------------------------------
  cudaStream_t s = c->stream;
  TSQ_CUDA(c, c->d_blob.reserve(c->h_blob.cap));
  uint8_t* const db = c->d_blob.p;
  c->d_lin.p = db;
  c->d_goff.p = reinterpret_cast<uint32_t*>(db + o_goff);
  c->d_loff.p = reinterpret_cast<uint32_t*>(db + o_loff);
  c->d_lens.p = reinterpret_cast<uint32_t*>(db + o_lens);
  c->d_perm.p = reinterpret_cast<uint32_t*>(db + o_perm);
  c->d_self.p = reinterpret_cast<int32_t*>(db + o_self);
  c->d_sbias.p = reinterpret_cast<uint32_t*>(db + o_sbias);
  c->d_prefix.p = reinterpret_cast<unsigned long long*>(db + o_prefix);
  TSQ_CUDA(c, c->d_dbw.reserve(c->dbw_size));
  TSQ_CUDA(c, c->d_counter.reserve(16));
  TSQ_CUDA(c, c->d_sorted.reserve(c->full_sorted ? npairs : c->part_end - c->part_begin));
  TSQ_CUDA(c, cudaMemcpyAsync(db, hb, c->blob_size, cudaMemcpyHostToDevice, s));
  c->st.upload_launches = 0;
  if (c->dbw_size && c->hi > c->lo) c->st.upload_launches = 1;
  if (c->dbw_size && c->hi > c->lo)
    TSQ_CUDA(c, tsq::subject_db_launch(c->d_lin.p, c->d_loff.p, c->d_lens.p, c->d_goff.p, c->d_dbw.p, c->n, c->lo, c->hi,
                                       c->db_scale, s));
  if (!c->tasks16w.empty()) {
    TSQ_CUDA(c, c->d_tasks16w.reserve(c->tasks16w.size()));
    TSQ_CUDA(c, cudaMemcpyAsync(c->d_tasks16w.p, c->tasks16w.data(), c->tasks16w.size() * sizeof(uint4), cudaMemcpyHostToDevice, s));
  }
  if (!c->pairs32.empty() || c->use_g32) {
    const std::vector<int32_t>& smat = c->smat_k;   // key units (= plain scores unless identity mode)
    TSQ_CUDA(c, c->d_smat.reserve(smat.size()));
    TSQ_CUDA(c, c->d_pairs32.reserve(c->pairs32.size() + 1));
    TSQ_CUDA(c, cudaMemcpyAsync(c->d_smat.p, smat.data(), smat.size() * 4, cudaMemcpyHostToDevice, s));
    if (!c->pairs32.empty())
      TSQ_CUDA(c, cudaMemcpyAsync(c->d_pairs32.p, c->pairs32.data(), c->pairs32.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
  }
  // no host wait here: the kernels of tsq_compute queue up behind the copy on the same stream, and the next
  // tsq_upload waits for the stream before it touches the staging blob again
  c->st.h2d_bytes = c->blob_size + c->tasks16w.size() * sizeof(uint4) + c->pairs32.size() * sizeof(uint2) +
                    ((!c->pairs32.empty() || c->use_g32) ? c->smat_k.size() * 4 : 0);
  return TSQ_OK;
}

}  // namespace

namespace {

struct StreamChunk {   // one launch of a streamed compute: tasks [t0, t1) = sorted rows [row0, row1)
  unsigned long long t0, t1;
  uint32_t row0, row1;
};
int stream_chunk_out(tsq_ctx* c, const StreamChunk& ch, cudaStream_t compute_stream, size_t index);   // defined below
int stream_ranges_out(tsq_ctx* c, const std::vector<StreamChunk>& ranges);                            // defined below

// cuStreamWaitValue32, resolved through the runtime (the library does not link libcuda): lets the copy stream wait
// for a counter the RUNNING kernel advances, so that row ranges leave without cutting the launch into pieces.
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
  static const StreamWaitValue32Fn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      f = nullptr;
    }
    return reinterpret_cast<StreamWaitValue32Fn>(f);
  }();
  return fn;
}
constexpr unsigned kDoneSlots = 32;

// kernel_ms runs from the first launch of a tsq_compute, not from its first allocation: a context's first compute
// reserves its scratch between the two.
#define TSQ_MARK_START(c, s)                          \
  do {                                                \
    if (!(c)->ev0_set) {                              \
      TSQ_CUDA((c), cudaEventRecord((c)->ev0, (s))); \
      (c)->ev0_set = true;                            \
    }                                                 \
  } while (0)

// Whether the results of this job can leave in row-range chunks: the packed kernel only, scores in place
// (sorted order = submitted order, no identity keys), the whole triangle or a sharded slab.
bool can_stream(const tsq_ctx* c) {
  return c->stream_out && !c->use_g32 && c->idshift == 0 && c->tasks16w.empty() && c->pairs32.empty() &&
         (c->prm.part_world == 1 ? c->perm_identity : c->slab_mode);
}

// ---- tsq_compute: the three score kernels, each enqueued on stream s when it has tasks ---------------
int enqueue_gotoh16(tsq_ctx* c, cudaStream_t s, uint32_t& launches) {
  const uint32_t nq = c->q_end - c->q_begin;
  const unsigned long long ntasks = c->task_prefix.empty() ? 0 : c->task_prefix[nq];
  if (ntasks > 0) {
    tsq::G16Launch v;
    if (!tsq::g16_variant(c->K, (uint32_t)c->nsym, &v)) return fail(c, TSQ_ERR_INVALID, "no kernel variant K=%d", c->K);
    const int warps_per_cta = v.tpb / 32;
    int grid = c->sm_count * v.ctas_sm;
    const unsigned long long need = (ntasks + warps_per_cta - 1) / warps_per_cta;
    if ((unsigned long long)grid > need) grid = (int)need;
    // (r01: shrinking the grid so that every warp runs a whole number of tasks was measured and is
    //  slower -- 7.8 vs 8.2 TCUPS on C2: warps of a partly filled last wave speed up on their own.)
    if (const char* eg = getenv("TSQ_GRID")) grid = std::max(1, atoi(eg));
    const uint32_t maxlen = c->hi > c->lo ? c->lens[c->hi - 1] : 0;
    const uint32_t bnd_rows = maxlen + 8;  // the row loop prefetches up to 3 rows past the end
    TSQ_CUDA(c, c->d_bnd.reserve_zeroed((size_t)grid * warps_per_cta * bnd_rows * 32));
    TSQ_CUDA(c, cudaMemsetAsync(c->d_counter.p, 0, sizeof(unsigned long long), s));
    if (c->use_g32) {   // 32-bit inter-task kernel (identity mode, or parameters too wide for 16 bits)
      tsq::G32Params g{};
      g.dbw = c->d_dbw.p;
      g.goff = c->d_goff.p;
      g.lin = c->d_lin.p;
      g.loff = c->d_loff.p;
      g.lens = c->d_lens.p;
      g.task_prefix = c->d_prefix.p;
      g.counter = c->d_counter.p;
      g.cancel = c->d_cancel;
      g.bnd = reinterpret_cast<int2*>(c->d_bnd.p);
      g.smat = c->d_smat.p;
      g.out = sorted_base(c);
      g.ntasks = ntasks;
      g.bnd_rows = bnd_rows;
      g.n_total = c->n;
      g.lo = c->lo;
      g.hi = c->hi;
      g.q_begin = c->q_begin;
      g.q_end = c->q_end;
      g.nsym = (uint32_t)c->nsym;
      g.go = c->go_k;
      g.ge = c->ge_k;
      g.one = 1;
      TSQ_MARK_START(c, s);
      TSQ_CUDA(c, tsq::g32_launch(c->K, grid, g, s));
      launches++;
      return TSQ_OK;
    }
    const uint32_t lpad = ((maxlen + c->K - 1) / c->K) * c->K;
    tsq::G16Params p{};
    p.dbw = c->d_dbw.p;
    p.goff = c->d_goff.p;
    p.lin = c->d_lin.p;
    p.loff = c->d_loff.p;
    p.lens = c->d_lens.p;
    p.task_prefix = c->d_prefix.p;
    p.counter = c->d_counter.p;
    p.cancel = c->d_cancel;
    p.bnd = c->d_bnd.p;
    p.sbias = c->d_sbias.p;
    p.out = sorted_base(c);
    p.ntasks = ntasks;
    p.bnd_rows = bnd_rows;
    p.n_total = c->n;
    p.lo = c->lo;
    p.hi = c->hi;
    p.q_begin = c->q_begin;
    p.q_end = c->q_end;
    p.nsym = (uint32_t)c->nsym;
    p.bias = bias_for(c, lpad);
    p.delta = c->delta;
    p.go = c->go;
    p.gep = c->ge - c->delta;
    p.negge2 = ((uint32_t)(-(c->ge - c->delta)) & 0xffffu) * 0x10001u;
    p.goe2 = (uint32_t)(c->go + c->ge - c->delta) * 0x10001u;
    if (!fits16(c, lpad)) return fail(c, TSQ_ERR_RANGE, "internal: padded length %u outside the 16-bit bound", lpad);
    // (r02: an L2 access-policy window over the scratch column -- persisting for the share that fits the carve-out,
    //  streaming for the rest -- was measured on configs[1] and changed neither the time nor the DRAM bytes
    //  (profiles/l2_policy_ab_r02.txt), so the launch carries no policy.)
    c->streamed = false;
    // how many launches: one, unless the results are streamed out and the job is long enough that every launch
    // still fills the device for many waves (a launch ends in a partly filled wave: ~half a task time lost)
    size_t nchunks = 1;
    if (can_stream(c)) {
      const unsigned long long resident = (unsigned long long)grid * warps_per_cta;
      nchunks = (size_t)std::min<unsigned long long>(8, ntasks / (16 * std::max<unsigned long long>(resident, 1)));
      if (nchunks < 1) nchunks = 1;
      if (const char* e = getenv("TSQ_STREAM_CHUNKS")) nchunks = (size_t)std::min(8, std::max(1, atoi(e)));   // tests: force a split
      if (nchunks > nq) nchunks = std::max<size_t>(nq, 1);
    }
    const bool by_counters = can_stream(c) && stream_wait_value32() != nullptr && getenv("TSQ_STREAM_LAUNCHES") == nullptr;
    if (nchunks == 1 && !can_stream(c)) {
      TSQ_MARK_START(c, s);
      TSQ_CUDA(c, tsq::g16_launch(c->K, grid, p, s, nullptr));
      launches++;
    } else if (by_counters) {
      // ---- ONE launch; row ranges leave as their tasks finish -----------------------------------------------
      // The kernel writes the distance next to every score and ticks its row range's counter at the end of a
      // task; the copy stream waits for a range's count (stream memory operation) and copies the range out while
      // the same launch works on the next rows.  No launch boundary, so no partly filled wave per range, and
      // short jobs (configs[1]: 4.4 waves) stream too.
      const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
      const uint64_t n = c->n, npairs = n < 2 ? 0 : n * (n - 1) / 2;
      const uint64_t first = c->full_sorted ? 0 : c->part_begin;
      const uint64_t out_bytes = c->pairs_part * (want_dist ? 12ull : 4ull);
      size_t nr = (size_t)std::min<uint64_t>(tsq::kG16MaxRanges, std::max<uint64_t>(1, out_bytes / (384u << 10)));
      if (const char* e = getenv("TSQ_STREAM_CHUNKS")) nr = (size_t)std::min(tsq::kG16MaxRanges, std::max(1, atoi(e)));
      if (nr > nq) nr = std::max<size_t>(nq, 1);
      if (c->chunk_ev.empty()) {
        c->chunk_ev.resize(1, nullptr);
        TSQ_CUDA(c, cudaEventCreateWithFlags(&c->chunk_ev[0], cudaEventDisableTiming));
      }
      if (!c->copy_stream) TSQ_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      TSQ_CUDA(c, c->d_done.reserve(kDoneSlots));
      if (want_dist) {
        TSQ_CUDA(c, c->d_dist.reserve(c->full_sorted ? npairs : c->part_end - c->part_begin));
        p.self = c->d_self.p;
        p.out_dist = biased(c->d_dist.p, first);
      }
      if (c->prm.flags & TSQ_FLAG_SCORES_I16) {
        TSQ_CUDA(c, c->d_scores16.reserve(c->full_sorted ? npairs : c->part_end - c->part_begin));
        p.out16 = biased(c->d_scores16.p, first);
      }
      p.done = c->d_done.p;
      p.nranges = (uint32_t)nr;
      std::vector<StreamChunk> ranges(nr);
      // Range boundaries by TASK count (= output bytes: every task is 64 pairs), not by query pairs -- the first rows of
      // the triangle are the longest and come last in task order -- and the last of the equal parts is cut again into
      // 1/2, 1/4, 1/8, 1/8: what is still on the device when the kernel ends is under 1 % of the result.
      std::vector<uint32_t> rcut(nr + 1, nq);
      rcut[0] = 0;
      for (size_t k = 1; k < nr; k++) {
        double f;
        if (nr < 5) f = (double)k / (double)nr;
        else {
          const size_t base = nr - 3;                       // equal parts; the last one holds four ranges
          if (k < base) f = (double)k / (double)base;
          else {
            static const double sub[3] = {0.5, 0.75, 0.875};
            f = ((double)(base - 1) + sub[k - base]) / (double)base;
          }
        }
        const unsigned long long target = (unsigned long long)(f * (double)ntasks);
        const uint32_t r = (uint32_t)(std::lower_bound(c->task_prefix.begin(), c->task_prefix.begin() + nq + 1, target) - c->task_prefix.begin());
        rcut[k] = std::max(rcut[k - 1], std::min(r, nq));
      }
      for (size_t k = 0; k < nr; k++) {
        const uint32_t r0 = rcut[k], r1 = rcut[k + 1];
        StreamChunk& ch = ranges[k];
        ch.t0 = c->task_prefix[r0];
        ch.t1 = c->task_prefix[r1];
        ch.row0 = c->lo + 2 * (c->q_end - r1);           // query pairs q_end-r1 .. q_end-1-r0
        ch.row1 = c->lo + 2 * (c->q_end - r0);
        if (k + 1 == nr) ch.row0 = std::min(ch.row0, c->row_a);
        if (k == 0) ch.row1 = std::max(ch.row1, std::min(c->row_b, c->n));
        p.range_end[k] = ch.t1;
      }
      // the copies of an earlier streamed compute read these counters: they must be through before the reset
      TSQ_CUDA(c, cudaStreamWaitEvent(s, c->fin_ev, 0));
      TSQ_CUDA(c, cudaMemsetAsync(c->d_done.p, 0, kDoneSlots * sizeof(unsigned int), s));
      TSQ_CUDA(c, cudaEventRecord(c->chunk_ev[0], s));
      TSQ_MARK_START(c, s);
      TSQ_CUDA(c, tsq::g16_launch(c->K, grid, p, s, nullptr));
      launches++;
      TSQ_CUDA(c, cudaEventRecord(c->ev1, s));
      const int rcs = stream_ranges_out(c, ranges);
      if (rcs != TSQ_OK) return rcs;
      c->streamed = true;
    } else {
      // chunk k = task rows r in [r_k, r_k+1): tasks [prefix[r_k], prefix[r_k+1]), i.e. query pairs
      // q_end-1-r, which are consecutive sorted rows -- long queries first, as the task order has it
      if (c->chunk_ev.size() < nchunks) {
        const size_t old = c->chunk_ev.size();
        c->chunk_ev.resize(nchunks, nullptr);
        for (size_t k = old; k < nchunks; k++) TSQ_CUDA(c, cudaEventCreateWithFlags(&c->chunk_ev[k], cudaEventDisableTiming));
      }
      if (!c->copy_stream) TSQ_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      if (!c->h_starts) TSQ_CUDA(c, BlockCache::get().take(-1, 8 * sizeof(unsigned long long), (void**)&c->h_starts));
      c->streamed_bytes = 0;
      for (size_t k = 0; k < nchunks; k++) {
        const uint32_t r0 = (uint32_t)((unsigned long long)nq * k / nchunks), r1 = (uint32_t)((unsigned long long)nq * (k + 1) / nchunks);
        if (r1 <= r0) continue;
        StreamChunk ch;
        ch.t0 = c->task_prefix[r0];
        ch.t1 = c->task_prefix[r1];
        ch.row0 = c->lo + 2 * (c->q_end - r1);           // query pairs q_end-r1 .. q_end-1-r0
        ch.row1 = c->lo + 2 * (c->q_end - r0);
        if (k + 1 == nchunks) ch.row0 = std::min(ch.row0, c->row_a);            // rows without tasks of their own (none today)
        if (k == 0) ch.row1 = std::max(ch.row1, std::min(c->row_b, c->n));     // a last odd row has no pairs: empty range
        c->h_starts[k] = ch.t0;
        TSQ_CUDA(c, cudaMemcpyAsync(c->d_counter.p, &c->h_starts[k], sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        p.ntasks = ch.t1;
        TSQ_MARK_START(c, s);
        TSQ_CUDA(c, tsq::g16_launch(c->K, grid, p, s, nullptr));
        launches++;
        if (k + 1 == nchunks) TSQ_CUDA(c, cudaEventRecord(c->ev1, s));   // kernel_ms: the kernels, not the enqueues behind them
        const int rc = stream_chunk_out(c, ch, s, k);
        if (rc != TSQ_OK) return rc;
      }
      c->streamed = true;
    }
  }
  return TSQ_OK;
}

int enqueue_wave16(tsq_ctx* c, cudaStream_t s, uint32_t& launches) {
  if (!c->tasks16w.empty()) {
    tsq::W32Launch v;
    if (!tsq::w16_variant((uint32_t)c->nsym, lipschitz_of(c), &v)) return fail(c, TSQ_ERR_INVALID, "no packed wavefront kernel variant");
    const int warps_per_cta = v.tpb / 32;
    int grid = c->sm_count * v.ctas_sm;
    const unsigned long long need = (c->tasks16w.size() + warps_per_cta - 1) / warps_per_cta;
    if ((unsigned long long)grid > need) grid = (int)need;
    const uint32_t bnd_rows = c->lens[c->n - 1] + 8;
    TSQ_CUDA(c, c->d_bnd16w.reserve_zeroed((size_t)grid * warps_per_cta * tsq::w16_slot_elems(bnd_rows)));
    TSQ_CUDA(c, cudaMemsetAsync(c->d_counter.p + 2, 0, 12 * sizeof(unsigned long long), s));
    tsq::W16Params w{};
    w.lin = c->d_lin.p;
    w.loff = c->d_loff.p;
    w.lens = c->d_lens.p;
    w.tasks = c->d_tasks16w.p;
    w.counter = c->d_counter.p + 2;
    w.cancel = c->d_cancel;
    w.fault = c->d_cancel + 1;
    {   // test hook (tests/test_gpu_parity.py): the first task arms a tile barrier whose copy never comes
      const char* fi = getenv("TSQ_FAULT_INJECT");
      w.inject_fault = (fi && strcmp(fi, "tma") == 0) ? 1u : 0u;
    }
    w.bnd = c->d_bnd16w.p;
    w.sbias = c->d_sbias.p;
    w.out = sorted_base(c);
    w.ntasks = c->tasks16w.size();
    w.bnd_rows = bnd_rows;
    w.n_total = c->n;
    w.nsym = (uint32_t)c->nsym;
    w.delta = c->delta;
    w.go = c->go;
    w.gep = c->ge - c->delta;
    w.goep = c->go + c->ge - c->delta;
    w.negge2 = ((uint32_t)(-(c->ge - c->delta)) & 0xffffu) * 0x10001u;
    TSQ_MARK_START(c, s);
    TSQ_CUDA(c, tsq::w16_launch(grid, w, lipschitz_of(c), s));
    launches++;
  }
  return TSQ_OK;
}

int enqueue_wave32(tsq_ctx* c, cudaStream_t s, uint32_t& launches) {
  if (!c->pairs32.empty()) {
    tsq::W32Launch v;
    if (!tsq::w32_variant((uint32_t)c->nsym, &v)) return fail(c, TSQ_ERR_INVALID, "no wavefront kernel variant");
    const int warps_per_cta = v.tpb / 32;
    int grid = c->sm_count * v.ctas_sm;
    const unsigned long long need = (c->pairs32.size() + warps_per_cta - 1) / warps_per_cta;
    if ((unsigned long long)grid > need) grid = (int)need;
    const uint32_t bnd_rows = c->lens[c->n - 1] + 8;
    TSQ_CUDA(c, c->d_bnd32.reserve_zeroed((size_t)grid * warps_per_cta * bnd_rows));
    TSQ_CUDA(c, cudaMemsetAsync(c->d_counter.p + 1, 0, sizeof(unsigned long long), s));
    tsq::W32Params w{};
    w.lin = c->d_lin.p;
    w.loff = c->d_loff.p;
    w.lens = c->d_lens.p;
    w.pairs = c->d_pairs32.p;
    w.counter = c->d_counter.p + 1;
    w.cancel = c->d_cancel;
    w.bnd = c->d_bnd32.p;
    w.smat = c->d_smat.p;
    w.out = sorted_base(c);
    w.ntasks = c->pairs32.size();
    w.bnd_rows = bnd_rows;
    w.n_total = c->n;
    w.nsym = (uint32_t)c->nsym;
    w.go = c->go_k;
    w.ge = c->ge_k;
    w.one = 1;
    TSQ_MARK_START(c, s);
    TSQ_CUDA(c, tsq::w32_launch(grid, w, s));
    launches++;
  }
  return TSQ_OK;
}

}  // namespace

extern "C" {

int tsq_version(int* major, int* minor) {
  if (major) *major = TSQ_VERSION_MAJOR;
  if (minor) *minor = TSQ_VERSION_MINOR;
  return TSQ_OK;
}

const char* tsq_version_string(void) { return "tsq-b200 0.1 (sm_100a Gotoh all-vs-all)"; }

const char* tsq_status_string(int s) {
  switch (s) {
    case TSQ_OK: return "ok";
    case TSQ_ERR_INVALID: return "invalid argument";
    case TSQ_ERR_NO_DEVICE: return "no sm_100 CUDA device";
    case TSQ_ERR_CUDA: return "CUDA error";
    case TSQ_ERR_NOMEM: return "out of memory";
    case TSQ_ERR_CANCELLED: return "cancelled";
    case TSQ_ERR_STATE: return "call out of order";
    case TSQ_ERR_IO: return "I/O error";
    case TSQ_ERR_MATRIX: return "bad substitution matrix";
    case TSQ_ERR_RANGE: return "score range exceeded";
    default: return "unknown status";
  }
}

int tsq_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; d++) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ok++;
  }
  return ok;
}

void tsq_default_params(tsq_params* p) {
  if (!p) return;
  memset(p, 0, sizeof *p);
  p->struct_size = (uint32_t)sizeof(tsq_params);
  p->alphabet = TSQ_PROTEIN;
  p->gap_open = -1;
  p->gap_extend = -1;
  p->matrix = nullptr;
  p->device = 0;
  p->part_rank = 0;
  p->part_world = 1;
  p->flags = 0;
  p->n_devices = 1;
}

}  // extern "C"

namespace {

// Usable devices: compute capability 10.x (sm_100a SASS only: nothing else can run these kernels).
bool device_usable(int d) {
  int major = 0;
  return cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10;
}

void copy_error(tsq_ctx* to, const tsq_ctx* from) {
  if (to && from) to->err = from->err;
}

// ---- host result buffers ---------------------------------------------------------------------------------
// Who owns them: a child of a multi-device context writes into its leader's; everything else into its own.
tsq_ctx* result_owner(tsq_ctx* c) { return c->leader ? c->leader : c; }

// A partitioned rank without caller buffers keeps only its slab on the host (a rank of world 8 at 100 000
// sequences: 2.5 GB, not 20 GB of page-locked memory): [first, first + count) of the packed triangle.
void host_extent(const tsq_ctx* c, uint64_t* first, uint64_t* count) {
  const uint64_t n = c->n, npairs = n < 2 ? 0 : n * (n - 1) / 2;
  const bool slab_only = c->slab_mode && c->kids.empty() && !c->leader && !c->ext_scores;
  *first = slab_only ? c->part_begin : 0;
  *count = slab_only ? c->part_end - c->part_begin : npairs;
}

// Page-locks the whole pages INSIDE [p, p + bytes) of caller memory, unless an earlier call already did.  The
// partial pages at either end stay pageable: they may be shared with a neighbouring array or with another rank's
// slab (whose owner would register them too, and overlapping registrations are refused); copy_out() moves them
// with small pageable copies, because a copy may not span page-locked and pageable memory.
constexpr uintptr_t kPage = 4096;
inline uintptr_t page_up(uintptr_t x) { return (x + kPage - 1) & ~(kPage - 1); }
inline uintptr_t page_down(uintptr_t x) { return x & ~(kPage - 1); }

int register_range(tsq_ctx* c, void* p, size_t bytes) {
  if (!p || bytes == 0) return TSQ_OK;
  const uintptr_t a = page_up(reinterpret_cast<uintptr_t>(p));
  const uintptr_t b = page_down(reinterpret_cast<uintptr_t>(p) + bytes);
  if (b <= a) return TSQ_OK;
  for (void* r : c->ext_registered)
    if (r == reinterpret_cast<void*>(a)) return TSQ_OK;
  const cudaError_t e = cudaHostRegister(reinterpret_cast<void*>(a), b - a, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {   // the caller pinned it already (cudaHostAlloc, or its own register)
    cudaGetLastError();
    return TSQ_OK;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(c, TSQ_ERR_CUDA, "cudaHostRegister of the caller's result buffer failed: %s", cudaGetErrorString(e));
  }
  c->ext_registered.push_back(reinterpret_cast<void*>(a));
  return TSQ_OK;
}

// Device -> host copy of one slab, by context c.  Into the library's own pinned buffers: one copy.  Into caller
// memory (registered as above): the whole pages go straight to their place; the partial pages at either end (which
// may be shared with a neighbour, so they are not page-locked, and a copy may not span page-locked and pageable
// memory) go into c's pinned bounce block, asynchronously like everything else, and the host moves those few
// bytes to their place after the synchronize (apply_fixups).  A pageable cudaMemcpyAsync would block the host until
// the kernels before it have finished.
constexpr size_t kBounceBytes = 64 * 4096;
cudaError_t copy_out(tsq_ctx* c, const tsq_ctx* owner, void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return cudaSuccess;
  if (!owner->ext_scores) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
  const uintptr_t p = reinterpret_cast<uintptr_t>(dst);
  const uintptr_t a = std::min(page_up(p), p + bytes), b = std::max(page_down(p + bytes), a);
  const char* sp = static_cast<const char*>(src);
  cudaError_t e = cudaSuccess;
  auto piece = [&](uintptr_t to, const char* from, size_t n) -> cudaError_t {
    if (n == 0) return cudaSuccess;
    if (c->bounce.cap == 0 && c->bounce.reserve(kBounceBytes) != cudaSuccess) cudaGetLastError();
    if (c->bounce.cap >= c->bounce_used + n) {
      const size_t off = c->bounce_used;
      c->bounce_used += (n + 15) & ~(size_t)15;
      c->fixups.push_back({reinterpret_cast<void*>(to), off, n});
      return cudaMemcpyAsync(c->bounce.p + off, from, n, cudaMemcpyDeviceToHost, s);
    }
    return cudaMemcpyAsync(reinterpret_cast<void*>(to), from, n, cudaMemcpyDeviceToHost, s);   // out of slots: pageable (blocks)
  };
  e = piece(p, sp, a - p);
  if (e == cudaSuccess && b > a) e = cudaMemcpyAsync(reinterpret_cast<void*>(a), sp + (a - p), b - a, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = piece(b, sp + (b - p), p + bytes - b);
  return e;
}

// After the streams of c have been synchronized: the partial pages wait in the bounce block.
void apply_fixups(tsq_ctx* c) {
  for (const tsq_ctx::Fixup& f : c->fixups) memcpy(f.dst, c->bounce.p + f.off, f.bytes);
  c->fixups.clear();
  c->bounce_used = 0;
}

void unregister_all(tsq_ctx* c) {
  for (void* r : c->ext_registered)
    if (cudaHostUnregister(r) != cudaSuccess) cudaGetLastError();
  c->ext_registered.clear();
}

// Host destinations of context c's results, as pointers indexed by ABSOLUTE packed index (so that a slab
// lands at its place whatever the buffer really spans).  Reserves / page-locks on first use.
struct HostDst {
  int32_t* scores = nullptr;
  int16_t* scores16 = nullptr;   // TSQ_FLAG_SCORES_I16: instead of scores
  int32_t* nid = nullptr;
  double* dist = nullptr;
};
int host_results(tsq_ctx* c, bool want_dist, bool want_nid, HostDst* out) {
  tsq_ctx* o = result_owner(c);
  uint64_t first = 0, count = 0;
  host_extent(o, &first, &count);   // a leader's buffers span the whole triangle
  if (o->ext_scores) {
    if (c->prm.flags & TSQ_FLAG_SCORES_I16) return fail(c, TSQ_ERR_INVALID, "TSQ_FLAG_SCORES_I16 delivers into the library's own buffer (tsq_scores16), not into tsq_set_result_buffers");
    if (o->ext_count < count) return fail(c, TSQ_ERR_INVALID, "result buffers hold %llu pairs, the job has %llu",
                                          (unsigned long long)o->ext_count, (unsigned long long)count);
    if (want_dist && !o->ext_dist) return fail(c, TSQ_ERR_INVALID, "tsq_set_result_buffers: no distance buffer, but distances are on");
    // page-lock what this process writes: a lone rank's slab, or everything
    uint64_t lb = 0, le = count;
    if (!c->leader && c->slab_mode) {
      lb = c->part_begin;
      le = c->part_end;
    }
    int rc = register_range(o, o->ext_scores + lb, (size_t)(le - lb) * sizeof(int32_t));
    if (rc == TSQ_OK && want_dist) rc = register_range(o, o->ext_dist + lb, (size_t)(le - lb) * sizeof(double));
    if (rc != TSQ_OK) {
      if (o != c) copy_error(c, o);
      return rc;
    }
    out->scores = o->ext_scores;
    out->dist = want_dist ? o->ext_dist : nullptr;
  } else {
    if (c->prm.flags & TSQ_FLAG_SCORES_I16) {
      TSQ_CUDA(c, o->h_scores16.reserve(count));
      out->scores16 = biased(o->h_scores16.p, first);
    } else {
      TSQ_CUDA(c, o->h_scores.reserve(count));
      out->scores = biased(o->h_scores.p, first);
    }
    if (want_dist) {
      TSQ_CUDA(c, o->h_dist.reserve(count));
      out->dist = biased(o->h_dist.p, first);
    }
  }
  if (want_nid) {
    TSQ_CUDA(c, o->h_nid.reserve(count));
    out->nid = biased(o->h_nid.p, first);
  }
  return TSQ_OK;
}

// ---- finalize --------------------------------------------------------------------------------------------
// Sorted rows [row_begin, row_end) of context c, read from ITS sorted buffer, written to the given packed
// triangles (absolute index; possibly another device's memory).
int launch_finalize(tsq_ctx* c, uint32_t row_begin, uint32_t row_end, int32_t* out_scores, int32_t* out_nid, double* out_dist) {
  tsq::FinalizeParams f{};
  f.sorted = sorted_base(c);
  f.lens = c->d_lens.p;
  f.perm = c->d_perm.p;
  f.self = c->d_self.p;
  f.out_scores = out_scores;
  f.out_nid = out_nid;
  f.out_dist = out_dist;
  f.idshift = c->idshift;
  f.n = c->n;
  f.row_begin = row_begin;
  f.row_end = row_end;
  f.go = c->go;
  f.ge = c->ge;
  f.perm_identity = c->perm_identity ? 1u : 0u;
  f.kimura = (c->prm.flags & TSQ_FLAG_KIMURA) ? 1u : 0u;
  f.kimura_oob = c->d_cancel + 2;
  if (f.kimura) TSQ_CUDA(c, cudaMemsetAsync(c->d_cancel + 2, 0, sizeof(int), c->stream));
  TSQ_CUDA(c, tsq::finalize_launch(f, c->stream));
  c->st.launches++;
  return TSQ_OK;
}

// The fault word a kernel sets when it gives up on a TMA barrier (tma_stage.cuh): read back with every
// synchronize so that a protocol failure is an error code, never a silently wrong score.
int check_device_fault(tsq_ctx* c) {
  if (!c->d_cancel || !c->h_fault) return TSQ_OK;
  c->h_fault[0] = c->h_fault[1] = 0;
  TSQ_CUDA(c, cudaMemcpyAsync(c->h_fault, c->d_cancel + 1, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  TSQ_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_fault[0] != 0) {
    c->computed = c->finalized = c->downloaded = false;
    return fail(c, TSQ_ERR_CUDA, "device fault 0x%x: a wavefront kernel timed out on a TMA tile barrier; results discarded", c->h_fault[0]);
  }
  if ((c->prm.flags & TSQ_FLAG_KIMURA) && c->finalized && c->h_fault[1] != 0) {
    c->finalized = c->downloaded = false;
    return fail(c, TSQ_ERR_RANGE, "%d pair(s) have an identity distance of 0.75 or more: Kimura's correction -ln(1 - D - D^2/5) does not "
                "apply there (ClustalW switches to a lookup table that cannot be restated); run without TSQ_FLAG_KIMURA", c->h_fault[1]);
  }
  return TSQ_OK;
}

// Behind one launch of a streamed compute: on the side stream, once that launch is done, finalize its rows
// (distances next to the scores) and copy both to their place in the host result -- while the next launch runs.
// The scores of packed indices [pb, pe) to the host result: as they are, or narrowed to int16 first
// (TSQ_FLAG_SCORES_I16).  src: the device's final int32 scores, src[0] = packed index `first`; ready16: the
// int16 copy is there already (the packed kernel wrote it next to the int32 one).
int scores_out(tsq_ctx* c, tsq_ctx* owner, const HostDst& h, uint64_t pb, uint64_t pe, const int32_t* src, uint64_t first,
               uint64_t extent, bool ready16, cudaStream_t s) {
  if (pe <= pb) return TSQ_OK;
  if (!(c->prm.flags & TSQ_FLAG_SCORES_I16)) {
    TSQ_CUDA(c, copy_out(c, owner, h.scores + pb, src + (pb - first), (pe - pb) * sizeof(int32_t), s));
    return TSQ_OK;
  }
  TSQ_CUDA(c, c->d_scores16.reserve(extent));
  if (!ready16) {
    TSQ_CUDA(c, tsq::narrow_scores_launch(src + (pb - first), c->d_scores16.p + (pb - first), pe - pb, s));
    c->st.launches++;
  }
  TSQ_CUDA(c, copy_out(c, owner, h.scores16 + pb, c->d_scores16.p + (pb - first), (pe - pb) * sizeof(int16_t), s));
  return TSQ_OK;
}

int stream_chunk_out(tsq_ctx* c, const StreamChunk& ch, cudaStream_t compute_stream, size_t index) {
  const uint64_t n = c->n, npairs = n < 2 ? 0 : n * (n - 1) / 2;
  auto start_of = [&](uint32_t row) -> uint64_t { return (n >= 2 && (uint64_t)row + 1 < n) ? tri(row, row + 1, n) : npairs; };
  const uint64_t pb = start_of(ch.row0), pe = start_of(ch.row1);
  TSQ_CUDA(c, cudaEventRecord(c->chunk_ev[index], compute_stream));
  if (pe <= pb) return TSQ_OK;
  const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
  HostDst h;
  int rc = host_results(c, want_dist, false, &h);
  if (rc != TSQ_OK) return rc;
  TSQ_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[index], 0));
  const uint64_t first = c->full_sorted ? 0 : c->part_begin;   // what the device buffers start at
  if (want_dist) {
    TSQ_CUDA(c, c->d_dist.reserve(c->full_sorted ? npairs : c->part_end - c->part_begin));
    tsq::FinalizeParams f{};
    f.sorted = sorted_base(c);
    f.lens = c->d_lens.p;
    f.perm = c->d_perm.p;
    f.self = c->d_self.p;
    f.out_scores = sorted_base(c);
    f.out_nid = nullptr;
    f.out_dist = biased(c->d_dist.p, first);
    f.idshift = 0;
    f.n = c->n;
    f.row_begin = ch.row0;
    f.row_end = ch.row1;
    f.go = c->go;
    f.ge = c->ge;
    f.perm_identity = 1u;
    f.kimura = 0;
    f.kimura_oob = c->d_cancel + 2;
    TSQ_CUDA(c, tsq::finalize_launch(f, c->copy_stream));
    c->st.launches++;
  }
  tsq_ctx* owner = result_owner(c);
  rc = scores_out(c, owner, h, pb, pe, c->d_sorted.p, first, c->full_sorted ? npairs : c->part_end - c->part_begin, false, c->copy_stream);
  if (rc != TSQ_OK) return rc;
  if (want_dist) TSQ_CUDA(c, copy_out(c, owner, h.dist + pb, c->d_dist.p + (pb - first), (pe - pb) * sizeof(double), c->copy_stream));
  c->streamed_bytes += (pe - pb) * (score_bytes(c) + (want_dist ? 8ull : 0ull));
  return TSQ_OK;
}

// Counter-driven flavour (one launch, enqueue_gotoh16): behind the event that says "counters zeroed", the copy stream
// waits for each row range's task count and copies the range -- scores and the distances the kernel wrote -- out.
int stream_ranges_out(tsq_ctx* c, const std::vector<StreamChunk>& ranges) {
  const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
  const uint64_t n = c->n, npairs = n < 2 ? 0 : n * (n - 1) / 2;
  const uint64_t first = c->full_sorted ? 0 : c->part_begin;
  const size_t nr = ranges.size();
  HostDst h;
  const int rch = host_results(c, want_dist, false, &h);
  if (rch != TSQ_OK) return rch;
  TSQ_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->chunk_ev[0], 0));   // counters zeroed (not: kernel finished)
  tsq_ctx* owner = result_owner(c);
  auto start_of = [&](uint32_t row) -> uint64_t { return (n >= 2 && (uint64_t)row + 1 < n) ? tri(row, row + 1, n) : npairs; };
  c->streamed_bytes = 0;
  for (size_t k = 0; k < nr; k++) {
    const StreamChunk& ch = ranges[k];
    const uint64_t pb = start_of(ch.row0), pe = start_of(ch.row1);
    if (ch.t1 > ch.t0) {
      const CUresult wr = stream_wait_value32()((CUstream)c->copy_stream, (CUdeviceptr)(uintptr_t)(c->d_done.p + k),
                                                (cuuint32_t)(ch.t1 - ch.t0), CU_STREAM_WAIT_VALUE_GEQ);
      if (wr != CUDA_SUCCESS) return fail(c, TSQ_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)wr);
    }
    if (pe <= pb) continue;
    const int rco = scores_out(c, owner, h, pb, pe, c->d_sorted.p, first, c->full_sorted ? npairs : c->part_end - c->part_begin, true, c->copy_stream);
    if (rco != TSQ_OK) return rco;
    if (want_dist) TSQ_CUDA(c, copy_out(c, owner, h.dist + pb, c->d_dist.p + (pb - first), (pe - pb) * sizeof(double), c->copy_stream));
    c->streamed_bytes += (pe - pb) * (score_bytes(c) + (want_dist ? 8ull : 0ull));
  }
  return TSQ_OK;
}

// ---- multi-device leader: every entry point fans out to the per-device children ----------------------------
template <typename F>
int for_each_kid_parallel(tsq_ctx* c, F&& body) {
  const size_t nk = c->kids.size();
  if (!c->pool) c->pool = new KidPool(c->kids);
  c->pool->run(std::function<int(tsq_ctx*)>(body));
  for (size_t r = 0; r < nk; r++)
    if (c->pool->rc[r] != TSQ_OK) {
      c->err = "device " + std::to_string(c->kids[r]->device) + ": " + c->kids[r]->err;
      return c->pool->rc[r];
    }
  return TSQ_OK;
}

void multi_share_sequences(tsq_ctx* c) {
  for (tsq_ctx* k : c->kids) {
    k->encp = &c->enc;
    k->selfp = &c->self_input;
    k->n = c->n;
    k->have_seqs = true;
    k->uploaded = k->computed = k->finalized = k->downloaded = false;
    k->err.clear();
  }
}

int multi_upload(tsq_ctx* c) {
  const double t0 = now_ms();
  // one host thread per device: each sorts, plans ITS rows, packs and copies over its own link
  int rc = for_each_kid_parallel(c, [](tsq_ctx* k) { return tsq_upload(k); });
  if (rc != TSQ_OK) return rc;
  const tsq_ctx* k0 = c->kids[0];
  c->perm_identity = k0->perm_identity;
  c->idshift = k0->idshift;
  c->slab_mode = k0->slab_mode;
  c->uploaded = true;
  c->computed = c->finalized = c->downloaded = false;
  c->st.upload_ms = now_ms() - t0;
  return TSQ_OK;
}

int multi_compute(tsq_ctx* c) {
  // the host result must exist before the devices start streaming into it (one owner, several writers: reserve or
  // page-lock it here, on one thread)
  tsq_ctx* k0 = c->kids[0];
  if (k0->stream_out && c->slab_mode && c->idshift == 0 && c->n >= 2) {
    HostDst h;
    const int rc = host_results(k0, !(c->prm.flags & TSQ_FLAG_NO_DISTANCES), false, &h);
    if (rc != TSQ_OK) {
      copy_error(c, k0);
      return rc;
    }
  }
  // every device's launches (and, when streamed, its finalize and copy-out enqueues) from its own host thread
  const int rc = for_each_kid_parallel(c, [](tsq_ctx* k) { return tsq_compute(k); });
  if (rc != TSQ_OK) return rc;
  c->computed = true;
  c->finalized = c->downloaded = false;
  return TSQ_OK;
}

int multi_synchronize(tsq_ctx* c) {
  double worst = 0;
  for (tsq_ctx* k : c->kids) {
    const int rc = tsq_synchronize(k);
    if (rc != TSQ_OK) {
      copy_error(c, k);
      if (rc == TSQ_ERR_CUDA) c->computed = c->finalized = c->downloaded = false;
      return rc;
    }
    worst = std::max(worst, k->st.kernel_ms);
  }
  c->st.kernel_ms = worst;
  return TSQ_OK;
}

int multi_finalize(tsq_ctx* c) {
  const uint32_t n = c->n;
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
  tsq_ctx* k0 = c->kids[0];
  if (c->slab_mode) {
    // fixed-length input: a device's slab is a contiguous piece of the final triangle -- finalize in place
    // (children that streamed their rows out behind the launches have done it already)
    for (tsq_ctx* k : c->kids) {
      const int rc = tsq_finalize(k);
      if (rc != TSQ_OK) {
        copy_error(c, k);
        return rc;
      }
    }
  } else if (npairs > 0) {
    // the un-sort scatters every slab over the triangle: each device un-sorts its rows STRAIGHT INTO device
    // 0's matrix with peer stores over NVLink (finalize + gather in one kernel, no staging copy)
    TSQ_CUDA(c, cudaSetDevice(k0->device));
    TSQ_CUDA(c, k0->d_scores.reserve(npairs));
    if (c->idshift) TSQ_CUDA(c, k0->d_nid.reserve(npairs));
    if (want_dist) TSQ_CUDA(c, k0->d_dist.reserve(npairs));
    for (tsq_ctx* k : c->kids) {
      if (k->device != k0->device) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, k->device, k0->device) != cudaSuccess || !can) {
          cudaGetLastError();
          return fail(c, TSQ_ERR_CUDA, "device %d cannot reach device %d's memory (no NVLink/PCIe peer access): "
                      "ragged input on several devices needs it", k->device, k0->device);
        }
      }
      TSQ_CUDA(c, cudaSetDevice(k->device));
      const int rc = launch_finalize(k, k->row_a, k->row_b, k0->d_scores.p, c->idshift ? k0->d_nid.p : nullptr,
                                     want_dist ? k0->d_dist.p : nullptr);
      if (rc != TSQ_OK) {
        copy_error(c, k);
        return rc;
      }
      TSQ_CUDA(c, cudaEventRecord(k->fin_ev, k->stream));
      k->finalized = true;
    }
    TSQ_CUDA(c, cudaSetDevice(k0->device));
    for (tsq_ctx* k : c->kids)
      if (k != k0) TSQ_CUDA(c, cudaStreamWaitEvent(k0->stream, k->fin_ev, 0));
  }
  for (tsq_ctx* k : c->kids) {
    k->finalized = true;
    k->have_tree = k->have_msa = false;
  }
  k0->d_dist_full.release();
  c->finalized = true;
  return TSQ_OK;
}

int multi_download(tsq_ctx* c) {
  if (!c->finalized) {
    const int rc = multi_finalize(c);
    if (rc != TSQ_OK) return rc;
  }
  const double t0 = now_ms();
  const uint32_t n = c->n;
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
  uint64_t bytes = 0;
  if (npairs > 0) {
    if (c->slab_mode) {
      // every device copies ITS slab over ITS OWN PCIe link into the one host result
      for (tsq_ctx* k : c->kids) {
        const int rc = tsq_download(k);   // enqueues; see the slab branch there
        if (rc != TSQ_OK) {
          copy_error(c, k);
          return rc;
        }
      }
    } else {
      tsq_ctx* k0 = c->kids[0];
      HostDst h;
      int rc = host_results(k0, want_dist, c->idshift != 0, &h);
      if (rc != TSQ_OK) {
        copy_error(c, k0);
        return rc;
      }
      TSQ_CUDA(c, cudaSetDevice(k0->device));
      {
        const int rco = scores_out(k0, c, h, 0, npairs, k0->d_scores.p, 0, npairs, false, k0->stream);
        if (rco != TSQ_OK) {
          copy_error(c, k0);
          return rco;
        }
      }
      if (c->idshift) TSQ_CUDA(c, cudaMemcpyAsync(h.nid, k0->d_nid.p, npairs * sizeof(int32_t), cudaMemcpyDeviceToHost, k0->stream));
      if (want_dist) TSQ_CUDA(c, copy_out(k0, c, h.dist, k0->d_dist.p, npairs * sizeof(double), k0->stream));
    }
    bytes = npairs * (want_dist ? 12ull : 4ull) + (c->idshift ? npairs * 4ull : 0ull);
  }
  int rc = multi_synchronize(c);
  if (rc != TSQ_OK) return rc;
  for (tsq_ctx* k : c->kids) k->downloaded = true;
  c->downloaded = true;
  c->st.download_ms = now_ms() - t0;
  c->st.d2h_bytes = bytes;
  return TSQ_OK;
}

// Guide tree, alignment, traceback and consensus of a multi-device context run on its first device, which
// holds the whole database; the distance matrix is assembled there from the slabs with peer copies (NVLink).
int multi_prepare_first_device(tsq_ctx* c) {
  if (!c->finalized) return fail(c, TSQ_ERR_STATE, "no results yet: call tsq_run first");
  tsq_ctx* k0 = c->kids[0];
  const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
  const uint32_t n = c->n;
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  if (c->slab_mode && want_dist && npairs > 0 && !k0->d_dist_full.p) {
    TSQ_CUDA(c, cudaSetDevice(k0->device));
    TSQ_CUDA(c, k0->d_dist_full.reserve(npairs));
    for (tsq_ctx* k : c->kids) {
      const uint64_t cnt = k->part_end - k->part_begin;
      if (cnt == 0) continue;
      TSQ_CUDA(c, cudaStreamWaitEvent(k0->stream, k->fin_ev, 0));   // that device's distances are complete
      TSQ_CUDA(c, cudaMemcpyPeerAsync(k0->d_dist_full.p + k->part_begin, k0->device, k->d_dist.p, k->device,
                                      cnt * sizeof(double), k0->stream));
    }
    TSQ_CUDA(c, cudaStreamSynchronize(k0->stream));
  }
  return TSQ_OK;
}

}  // namespace

extern "C" {

int tsq_detect_alphabet(const char* const* residues, const uint32_t* lengths, uint32_t n) {
  // class of a byte: 1 = letter, 3 = letter of ACGTUN; counted per byte value first (one table look-up per residue)
  static const struct Cls {
    uint8_t v[256];
    Cls() {
      for (int b = 0; b < 256; b++) {
        const bool letter = (b >= 'A' && b <= 'Z') || (b >= 'a' && b <= 'z');
        v[b] = !letter ? 0 : (b != 0 && strchr("ACGTUNacgtun", b)) ? 3 : 1;
      }
    }
  } cls;
  unsigned long long letters = 0, nuc = 0;
  for (uint32_t i = 0; i < n && residues && lengths; i++) {
    if (!residues[i]) continue;
    const unsigned char* r = reinterpret_cast<const unsigned char*>(residues[i]);
    unsigned long long l = 0, u = 0;
    for (uint32_t k = 0; k < lengths[i]; k++) {
      const unsigned v = cls.v[r[k]];
      l += v & 1u;
      u += v >> 1;
    }
    letters += l;
    nuc += u;
  }
  return (letters > 0 && nuc * 10 >= letters * 9) ? TSQ_NUCLEOTIDE : TSQ_PROTEIN;
}

int tsq_encode(int alphabet, const char* residues, uint64_t len, uint8_t* out, uint64_t* out_len, int64_t* self_score) {
  if ((alphabet != TSQ_PROTEIN && alphabet != TSQ_NUCLEOTIDE) || !out_len || (len > 0 && (!residues || !out))) return TSQ_ERR_INVALID;
  static const tsq::EncodeTables* const tabs = [] {   // default matrices; built once, thread-safe
    static tsq::EncodeTables t[2];
    for (int a = 0; a < 2; a++) {
      const uint8_t* lut = a == TSQ_NUCLEOTIDE ? kLut.nuc : kLut.prot;
      const int nsym = a == TSQ_NUCLEOTIDE ? 5 : 23;
      const int8_t* m = a == TSQ_NUCLEOTIDE ? kDna : kBlosum62;
      for (int b = 0; b < 256; b++) {
        t[a].lut[b] = lut[b];
        t[a].diag_by_byte[b] = lut[b] == 0xff ? 0 : m[lut[b] * (nsym + 1)];
      }
      tsq::encode_tables_finish(&t[a]);
    }
    return t;
  }();
  int64_t self = 0;
  // (stores of the vector path stay inside out[0, len): a block is written at or before where it was read)
  const bool scalar = getenv("TSQ_ENCODE_SCALAR") != nullptr;   // test hook: the table loop
  *out_len = scalar ? tsq::encode_residues_scalar(tabs[alphabet], residues, (size_t)len, out, &self)
                    : tsq::encode_residues(tabs[alphabet], residues, (size_t)len, out, &self);
  if (self_score) *self_score = self;
  return TSQ_OK;
}

int tsq_create(tsq_ctx** out, const tsq_params* params) {
  if (!out) return TSQ_ERR_INVALID;
  *out = nullptr;
  tsq_params p;
  tsq_default_params(&p);
  if (params) {
    if (params->struct_size < 8 || params->struct_size > sizeof(tsq_params)) return TSQ_ERR_INVALID;
    memcpy(&p, params, params->struct_size);
    p.struct_size = (uint32_t)sizeof(tsq_params);
  }
  if (p.alphabet != TSQ_PROTEIN && p.alphabet != TSQ_NUCLEOTIDE) return TSQ_ERR_INVALID;
  if ((p.flags & TSQ_FLAG_KIMURA) && !(p.flags & TSQ_FLAG_IDENTITY)) return TSQ_ERR_INVALID;   // corrects the identity distance
  if ((p.flags & TSQ_FLAG_SCORES_I16) && (p.flags & TSQ_FLAG_IDENTITY)) return TSQ_ERR_INVALID;   // identity keys are decoded into int32 arrays
  if (p.part_world < 1) p.part_world = 1;
  if (p.part_rank < 0 || p.part_rank >= p.part_world) return TSQ_ERR_INVALID;
  const int nsym = p.alphabet == TSQ_NUCLEOTIDE ? 5 : 23;
  const int8_t* m = p.matrix ? p.matrix : (p.alphabet == TSQ_NUCLEOTIDE ? kDna : kBlosum62);
  for (int a = 0; a < nsym; a++)
    for (int b = 0; b < nsym; b++)
      if (m[a * nsym + b] != m[b * nsym + a]) return TSQ_ERR_MATRIX;
  const int go = p.gap_open < 0 ? (p.alphabet == TSQ_NUCLEOTIDE ? 10 : 11) : p.gap_open;
  const int ge = p.gap_extend < 0 ? 1 : p.gap_extend;
  if (go > 4096 || ge > 1024) return TSQ_ERR_INVALID;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return TSQ_ERR_NO_DEVICE;
  }
  if (p.device < 0 || p.device >= ndev) return TSQ_ERR_NO_DEVICE;
  if (!device_usable(p.device)) return TSQ_ERR_NO_DEVICE;
  // ---- several devices behind one context ----------------------------------------------------------------
  if (p.n_devices < 0) {   // all usable devices from `device` on
    int cnt = 0;
    for (int d = p.device; d < ndev && device_usable(d); d++) cnt++;
    p.n_devices = cnt;
  }
  if (p.n_devices == 0) p.n_devices = 1;
  // test hook: all children on the one device, so that the multi-device host logic (slab finalize, peer-store
  // finalize, own-link downloads, collected distances) also runs on a one-GPU box
  const bool same_device = p.n_devices > 1 && getenv("TSQ_MULTI_SAME_DEVICE") != nullptr;
  if (p.n_devices > 1) {
    if (p.part_world != 1) return TSQ_ERR_INVALID;          // one way of partitioning at a time
    if (!same_device) {
      if (p.device + p.n_devices > ndev) return TSQ_ERR_NO_DEVICE;
      for (int d = p.device; d < p.device + p.n_devices; d++)
        if (!device_usable(d)) return TSQ_ERR_NO_DEVICE;
    }
  }
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p.device);

  tsq_ctx* c = new (std::nothrow) tsq_ctx();
  if (!c) return TSQ_ERR_NOMEM;
  c->prm = p;
  c->nsym = nsym;
  c->matrix.assign(m, m + nsym * nsym);
  c->prm.matrix = nullptr;
  c->go = go;
  c->ge = ge;
  c->device = p.device;
  c->sm_count = sms;
  c->encp = &c->enc;
  c->selfp = &c->self_input;
  c->smin = *std::min_element(c->matrix.begin(), c->matrix.end());
  c->smax = *std::max_element(c->matrix.begin(), c->matrix.end());
  c->delta = c->smin < 0 ? (-c->smin + 1) / 2 : 0;
  {
    const uint8_t* lut = p.alphabet == TSQ_NUCLEOTIDE ? kLut.nuc : kLut.prot;
    for (int b = 0; b < 256; b++) {
      c->enct.lut[b] = lut[b];
      c->enct.diag_by_byte[b] = lut[b] == 0xff ? 0 : c->matrix[lut[b] * (nsym + 1)];
    }
    tsq::encode_tables_finish(&c->enct);
  }
  c->max_len16 = (p.flags & TSQ_FLAG_FORCE_S32) ? 0 : max_len16_of(c);
  if (p.n_devices > 1) {
    // the leader owns no device state: one child context per device does the work
    for (int r = 0; r < p.n_devices; r++) {
      tsq_params kp = p;
      kp.matrix = c->matrix.data();
      kp.device = same_device ? p.device : p.device + r;
      kp.n_devices = 1;
      kp.part_rank = r;
      kp.part_world = p.n_devices;
      tsq_ctx* k = nullptr;
      const int rc = tsq_create(&k, &kp);
      if (rc != TSQ_OK) {
        tsq_destroy(c);
        return rc;
      }
      k->leader = c;
      c->kids.push_back(k);
    }
    // peer access towards the first device (ragged input: every device un-sorts its rows into device 0's
    // matrix) and back (the distance slabs collected for the guide tree); absence is an error only when used
    for (int r = 1; r < p.n_devices && !same_device; r++) {
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, p.device + r, p.device) == cudaSuccess && can) {
        cudaSetDevice(p.device + r);
        if (cudaDeviceEnablePeerAccess(p.device, 0) != cudaSuccess) cudaGetLastError();
        cudaSetDevice(p.device);
        if (cudaDeviceEnablePeerAccess(p.device + r, 0) != cudaSuccess) cudaGetLastError();
      } else {
        cudaGetLastError();
      }
    }
    *out = c;
    return TSQ_OK;
  }
  if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      cudaEventCreate(&c->tev0) != cudaSuccess || cudaEventCreate(&c->tev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->fin_ev, cudaEventDisableTiming) != cudaSuccess) {
    cudaGetLastError();
    tsq_destroy(c);
    return TSQ_ERR_CUDA;
  }
  if (BlockCache::get().take(c->device, 4 * sizeof(int), (void**)&c->d_cancel) != cudaSuccess ||
      cudaMemset(c->d_cancel, 0, 4 * sizeof(int)) != cudaSuccess ||
      BlockCache::get().take(-1, 4 * sizeof(int), (void**)&c->h_one) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->cancel_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    tsq_destroy(c);
    return TSQ_ERR_CUDA;
  }
  c->h_one[0] = 1;
  c->h_one[1] = c->h_one[2] = c->h_one[3] = 0;
  c->h_fault = c->h_one + 1;
  c->stream = c->own_stream;
  *out = c;
  return TSQ_OK;
}

int tsq_destroy(tsq_ctx* c) {
  if (!c) return TSQ_OK;
  if (c->pool) {
    delete c->pool;
    c->pool = nullptr;
  }
  for (tsq_ctx* k : c->kids) tsq_destroy(k);
  c->kids.clear();
  cudaSetDevice(c->device);
  if (c->own_stream) cudaStreamSynchronize(c->own_stream);
  if (c->stream && c->stream != c->own_stream && cudaStreamSynchronize(c->stream) != cudaSuccess) cudaGetLastError();
  if (c->copy_stream && cudaStreamSynchronize(c->copy_stream) != cudaSuccess) cudaGetLastError();
  unregister_all(c);
  c->d_dbw.release(); c->d_blob.release(); c->h_blob.release();
  c->d_sorted.release(); c->d_scores.release(); c->d_nid.release(); c->h_nid.release(); c->d_dist.release(); c->d_dist_full.release();
  c->d_counter.release(); c->d_done.release(); c->d_bnd.release(); c->h_scores.release(); c->h_scores16.release(); c->d_scores16.release(); c->h_dist.release(); c->bounce.release();
  c->d_treeD.release(); c->d_treemin.release(); c->d_treeh.release(); c->d_treeu.release(); c->d_merges.release();
  c->d_pairs32.release(); c->d_tasks16w.release(); c->d_bnd16w.release(); c->d_bnd32.release(); c->d_smat.release();
  if (c->d_cancel) BlockCache::get().give(c->device, c->d_cancel);
  if (c->h_one) BlockCache::get().give(-1, c->h_one);
  if (c->copy_stream) {
    cudaStreamSynchronize(c->copy_stream);
    cudaStreamDestroy(c->copy_stream);
  }
  for (cudaEvent_t e : c->chunk_ev)
    if (e) cudaEventDestroy(e);
  if (c->h_starts) BlockCache::get().give(-1, c->h_starts);
  if (c->cancel_stream) cudaStreamDestroy(c->cancel_stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->tev0) cudaEventDestroy(c->tev0);
  if (c->tev1) cudaEventDestroy(c->tev1);
  if (c->fin_ev) cudaEventDestroy(c->fin_ev);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
  return TSQ_OK;
}

const char* tsq_last_error(const tsq_ctx* c) { return c ? c->err.c_str() : "null context"; }

int tsq_set_stream(tsq_ctx* c, void* s) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) return fail(c, TSQ_ERR_INVALID, "a multi-device context runs on its own per-device streams");
  cudaStream_t ns = s ? (cudaStream_t)s : c->own_stream;
  if (ns != c->stream && c->stream) {   // work queued on the old stream (an upload's copies) is not ordered before the new one
    TSQ_CUDA(c, cudaSetDevice(c->device));
    TSQ_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  c->stream = ns;
  return TSQ_OK;
}

}  // extern "C"

namespace {

// Encodes all sequences (gap stripping, letter map, self scores).  Large inputs are cut into byte-balanced ranges
// of sequences for a few host threads: at 10^5 sequences this pass is the longest host step of a job, and in a
// weak-scaled multi-GPU run every rank repeats it for the whole input.
template <typename Get>
void encode_all(tsq_ctx* c, uint32_t n, uint64_t total_bytes, Get&& get) {
  auto work = [&](uint32_t a, uint32_t b) {
    std::vector<tsq::EncodeJob> jobs(b - a);
    for (uint32_t i = a; i < b; i++) {
      const char* p = nullptr;
      size_t len = 0;
      get(i, &p, &len);
      if (c->enc[i].size() < len) c->enc[i].resize(len);
      jobs[i - a] = tsq::EncodeJob{p, len, c->enc[i].data(), 0, 0};
    }
    tsq::encode_many(c->enct, jobs.data(), jobs.size());   // gap stripping, letter map, S(x, x): 32 bytes per step
    for (uint32_t i = a; i < b; i++) {
      c->enc[i].resize(jobs[i - a].out_len);
      c->self_input[i] = (int32_t)jobs[i - a].self;
    }
  };
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  unsigned nt = total_bytes >= (2u << 20) ? std::min<unsigned>({4u, hw, n}) : 1u;   // ~0.1 ms per MB on one thread
  if (nt <= 1) {
    work(0, n);
    return;
  }
  std::vector<uint32_t> cut(nt + 1, n);
  cut[0] = 0;
  {
    uint64_t acc = 0;
    unsigned k = 1;
    for (uint32_t i = 0; i < n && k < nt; i++) {
      const char* p = nullptr;
      size_t len = 0;
      get(i, &p, &len);
      acc += len;
      if (acc >= total_bytes * k / nt) cut[k++] = i + 1;
    }
  }
  std::vector<std::thread> th;
  for (unsigned k = 1; k < nt; k++) th.emplace_back(work, cut[k], cut[k + 1]);
  work(cut[0], cut[1]);
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

int tsq_set_sequences(tsq_ctx* c, const char* const* residues, const uint32_t* lengths, uint32_t n) {
  if (!c) return TSQ_ERR_INVALID;
  if (n > 0 && (!residues || !lengths)) return fail(c, TSQ_ERR_INVALID, "null sequence arrays");
  const double t0 = now_ms();
  // a failed call leaves the context without sequences rather than with half of the new set
  c->have_seqs = c->uploaded = c->computed = c->finalized = c->downloaded = false;
  c->n = 0;
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (lengths[i] > 0 && !residues[i]) return fail(c, TSQ_ERR_INVALID, "sequence %u is null", i);
    total += lengths[i];
  }
  c->enc.resize(n);
  c->self_input.resize(n);
  encode_all(c, n, total, [&](uint32_t i, const char** p, size_t* len) {
    *p = residues[i];
    *len = lengths[i];
  });
  c->n = n;
  c->have_seqs = true;
  c->uploaded = c->computed = c->finalized = c->downloaded = false;
  c->err.clear();
  multi_share_sequences(c);   // children of a multi-device context read the leader's encoding
  c->st.encode_ms = now_ms() - t0;
  return TSQ_OK;
}

int tsq_set_sequences_flat(tsq_ctx* c, const char* residues, const uint64_t* offsets, uint32_t n) {
  if (!c) return TSQ_ERR_INVALID;
  if (n > 0 && (!residues || !offsets)) return fail(c, TSQ_ERR_INVALID, "null sequence buffer");
  const double t0 = now_ms();
  c->have_seqs = c->uploaded = c->computed = c->finalized = c->downloaded = false;
  c->n = 0;
  for (uint32_t i = 0; i < n; i++)
    if (offsets[i + 1] < offsets[i]) return fail(c, TSQ_ERR_INVALID, "offsets not ascending at %u", i);
  c->enc.resize(n);
  c->self_input.resize(n);
  encode_all(c, n, n ? offsets[n] - offsets[0] : 0, [&](uint32_t i, const char** p, size_t* len) {
    *p = residues + offsets[i];
    *len = (size_t)(offsets[i + 1] - offsets[i]);
  });
  c->n = n;
  c->have_seqs = true;
  c->uploaded = c->computed = c->finalized = c->downloaded = false;
  c->err.clear();
  multi_share_sequences(c);
  c->st.encode_ms = now_ms() - t0;
  return TSQ_OK;
}

int tsq_upload(tsq_ctx* c) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->have_seqs) return fail(c, TSQ_ERR_STATE, "tsq_upload before tsq_set_sequences");
  if (!c->kids.empty()) return multi_upload(c);
  const double t0 = now_ms();
  TSQ_CUDA(c, cudaSetDevice(c->device));
  // an earlier job may still be reading the staging blob or the device buffers this call re-uses (or returns to the
  // block cache when they grow): wait for it; free when the stream is idle
  TSQ_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->copy_stream) {   // ... and the side stream of a streamed compute nobody waited for (it reads the buffers re-sized below)
    TSQ_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    apply_fixups(c);
  }
  int rc = host_sort_and_pack(c);
  if (rc == TSQ_OK) rc = host_plan_work(c);
  if (rc == TSQ_OK) rc = host_build_subject_db(c);
  if (rc == TSQ_OK) rc = device_upload(c);
  if (rc != TSQ_OK) return rc;
  c->uploaded = true;
  c->computed = c->finalized = c->downloaded = false;
  c->st.upload_ms = now_ms() - t0;
  return TSQ_OK;
}

int tsq_compute(tsq_ctx* c) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_compute before tsq_upload");
  if (!c->kids.empty()) return multi_compute(c);
  TSQ_CUDA(c, cudaSetDevice(c->device));
  cudaStream_t s = c->stream;
  uint32_t launches = 0;
  c->st.launches = 0;
  c->streamed = false;
  if (!c->fixups.empty()) {   // an earlier compute streamed into caller memory and was never waited for
    const int rcs = tsq_synchronize(c);
    if (rcs != TSQ_OK) return rcs;
  }
  TSQ_CUDA(c, cudaMemsetAsync(c->d_cancel, 0, 4 * sizeof(int), s));   // cancel flag, fault word, Kimura out-of-range count
  c->ev0_set = false;
  int rc = enqueue_gotoh16(c, s, launches);                 // regime 1: packed inter-task kernel
  if (rc == TSQ_OK) rc = enqueue_wave16(c, s, launches);    // regime 2: packed wavefront kernel
  if (rc == TSQ_OK) rc = enqueue_wave32(c, s, launches);    // regime 2 fallback: 32-bit wavefront kernel
  if (rc != TSQ_OK) return rc;
  TSQ_MARK_START(c, s);   // (a job without a single task)
  if (!c->streamed) TSQ_CUDA(c, cudaEventRecord(c->ev1, s));
  c->st.launches += launches;
  c->computed = true;
  c->finalized = c->downloaded = false;
  if (c->streamed) {   // rows finalized and on their way to the host: what tsq_finalize / tsq_download would have done
    TSQ_CUDA(c, cudaEventRecord(c->fin_ev, c->copy_stream));
    c->finalized = true;
    c->have_tree = c->have_msa = false;
  }
  return TSQ_OK;
}

int tsq_finalize(tsq_ctx* c) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->computed) return fail(c, TSQ_ERR_STATE, "tsq_finalize before tsq_compute");
  if (!c->kids.empty()) return multi_finalize(c);
  TSQ_CUDA(c, cudaSetDevice(c->device));
  if (c->streamed) {   // done chunk by chunk behind the launches; whatever follows on the main stream sees the result
    TSQ_CUDA(c, cudaStreamWaitEvent(c->stream, c->fin_ev, 0));
    c->finalized = true;
    return TSQ_OK;
  }
  const uint32_t n = c->n;
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
  if (c->slab_mode) {
    // this rank's slab is a contiguous piece of the final triangle: distances next to it, scores in place
    const uint64_t cnt = c->part_end - c->part_begin;
    if (cnt > 0 && want_dist) {
      TSQ_CUDA(c, c->d_dist.reserve(cnt));
      const int rc = launch_finalize(c, c->row_a, c->row_b, sorted_base(c), nullptr, biased(c->d_dist.p, c->part_begin));
      if (rc != TSQ_OK) return rc;
    }
    if (c->fin_ev) TSQ_CUDA(c, cudaEventRecord(c->fin_ev, c->stream));
  } else if (c->prm.part_world > 1 && (c->prm.part_rank != 0 || c->leader)) {
    // a rank whose slab is gathered elsewhere (rank 0 / the leader's peer-store finalize) has nothing to do
    return fail(c, TSQ_ERR_STATE, "rank %d holds a slab only: gather it into rank 0 (tsq_device_slab) and finalize there", c->prm.part_rank);
  } else if (npairs > 0 && (want_dist || !c->perm_identity || c->idshift)) {
    const bool inplace = c->perm_identity && c->idshift == 0;   // identity keys are decoded into a separate buffer
    if (!inplace) TSQ_CUDA(c, c->d_scores.reserve(npairs));
    if (c->idshift) TSQ_CUDA(c, c->d_nid.reserve(npairs));
    if (want_dist) TSQ_CUDA(c, c->d_dist.reserve(npairs));
    const int rc = launch_finalize(c, 0, n, inplace ? c->d_sorted.p : c->d_scores.p, c->idshift ? c->d_nid.p : nullptr,
                                   want_dist ? c->d_dist.p : nullptr);
    if (rc != TSQ_OK) return rc;
  }
  c->finalized = true;
  c->have_tree = false;
  c->have_msa = false;
  return TSQ_OK;
}

int tsq_synchronize(tsq_ctx* c) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) return multi_synchronize(c);
  TSQ_CUDA(c, cudaSetDevice(c->device));
  TSQ_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->copy_stream) TSQ_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  apply_fixups(c);
  if (c->computed) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) == cudaSuccess) c->st.kernel_ms = ms;
    else cudaGetLastError();
    return check_device_fault(c);
  }
  return TSQ_OK;
}

int tsq_download(tsq_ctx* c) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->computed) return fail(c, TSQ_ERR_STATE, "tsq_download before tsq_compute");
  if (!c->kids.empty()) return multi_download(c);
  if (c->streamed) {   // every row range was copied behind its launch: only the wait is left
    const double t0s = now_ms();
    c->st.d2h_bytes = c->streamed_bytes;
    if (c->leader) return TSQ_OK;   // the leader synchronizes all its devices
    const int rc = tsq_synchronize(c);
    if (rc != TSQ_OK) return rc;
    c->downloaded = true;
    c->st.download_ms = now_ms() - t0s;
    return TSQ_OK;
  }
  const bool gathered_elsewhere = c->prm.part_world > 1 && !c->slab_mode && (c->prm.part_rank != 0 || c->leader);
  if (!c->finalized && !gathered_elsewhere) {
    int rc = tsq_finalize(c);
    if (rc != TSQ_OK) return rc;
  }
  const double t0 = now_ms();
  TSQ_CUDA(c, cudaSetDevice(c->device));
  const uint32_t n = c->n;
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  const bool want_dist = !(c->prm.flags & TSQ_FLAG_NO_DISTANCES);
  cudaStream_t s = c->stream;
  uint64_t bytes = 0;
  if (c->slab_mode) {
    // this rank's slab goes over this device's own PCIe link to its place in the host result
    const uint64_t cnt = c->part_end - c->part_begin;
    if (cnt > 0) {
      HostDst h;
      int rc = host_results(c, want_dist, false, &h);
      if (rc != TSQ_OK) return rc;
      rc = scores_out(c, result_owner(c), h, c->part_begin, c->part_end, c->d_sorted.p, c->part_begin, cnt, false, s);
      if (rc != TSQ_OK) return rc;
      if (want_dist) TSQ_CUDA(c, copy_out(c, result_owner(c), h.dist + c->part_begin, c->d_dist.p, cnt * sizeof(double), s));
      bytes = cnt * (score_bytes(c) + (want_dist ? 8ull : 0ull));
    }
    if (c->leader) {   // the leader synchronizes all its devices once every copy is in flight
      c->st.d2h_bytes = bytes;
      return TSQ_OK;
    }
  } else if (npairs > 0 && c->finalized) {
    HostDst h;
    int rc = host_results(c, want_dist, c->idshift != 0, &h);
    if (rc != TSQ_OK) return rc;
    const int32_t* src = (c->perm_identity && c->idshift == 0) ? c->d_sorted.p : c->d_scores.p;
    rc = scores_out(c, c, h, 0, npairs, src, 0, npairs, false, s);
    if (rc != TSQ_OK) return rc;
    if (c->idshift) TSQ_CUDA(c, cudaMemcpyAsync(h.nid, c->d_nid.p, npairs * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (want_dist) TSQ_CUDA(c, copy_out(c, c, h.dist, c->d_dist.p, npairs * sizeof(double), s));
    bytes = npairs * (score_bytes(c) + (want_dist ? 8ull : 0ull)) + (c->idshift ? npairs * 4ull : 0ull);
  }
  int rc = tsq_synchronize(c);
  if (rc != TSQ_OK) return rc;
  c->downloaded = c->finalized;
  c->st.download_ms = now_ms() - t0;
  c->st.d2h_bytes = bytes;
  return TSQ_OK;
}

int tsq_run(tsq_ctx* c, tsq_progress_cb cb, void* user, volatile int* cancel) {
  if (!c) return TSQ_ERR_INVALID;
  auto cancelled = [&]() { return cancel && *cancel != 0; };
  if (cancelled()) return fail(c, TSQ_ERR_CANCELLED, "cancelled");
  if (cb) cb(user, 0.0, "packing sequences");
  int rc = tsq_upload(c);
  if (rc != TSQ_OK) return rc;
  if (cancelled()) return fail(c, TSQ_ERR_CANCELLED, "cancelled");
  if (cb) cb(user, 0.05, "computing pairwise scores");
  // the blocking call knows the results are wanted on the host: let finished row ranges leave while the rest computes
  const bool was_streaming = c->stream_out;
  tsq_stream_results(c, 1);
  rc = tsq_compute(c);
  tsq_stream_results(c, was_streaming ? 1 : 0);
  if (rc != TSQ_OK) return rc;
  // poll the stream(s) so that "Stop" (SeqEditMainWin.cpp:803-812) is honoured while kernels run
  std::vector<tsq_ctx*> devs;
  if (c->kids.empty()) devs.push_back(c);
  else devs = c->kids;
  for (;;) {
    bool busy = false;
    for (tsq_ctx* k : devs) {
      cudaError_t q = cudaStreamQuery(k->stream);
      if (q == cudaErrorNotReady) busy = true;
      else if (q != cudaSuccess) return fail(c, TSQ_ERR_CUDA, "kernel failed on device %d: %s", k->device, cudaGetErrorString(q));
    }
    if (!busy) break;
    if (cancelled()) {
      // the kernels poll the flag at every task fetch and drain within one task
      for (tsq_ctx* k : devs) cudaMemcpyAsync(k->d_cancel, k->h_one, sizeof(int), cudaMemcpyHostToDevice, k->cancel_stream);
      for (tsq_ctx* k : devs) {
        cudaStreamSynchronize(k->cancel_stream);
        cudaStreamSynchronize(k->stream);
        if (k->d_done.p && k->copy_stream) {
          // row ranges whose tasks were never fetched: let the copy stream's waits pass (the results are discarded)
          cudaMemsetAsync(k->d_done.p, 0x7f, kDoneSlots * sizeof(unsigned int), k->cancel_stream);
          cudaStreamSynchronize(k->cancel_stream);
          cudaStreamSynchronize(k->copy_stream);
          apply_fixups(k);
        }
        k->computed = false;
      }
      c->computed = false;
      return fail(c, TSQ_ERR_CANCELLED, "cancelled");
    }
    struct timespec ts = {0, 200000};
    nanosleep(&ts, nullptr);
  }
  if (cb) cb(user, 0.9, "collecting results");
  rc = tsq_download(c);
  if (rc != TSQ_OK) return rc;
  if (cb) cb(user, 1.0, "done");
  return TSQ_OK;
}

int tsq_stream_results(tsq_ctx* c, int enable) {
  if (!c) return TSQ_ERR_INVALID;
  c->stream_out = enable != 0;
  for (tsq_ctx* k : c->kids) k->stream_out = c->stream_out;
  return TSQ_OK;
}

int tsq_set_result_buffers(tsq_ctx* c, int32_t* scores, double* distances, uint64_t count) {
  if (!c) return TSQ_ERR_INVALID;
  if (c->leader) return fail(c, TSQ_ERR_INVALID, "set the buffers on the multi-device context, not on its children");
  if ((scores == nullptr) != (count == 0)) return fail(c, TSQ_ERR_INVALID, "scores buffer and count must both be given or both be absent");
  if (scores && (c->prm.flags & TSQ_FLAG_SCORES_I16)) return fail(c, TSQ_ERR_INVALID, "TSQ_FLAG_SCORES_I16 delivers into the library's own buffer (tsq_scores16)");
  if (c->kids.empty()) cudaSetDevice(c->device);
  unregister_all(c);
  c->ext_scores = scores;
  c->ext_dist = distances;
  c->ext_count = count;
  c->downloaded = false;
  for (tsq_ctx* k : c->kids) k->downloaded = false;
  return TSQ_OK;
}

int tsq_results_sharded(tsq_ctx* c, int* sharded) {
  if (!c || !sharded) return TSQ_ERR_INVALID;
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_results_sharded before tsq_upload");
  *sharded = c->slab_mode ? 1 : 0;
  return TSQ_OK;
}

int tsq_scores(tsq_ctx* c, const int32_t** out, uint64_t* count) {
  if (!c || !out) return TSQ_ERR_INVALID;
  if (!c->downloaded) return fail(c, TSQ_ERR_STATE, "no results: call tsq_run or tsq_download first");
  if (c->prm.flags & TSQ_FLAG_SCORES_I16) return fail(c, TSQ_ERR_STATE, "TSQ_FLAG_SCORES_I16: the scores are int16, ask tsq_scores16");
  uint64_t first = 0, cnt = 0;
  host_extent(c, &first, &cnt);
  *out = c->ext_scores ? c->ext_scores : c->h_scores.p;
  if (count) *count = cnt;
  return TSQ_OK;
}

int tsq_scores16(tsq_ctx* c, const int16_t** out, uint64_t* count) {
  if (!c || !out) return TSQ_ERR_INVALID;
  if (!(c->prm.flags & TSQ_FLAG_SCORES_I16)) return fail(c, TSQ_ERR_STATE, "tsq_scores16 needs TSQ_FLAG_SCORES_I16");
  if (!c->downloaded) return fail(c, TSQ_ERR_STATE, "no results: call tsq_run or tsq_download first");
  uint64_t first = 0, cnt = 0;
  host_extent(c, &first, &cnt);
  *out = c->h_scores16.p;
  if (count) *count = cnt;
  return TSQ_OK;
}

int tsq_distances(tsq_ctx* c, const double** out, uint64_t* count) {
  if (!c || !out) return TSQ_ERR_INVALID;
  if (!c->downloaded) return fail(c, TSQ_ERR_STATE, "no results: call tsq_run or tsq_download first");
  if (c->prm.flags & TSQ_FLAG_NO_DISTANCES) return fail(c, TSQ_ERR_STATE, "distances disabled by TSQ_FLAG_NO_DISTANCES");
  uint64_t first = 0, cnt = 0;
  host_extent(c, &first, &cnt);
  *out = c->ext_scores ? c->ext_dist : c->h_dist.p;
  if (count) *count = cnt;
  return TSQ_OK;
}

int tsq_identities(tsq_ctx* c, const int32_t** out, uint64_t* count) {
  if (!c || !out) return TSQ_ERR_INVALID;
  if (!(c->prm.flags & TSQ_FLAG_IDENTITY)) return fail(c, TSQ_ERR_STATE, "identities need TSQ_FLAG_IDENTITY");
  if (!c->downloaded) return fail(c, TSQ_ERR_STATE, "no results: call tsq_run or tsq_download first");
  *out = c->h_nid.p;
  if (count) *count = c->n < 2 ? 0 : (uint64_t)c->n * (c->n - 1) / 2;
  return TSQ_OK;
}

int tsq_self_scores(tsq_ctx* c, const int32_t** self, uint32_t* n) {
  if (!c || !self) return TSQ_ERR_INVALID;
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "no sequences uploaded");
  *self = c->selfp->data();
  if (n) *n = c->n;
  return TSQ_OK;
}

int tsq_device_scores(tsq_ctx* c, void** d, uint64_t* count) {
  if (!c || !d) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) return fail(c, TSQ_ERR_STATE, "a multi-device context has no single device buffer");
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_device_scores before tsq_upload");
  if (!c->full_sorted) return fail(c, TSQ_ERR_STATE, "rank %d holds only its slab: use tsq_device_slab", c->prm.part_rank);
  *d = c->d_sorted.p;
  if (count) *count = c->n < 2 ? 0 : (uint64_t)c->n * (c->n - 1) / 2;
  return TSQ_OK;
}

int tsq_device_slab(tsq_ctx* c, void** d, uint64_t* first, uint64_t* count) {
  if (!c || !d) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) return fail(c, TSQ_ERR_STATE, "a multi-device context has no single device buffer");
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_device_slab before tsq_upload");
  *d = c->d_sorted.p;
  if (first) *first = c->full_sorted ? 0 : c->part_begin;
  if (count) *count = c->full_sorted ? (c->n < 2 ? 0 : (uint64_t)c->n * (c->n - 1) / 2) : c->part_end - c->part_begin;
  return TSQ_OK;
}

int tsq_partition(tsq_ctx* c, uint64_t* b, uint64_t* e) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_partition before tsq_upload");
  if (!c->kids.empty()) {   // the whole triangle
    if (b) *b = 0;
    if (e) *e = c->n < 2 ? 0 : (uint64_t)c->n * (c->n - 1) / 2;
    return TSQ_OK;
  }
  if (b) *b = c->part_begin;
  if (e) *e = c->part_end;
  return TSQ_OK;
}

int tsq_partition_of(tsq_ctx* c, int32_t rank, uint64_t* b, uint64_t* e) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_partition_of before tsq_upload");
  if (!c->kids.empty()) {   // device number `rank` of a multi-device context
    if (rank < 0 || rank >= (int32_t)c->kids.size()) return fail(c, TSQ_ERR_INVALID, "device %d outside 0..%zu", rank, c->kids.size());
    if (b) *b = c->kids[(size_t)rank]->part_begin;
    if (e) *e = c->kids[(size_t)rank]->part_end;
    return TSQ_OK;
  }
  if (rank < 0 || rank >= c->prm.part_world) return fail(c, TSQ_ERR_INVALID, "rank %d outside 0..%d", rank, c->prm.part_world);
  const uint64_t n = c->n, npairs = n < 2 ? 0 : n * (n - 1) / 2;
  auto start_of = [&](uint32_t row) -> uint64_t { return (n >= 2 && (uint64_t)row + 1 < n) ? tri(row, row + 1, n) : npairs; };
  uint64_t lo = start_of(c->first_row[(size_t)rank]), hi = start_of(c->first_row[(size_t)rank + 1]);
  if (lo > hi) lo = hi;
  if (b) *b = lo;
  if (e) *e = hi;
  return TSQ_OK;
}

int tsq_device_results(tsq_ctx* c, void** d_scores, void** d_dist, uint64_t* count) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) return fail(c, TSQ_ERR_STATE, "a multi-device context has no single device buffer");
  if (!c->finalized) return fail(c, TSQ_ERR_STATE, "tsq_device_results before tsq_finalize");
  if (c->slab_mode) return fail(c, TSQ_ERR_STATE, "results of this partition stay sharded (tsq_results_sharded)");
  if (d_scores) *d_scores = (c->perm_identity && c->idshift == 0) ? (void*)c->d_sorted.p : (void*)c->d_scores.p;
  if (d_dist) *d_dist = (c->prm.flags & TSQ_FLAG_NO_DISTANCES) ? nullptr : (void*)c->d_dist.p;
  if (count) *count = c->n < 2 ? 0 : (uint64_t)c->n * (c->n - 1) / 2;
  return TSQ_OK;
}

int tsq_guide_tree(tsq_ctx* c, const tsq_merge** merges, uint32_t* count) {
  if (!c) return TSQ_ERR_INVALID;
  if (c->prm.flags & TSQ_FLAG_NO_DISTANCES) return fail(c, TSQ_ERR_STATE, "guide tree needs distances (TSQ_FLAG_NO_DISTANCES set)");
  if (!c->kids.empty()) {   // on the first device, from the distance slabs collected there
    int rc = multi_prepare_first_device(c);
    if (rc == TSQ_OK) rc = tsq_guide_tree(c->kids[0], merges, count);
    if (rc != TSQ_OK && !c->kids[0]->err.empty()) copy_error(c, c->kids[0]);
    c->tree_ms = c->kids[0]->tree_ms;
    return rc;
  }
  if (!c->finalized) return fail(c, TSQ_ERR_STATE, "tsq_guide_tree before tsq_run / tsq_finalize");
  if (c->slab_mode && !c->d_dist_full.p)
    return fail(c, TSQ_ERR_STATE, "this rank holds a slab of the matrix only (tsq_results_sharded): build the tree from the assembled host matrix");
  const uint32_t n = c->n;
  if (!c->have_tree && n > TSQ_GUIDE_TREE_MAX_N)
    return fail(c, TSQ_ERR_RANGE, "guide tree of %u sequences: the UPGMA kernel keeps a dense n x n fp64 matrix and runs its n-1 "
                "dependent merges on one SM; the limit is %u sequences (%.1f GB)", n, (unsigned)TSQ_GUIDE_TREE_MAX_N,
                (double)TSQ_GUIDE_TREE_MAX_N * TSQ_GUIDE_TREE_MAX_N * 8 / 1e9);
  if (!c->have_tree) {
    c->merges.assign(n >= 2 ? n - 1 : 0, tsq_merge{0, 0, 0.0});
    if (n >= 2) {
      TSQ_CUDA(c, cudaSetDevice(c->device));
      TSQ_CUDA(c, c->d_treeD.reserve((size_t)n * n));
      TSQ_CUDA(c, c->d_treemin.reserve(n));
      TSQ_CUDA(c, c->d_treeh.reserve(n));
      TSQ_CUDA(c, c->d_treeu.reserve((size_t)5 * n + 8));
      TSQ_CUDA(c, c->d_merges.reserve(n - 1));
      tsq::UpgmaParams u{};
      u.dist = c->d_dist_full.p ? c->d_dist_full.p : c->d_dist.p;   // child 0 of a multi-device context: the collected slabs
      u.D = c->d_treeD.p;
      u.rowmin = c->d_treemin.p;
      u.nheight = c->d_treeh.p;
      u.rowarg = c->d_treeu.p;
      u.active = c->d_treeu.p + n;
      u.csize = c->d_treeu.p + 2 * (size_t)n;
      u.node = c->d_treeu.p + 3 * (size_t)n;
      u.rescan = c->d_treeu.p + 4 * (size_t)n;
      u.merges = c->d_merges.p;
      u.n = n;
      TSQ_CUDA(c, cudaEventRecord(c->tev0, c->stream));
      TSQ_CUDA(c, tsq::upgma_launch(u, c->stream));
      TSQ_CUDA(c, cudaEventRecord(c->tev1, c->stream));
      TSQ_CUDA(c, cudaMemcpyAsync(c->merges.data(), c->d_merges.p, (size_t)(n - 1) * sizeof(tsq_merge), cudaMemcpyDeviceToHost, c->stream));
      TSQ_CUDA(c, cudaStreamSynchronize(c->stream));
      float ms = 0;
      if (cudaEventElapsedTime(&ms, c->tev0, c->tev1) != cudaSuccess) cudaGetLastError();
      c->tree_ms = ms;
      c->st.launches += 2;
    }
    c->have_tree = true;
  }
  if (merges) *merges = c->merges.data();
  if (count) *count = (uint32_t)c->merges.size();
  return TSQ_OK;
}

int tsq_write_newick(tsq_ctx* c, const char* const* labels, const char* path) {
  if (!c || !path) return TSQ_ERR_INVALID;
  const tsq_merge* mg = nullptr;
  uint32_t cnt = 0;
  int rc = tsq_guide_tree(c, &mg, &cnt);
  if (rc != TSQ_OK) return rc;
  const uint32_t n = c->n;
  FILE* f = fopen(path, "w");
  if (!f) return fail(c, TSQ_ERR_IO, "cannot write %s", path);
  auto leaf_name = [&](uint32_t i) -> std::string {
    if (!(labels && labels[i])) return "s" + std::to_string(i);
    // the label as the matrix file and the FASTA headers spell it; one that contains Newick structure
    // characters goes in single quotes (a quote inside doubled), as the format prescribes, so that
    // `clustalo --guidetree-in` and `--distmat-in` see the same names
    const std::string name = labels[i];
    if (name.empty()) return "s" + std::to_string(i);
    if (name.find_first_of("():;,[]'\" \t") == std::string::npos) return name;
    std::string q = "'";
    for (char ch : name) {
      q.push_back(ch);
      if (ch == '\'') q.push_back('\'');
    }
    q.push_back('\'');
    return q;
  };
  if (n == 0) {
    fputs(";\n", f);
  } else if (n == 1) {
    fprintf(f, "%s;\n", leaf_name(0).c_str());
  } else {
    // iterative post-order writer (trees of 10^5 leaves can be a caterpillar: no recursion)
    auto height_of = [&](uint32_t id) -> double { return id < n ? 0.0 : mg[id - n].height; };
    struct Frame { uint32_t id; int state; double parent_h; };
    std::vector<Frame> st;
    st.push_back({n + cnt - 1, 0, -1.0});
    while (!st.empty()) {
      Frame& fr = st.back();
      if (fr.id < n) {
        fprintf(f, "%s:%.6f", leaf_name(fr.id).c_str(), fr.parent_h);   // leaves sit at height 0
        st.pop_back();
        continue;
      }
      const tsq_merge& m = mg[fr.id - n];
      if (fr.state == 0) {
        fputc('(', f);
        fr.state = 1;
        st.push_back({m.left, 0, m.height});
      } else if (fr.state == 1) {
        fputc(',', f);
        fr.state = 2;
        st.push_back({m.right, 0, m.height});
      } else {
        if (fr.parent_h < 0) fputs(")", f);
        else fprintf(f, "):%.6f", fr.parent_h - height_of(fr.id));
        st.pop_back();
      }
    }
    fputs(";\n", f);
  }
  fclose(f);
  return TSQ_OK;
}

namespace {

// The device side of msa_host.h: persistent memory is bump-allocated from cudaMalloc'd chunks (a job
// makes ~2n small allocations), scratch is one growing block, everything runs on the context's stream.
class CudaMsaDevice : public tsq::MsaDevice {
 public:
  explicit CudaMsaDevice(cudaStream_t s) : s_(s), dev_(current_device()) {}
  ~CudaMsaDevice() override {
    if (!chunks_.empty() && cudaDeviceSynchronize() != cudaSuccess) cudaGetLastError();
    for (void* p : chunks_) BlockCache::get().give(dev_, p);
    scr_.release();
    stage_.release();
  }
  cudaError_t err = cudaSuccess;   // first CUDA error seen
  double t_alloc = 0, t_scratch = 0, t_copy = 0, t_wait = 0, t_launch = 0;   // host clock per kind of call (TSQ_MSA_DEBUG)

  void* alloc(size_t bytes) override {
    Timer tm(t_alloc);
    bytes = tsq::msa_align(std::max<size_t>(bytes, 1));
    if (bytes > left_) {
      const size_t chunk = std::max(bytes, (size_t)64 << 20);
      void* p = nullptr;
      if (!ok(BlockCache::get().take(dev_, chunk, &p))) return nullptr;
      chunks_.push_back(p);
      cur_ = (char*)p;
      left_ = chunk;
    }
    void* r = cur_;
    cur_ += bytes;
    left_ -= bytes;
    return r;
  }
  void* scratch(size_t bytes) override {
    Timer tm(t_scratch);
    if (bytes > scr_.cap) {
      if (!ok(cudaStreamSynchronize(s_))) return nullptr;   // nothing may still be using the old block
      if (!ok(scr_.reserve(bytes + bytes / 4))) return nullptr;
    }
    return scr_.p;
  }
  // Small copies (a level's task list up, its merged lengths back: once per tree level, a thousand times for a deep
  // tree) go through a page-locked staging block: a pageable cudaMemcpyAsync stages and waits inside the driver.
  // Safe to reuse per call: every level ends in d2h's synchronize before the next h2d writes the block.
  bool h2d(void* d, const void* h, size_t b) override {
    Timer tm(t_copy);
    if (b <= kStage && stage_up()) {
      if (!ok(cudaStreamSynchronize(s_))) return false;   // (an earlier staged upload of this stream has left the block)
      memcpy(stage_.p, h, b);
      return ok(cudaMemcpyAsync(d, stage_.p, b, cudaMemcpyHostToDevice, s_));
    }
    return ok(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, s_));
  }
  bool d2h(void* h, const void* d, size_t b) override {
    Timer tm(t_wait);
    if (b <= kStage && stage_up()) {
      if (!ok(cudaMemcpyAsync(stage_.p + kStage, d, b, cudaMemcpyDeviceToHost, s_)) || !ok(cudaStreamSynchronize(s_))) return false;
      memcpy(h, stage_.p + kStage, b);
      return true;
    }
    return ok(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, s_)) && ok(cudaStreamSynchronize(s_));
  }
  bool fill(void* d, int v, size_t b) override { return ok(cudaMemsetAsync(d, v, b, s_)); }
  bool launch_leaves(const tsq::MsaLeaf* l, uint32_t n, uint32_t nsym) override { return ok(tsq::msa_leaf_launch(l, n, nsym, s_)); }
  bool launch_merges(const tsq::MsaTask* t, uint32_t count, uint32_t threads, uint32_t smem_bytes, const tsq::MsaConst& k) override {
    Timer tm(t_launch);
    return ok(tsq::msa_merge_launch(t, count, threads, smem_bytes, k, s_));
  }
  bool launch_rows(const tsq::MsaRows& p) override { return ok(tsq::msa_rows_launch(p, s_)); }

 private:
  struct Timer {
    double& acc;
    double t0;
    explicit Timer(double& a) : acc(a), t0(now_ms()) {}
    ~Timer() { acc += now_ms() - t0; }
  };
  bool ok(cudaError_t e) {
    if (e != cudaSuccess && err == cudaSuccess) err = e;
    return e == cudaSuccess;
  }
  static constexpr size_t kStage = 256u << 10;   // bytes per direction
  bool stage_up() { return stage_.p != nullptr || stage_.reserve(2 * kStage) == cudaSuccess; }
  PinnedBuf<uint8_t> stage_;
  cudaStream_t s_;
  int dev_;
  std::vector<void*> chunks_;
  char* cur_ = nullptr;
  size_t left_ = 0;
  DevBuf<uint8_t> scr_;
};

const char* letters_of(const tsq_ctx* c) { return c->prm.alphabet == TSQ_NUCLEOTIDE ? "ACGTN" : "ARNDCQEGHILKMFPSTWYVBZX"; }

}  // namespace

int tsq_msa(tsq_ctx* c, const char** rows, uint32_t* nrows, uint32_t* ncols, const uint32_t** tree_order) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_msa before tsq_run");
  if (!c->kids.empty()) {   // the progressive alignment is a chain of dependent merges: one device (the first) runs it
    tsq_ctx* k0 = c->kids[0];
    int rc = multi_prepare_first_device(c);
    if (rc != TSQ_OK) return rc;
    k0->msa_cancel = c->msa_cancel;
    rc = tsq_msa(k0, rows, nrows, ncols, tree_order);
    k0->msa_cancel = nullptr;
    if (rc != TSQ_OK) copy_error(c, k0);
    c->msa_ms = k0->msa_ms;
    c->tree_ms = k0->tree_ms;
    return rc;
  }
  const tsq_merge* mg = nullptr;
  uint32_t cnt = 0;
  int rc = tsq_guide_tree(c, &mg, &cnt);
  if (rc != TSQ_OK) return rc;
  if (!c->have_msa) {
    const uint32_t n = c->n;
    TSQ_CUDA(c, cudaSetDevice(c->device));
    tsq::MsaJob job;
    job.n = n;
    job.d_sym = c->d_lin.p;
    job.sym_off.assign(n, 0);
    job.len.assign(n, 0);
    for (uint32_t k = 0; k < n; k++) {   // perm: sorted position -> submitted index (the tree's leaf id)
      job.sym_off[c->perm[k]] = c->loff[k];
      job.len[c->perm[k]] = c->lens[k];
    }
    job.left.resize(cnt);
    job.right.resize(cnt);
    for (uint32_t t = 0; t < cnt; t++) {
      job.left[t] = mg[t].left;
      job.right[t] = mg[t].right;
    }
    job.nsym = (uint32_t)c->nsym;
    job.smat.assign(c->matrix.begin(), c->matrix.end());
    job.go = c->go;
    job.ge = c->ge;
    job.letters = letters_of(c);
    job.cancel = c->msa_cancel;
    if (const char* e = getenv("TSQ_MSA_CELLS_PER_THREAD")) job.cells_per_thread = (uint32_t)std::max(1, atoi(e));   // tuning runs
    job.device_sms = (uint32_t)std::max(1, c->sm_count);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) job.scratch_budget = std::max<size_t>(free_b / 4, (size_t)256 << 20);
    else cudaGetLastError();
    tsq::MsaOut out;
    const double t0 = now_ms();
    int mrc;
    cudaError_t derr;
    {
      CudaMsaDevice dev(c->stream);
      mrc = tsq::msa_progressive(dev, job, out);
      derr = dev.err;
      if (mrc == tsq::MSA_OK && derr == cudaSuccess) derr = cudaStreamSynchronize(c->stream);
      if (getenv("TSQ_MSA_DEBUG"))
        fprintf(stderr, "tsq_msa: %.1f ms so far: alloc %.1f, scratch %.1f, copies %.1f, waiting on the device %.1f, launches %.1f; %u levels, %u launches\n",
                now_ms() - t0, dev.t_alloc, dev.t_scratch, dev.t_copy, dev.t_wait, dev.t_launch, out.levels, out.launches);
    }
    if (derr != cudaSuccess) {
      cudaGetLastError();
      return fail(c, derr == cudaErrorMemoryAllocation ? TSQ_ERR_NOMEM : TSQ_ERR_CUDA, "tsq_msa: %s", cudaGetErrorString(derr));
    }
    if (mrc == tsq::MSA_CANCELLED) return fail(c, TSQ_ERR_CANCELLED, "cancelled");
    if (mrc == tsq::MSA_NOMEM) return fail(c, TSQ_ERR_NOMEM, "tsq_msa: out of device memory");
    if (mrc != tsq::MSA_OK) return fail(c, TSQ_ERR_CUDA, "tsq_msa: internal error %d", mrc);
    c->msa_ms = now_ms() - t0;
    c->msa_rows.swap(out.rows);
    c->msa_order.swap(out.tree_order);
    c->msa_cols = out.ncols;
    c->st.launches += out.launches;
    c->have_msa = true;
  }
  if (rows) *rows = reinterpret_cast<const char*>(c->msa_rows.data());
  if (nrows) *nrows = c->n;
  if (ncols) *ncols = c->msa_cols;
  if (tree_order) *tree_order = c->msa_order.data();
  return TSQ_OK;
}

int tsq_write_msa_fasta(tsq_ctx* c, const char* const* headers, const char* const* residues, const uint32_t* lengths,
                        const char* path, int tree_order) {
  if (!c || !path) return TSQ_ERR_INVALID;
  const char* rows = nullptr;
  const uint32_t* order = nullptr;
  uint32_t n = 0, cols = 0;
  int rc = tsq_msa(c, &rows, &n, &cols, &order);
  if (rc != TSQ_OK) return rc;
  if (residues && !lengths) return fail(c, TSQ_ERR_INVALID, "residues without lengths");
  FILE* f = fopen(path, "w");
  if (!f) return fail(c, TSQ_ERR_IO, "cannot write %s", path);
  // the whole file is assembled in memory and written once (a thousand rows are seven thousand lines:
  // formatted writes per line cost more than the alignment kernels of a small job)
  std::string row, text;
  text.reserve((size_t)n * ((size_t)cols + cols / 60 + 64));
  for (uint32_t q = 0; q < n; q++) {
    const uint32_t r = tree_order ? order[q] : q;
    row.assign(rows + (size_t)r * cols, cols);
    if (residues && residues[r]) {
      // the caller's own spelling of every residue (case, J/O/U, ...): k-th kept input byte -> k-th non-gap column
      const unsigned char* in = reinterpret_cast<const unsigned char*>(residues[r]);
      uint32_t k = 0;
      for (uint32_t col = 0; col < cols; col++) {
        if (row[col] == '-') continue;
        while (k < lengths[r] && is_gap_or_space(in[k])) k++;
        if (k < lengths[r]) row[col] = (char)in[k++];
      }
    }
    if (headers && headers[r]) {
      const char* h = headers[r];
      if (h[0] != '>') text.push_back('>');
      text.append(h);
      text.push_back('\n');
    } else {
      text.append(">s").append(std::to_string(r)).push_back('\n');
    }
    for (uint32_t at = 0; at < cols; at += 60) {
      text.append(row, at, std::min<uint32_t>(60, cols - at));
      text.push_back('\n');
    }
  }
  if (!text.empty() && fwrite(text.data(), 1, text.size(), f) != text.size()) {
    fclose(f);
    return fail(c, TSQ_ERR_IO, "write to %s failed", path);
  }
  if (fclose(f) != 0) return fail(c, TSQ_ERR_IO, "write to %s failed", path);
  return TSQ_OK;
}

int tsq_align_pair(tsq_ctx* c, uint32_t i, uint32_t j, char* row_i, char* row_j, uint32_t capacity, uint32_t* columns,
                   int32_t* score) {
  if (!c || !row_i || !row_j) return TSQ_ERR_INVALID;
  if (!c->uploaded) return fail(c, TSQ_ERR_STATE, "tsq_align_pair before tsq_upload");
  if (!c->kids.empty()) {   // every device holds the whole database: the first one serves single pairs
    const int rc = tsq_align_pair(c->kids[0], i, j, row_i, row_j, capacity, columns, score);
    if (rc != TSQ_OK) copy_error(c, c->kids[0]);
    return rc;
  }
  if (i >= c->n || j >= c->n) return fail(c, TSQ_ERR_INVALID, "pair (%u, %u) outside 0..%u", i, j, c->n);
  uint32_t si = c->n, sj = c->n;   // sorted positions of the two submitted indices
  for (uint32_t k = 0; k < c->n; k++) {
    if (c->perm[k] == i) si = k;
    if (c->perm[k] == j) sj = k;
  }
  const uint32_t m = c->lens[si], n = c->lens[sj];
  if ((uint64_t)m + n + 1 > capacity) return fail(c, TSQ_ERR_INVALID, "capacity %u < %llu", capacity, (unsigned long long)m + n + 1);
  const size_t dir_bytes = ((size_t)m + n + 1) * ((size_t)std::min(m, n) + 1);   // diagonal-major, padded rows
  if (dir_bytes > ((size_t)16 << 30)) return fail(c, TSQ_ERR_NOMEM, "direction matrix of %u x %u cells exceeds 16 GiB", m, n);
  TSQ_CUDA(c, cudaSetDevice(c->device));
  DevBuf<uint8_t> d_dir, d_out;
  DevBuf<int32_t> d_work;   // 7 rolling diagonals, the plain score table, info[2]
  const uint32_t nsym = (uint32_t)c->nsym;
  const size_t w_diag = 7 * ((size_t)m + 1), w_smat = (size_t)nsym * nsym;
  std::vector<int32_t> smat(w_smat);
  for (size_t k = 0; k < w_smat; k++) smat[k] = c->matrix[k];
  std::vector<uint8_t> out(2 * ((size_t)m + n) + 1);
  int32_t info[2] = {0, 0};
  cudaError_t e = d_dir.reserve(dir_bytes);
  if (e == cudaSuccess) e = d_out.reserve(2 * ((size_t)m + n) + 1);
  if (e == cudaSuccess) e = d_work.reserve(w_diag + w_smat + 2);
  if (e == cudaSuccess) {
    cudaStream_t s = c->stream;
    tsq::TbParams t{};
    t.a = c->d_lin.p + c->loff[si];
    t.b = c->d_lin.p + c->loff[sj];
    t.m = m;
    t.n = n;
    t.diag = d_work.p;
    t.smat = d_work.p + w_diag;
    t.info = d_work.p + w_diag + w_smat;
    t.nsym = nsym;
    t.go = c->go;
    t.ge = c->ge;
    t.dir = d_dir.p;
    t.out_a = d_out.p;
    t.out_b = d_out.p + m + n;
    e = cudaMemcpyAsync(d_work.p + w_diag, smat.data(), w_smat * 4, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = tsq::traceback_launch(t, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out.data(), d_out.p, 2 * ((size_t)m + n), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(info, t.info, sizeof info, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  }
  d_dir.release(); d_out.release(); d_work.release();
  if (e != cudaSuccess) return fail(c, e == cudaErrorMemoryAllocation ? TSQ_ERR_NOMEM : TSQ_ERR_CUDA, "tsq_align_pair: %s", cudaGetErrorString(e));
  const uint32_t cols = (uint32_t)info[0];
  if (cols > m + n) return fail(c, TSQ_ERR_CUDA, "internal: traceback of %u columns for %u + %u residues", cols, m, n);
  const char* letters = c->prm.alphabet == TSQ_NUCLEOTIDE ? "ACGTN" : "ARNDCQEGHILKMFPSTWYVBZX";
  for (uint32_t k = 0; k < cols; k++) {   // the kernel wrote the path from the end: reverse
    const uint8_t a = out[cols - 1 - k], b = out[(size_t)m + n + cols - 1 - k];
    row_i[k] = a == 0xff ? '-' : letters[a];
    row_j[k] = b == 0xff ? '-' : letters[b];
  }
  row_i[cols] = row_j[cols] = 0;
  if (columns) *columns = cols;
  if (score) *score = info[1];
  c->st.launches += 1;
  return TSQ_OK;
}

int tsq_consensus(tsq_ctx* c, const char* const* rows, uint32_t nrows, uint32_t ncols, double plurality, char* out) {
  if (!c) return TSQ_ERR_INVALID;
  if ((nrows > 0 && !rows) || (ncols > 0 && !out)) return fail(c, TSQ_ERR_INVALID, "null alignment / output");
  if (!c->kids.empty()) {
    const int rc = tsq_consensus(c->kids[0], rows, nrows, ncols, plurality, out);
    if (rc != TSQ_OK) copy_error(c, c->kids[0]);
    return rc;
  }
  if (ncols == 0) return TSQ_OK;
  if (nrows == 0) {
    memset(out, '?', ncols);
    return TSQ_OK;
  }
  if (plurality < 0) plurality = (double)nrows / 2.0;   // Consensus.cpp:164-175
  TSQ_CUDA(c, cudaSetDevice(c->device));
  const size_t cells = (size_t)nrows * ncols;
  PinnedBuf<uint8_t> h;
  DevBuf<uint8_t> d_aln, d_out, d_tab;
  int rc = TSQ_OK;
  cudaError_t e = h.reserve(cells + ncols);
  if (e == cudaSuccess) e = d_aln.reserve(cells);
  if (e == cudaSuccess) e = d_out.reserve(ncols);
  if (e == cudaSuccess) e = d_tab.reserve(23 * 23 + 32);
  if (e == cudaSuccess) {
    for (uint32_t r = 0; r < nrows; r++) {
      if (!rows[r]) { rc = fail(c, TSQ_ERR_INVALID, "alignment row %u is null", r); break; }
      memcpy(h.p + (size_t)r * ncols, rows[r], ncols);
    }
  }
  if (e == cudaSuccess && rc == TSQ_OK) {
    cudaStream_t s = c->stream;
    e = cudaMemcpyAsync(d_aln.p, h.p, cells, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab.p, kBlosum62, 23 * 23, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_tab.p + 23 * 23, kProteinIndex, 26, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess)
      e = tsq::consensus_launch(d_aln.p, nrows, ncols, plurality, reinterpret_cast<const int8_t*>(d_tab.p), d_tab.p + 23 * 23,
                                d_out.p, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h.p + cells, d_out.p, ncols, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) memcpy(out, h.p + cells, ncols);
  }
  h.release(); d_aln.release(); d_out.release(); d_tab.release();
  if (rc != TSQ_OK) return rc;
  if (e != cudaSuccess) return fail(c, e == cudaErrorMemoryAllocation ? TSQ_ERR_NOMEM : TSQ_ERR_CUDA, "tsq_consensus: %s", cudaGetErrorString(e));
  return TSQ_OK;
}

int tsq_plan_partition(const tsq_params* params, const uint32_t* lengths, uint32_t n, int32_t world,
                       uint64_t* begins, uint64_t* ends) {
  if (world < 1 || (n > 0 && !lengths) || !begins || !ends) return TSQ_ERR_INVALID;
  tsq_ctx tmp;
  tsq_params p;
  tsq_default_params(&p);
  if (params) {
    if (params->struct_size < 8 || params->struct_size > sizeof(tsq_params)) return TSQ_ERR_INVALID;
    memcpy(&p, params, params->struct_size);
  }
  const int nsym = p.alphabet == TSQ_NUCLEOTIDE ? 5 : 23;
  const int8_t* m = p.matrix ? p.matrix : (p.alphabet == TSQ_NUCLEOTIDE ? kDna : kBlosum62);
  tmp.go = p.gap_open < 0 ? (p.alphabet == TSQ_NUCLEOTIDE ? 10 : 11) : p.gap_open;
  tmp.ge = p.gap_extend < 0 ? 1 : p.gap_extend;
  tmp.smin = *std::min_element(m, m + nsym * nsym);
  tmp.smax = *std::max_element(m, m + nsym * nsym);
  tmp.delta = tmp.smin < 0 ? (-tmp.smin + 1) / 2 : 0;
  const uint32_t max16 = (p.flags & TSQ_FLAG_FORCE_S32) ? 0 : max_len16_of(&tmp);
  std::vector<uint32_t> lens(lengths, lengths + n);
  std::stable_sort(lens.begin(), lens.end());
  uint32_t lo = 0;
  while (lo < n && lens[lo] == 0) lo++;
  uint32_t hi = lo;
  const uint32_t inter_max = inter_task_limit(max16, p.flags, nullptr);
  while (hi < n && lens[hi] <= inter_max) hi++;
  std::vector<uint32_t> first_row;
  const bool w16 = wave16_ok(&tmp, nsym, p.flags) && !(p.flags & TSQ_FLAG_IDENTITY);
  plan_rows(lens, lo, hi, world, first_row, w16 ? 1.5 : 2.4);
  const uint64_t npairs = n < 2 ? 0 : (uint64_t)n * (n - 1) / 2;
  auto start_of = [&](uint32_t row) -> uint64_t { return (n >= 2 && row + 1 < n) ? tri(row, row + 1, n) : npairs; };
  for (int r = 0; r < world; r++) {
    begins[r] = start_of(first_row[(size_t)r]);
    ends[r] = start_of(first_row[(size_t)r + 1]);
    if (begins[r] > ends[r]) begins[r] = ends[r];
  }
  return TSQ_OK;
}

int tsq_get_stats(tsq_ctx* c, tsq_stats* out) {
  if (!c || !out) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) {
    // the whole job: sums over the devices; kernel_ms = the slowest device (they run side by side)
    tsq_stats t = c->st;
    t.n_pairs = t.cells = t.cells_s16 = t.cells_s32 = 0;
    t.launches = t.upload_launches = 0;
    t.h2d_bytes = 0;
    t.kernel_ms = 0;
    for (tsq_ctx* k : c->kids) {
      tsq_stats ks;
      tsq_get_stats(k, &ks);
      t.n_pairs += ks.n_pairs;
      t.cells += ks.cells;
      t.cells_s16 += ks.cells_s16;
      t.cells_s32 += ks.cells_s32;
      t.launches += ks.launches;
      t.upload_launches += ks.upload_launches;
      t.h2d_bytes += ks.h2d_bytes;
      t.kernel_ms = std::max(t.kernel_ms, ks.kernel_ms);
      t.strip_width = ks.strip_width;
    }
    t.n_sequences = c->n;
    t.gcups_kernel = t.kernel_ms > 0 ? (double)t.cells / (t.kernel_ms * 1e6) : 0.0;
    t.sm_count = (uint32_t)c->sm_count * (uint32_t)c->kids.size();
    t.tree_ms = c->tree_ms;
    t.msa_ms = c->msa_ms;
    *out = t;
    return TSQ_OK;
  }
  c->st.n_sequences = c->n;
  c->st.n_pairs = c->pairs_part;
  c->st.cells_s16 = c->cells16;
  c->st.cells_s32 = c->cells32;
  c->st.cells = c->cells16 + c->cells32;
  c->st.gcups_kernel = c->st.kernel_ms > 0 ? (double)c->st.cells / (c->st.kernel_ms * 1e6) : 0.0;
  c->st.sm_count = (uint32_t)c->sm_count;
  c->st.strip_width = (uint32_t)c->K;
  c->st.tree_ms = c->tree_ms;
  c->st.msa_ms = c->msa_ms;
  *out = c->st;
  return TSQ_OK;
}

int tsq_get_device_stats(tsq_ctx* c, int32_t index, tsq_stats* out) {
  if (!c || !out) return TSQ_ERR_INVALID;
  if (c->kids.empty()) return index == 0 ? tsq_get_stats(c, out) : fail(c, TSQ_ERR_INVALID, "device index %d of a one-device context", index);
  if (index < 0 || index >= (int32_t)c->kids.size()) return fail(c, TSQ_ERR_INVALID, "device index %d outside 0..%zu", index, c->kids.size());
  return tsq_get_stats(c->kids[(size_t)index], out);
}

int tsq_get_limits(tsq_ctx* c, tsq_limits* out) {
  if (!c || !out) return TSQ_ERR_INVALID;
  memset(out, 0, sizeof *out);
  bool g32 = false;
  out->max_len_packed = c->max_len16;
  out->max_len_inter = inter_task_limit(c->max_len16, c->prm.flags, &g32);
  out->inter_is_32bit = g32 ? 1 : 0;
  const long long lip = std::max(std::abs(c->smax), std::abs(c->smin)) + c->go + c->ge + 2 * c->delta;
  out->wave_window = tsq::w16_window((uint32_t)c->nsym, lip);
  out->wave_window_max = kWave16WindowMax;
  out->wave_packed = (wave16_ok(c, c->nsym, c->prm.flags) && !(c->prm.flags & TSQ_FLAG_IDENTITY)) ? 1 : 0;
  out->delta = c->delta;
  out->bias_at_limit = (int32_t)bias_for(c, c->max_len16 + 64);
  return TSQ_OK;
}

int tsq_measure_dpx_rate(tsq_ctx* c, double* ops, double* mhz) {
  if (!c) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) return tsq_measure_dpx_rate(c->kids[0], ops, mhz);
  TSQ_CUDA(c, cudaSetDevice(c->device));
  TSQ_CUDA(c, tsq::dpx_probe(c->sm_count, ops, mhz, c->stream));
  return TSQ_OK;
}

int tsq_measure_pipe_rates(tsq_ctx* c, tsq_pipe_rates* out) {
  if (!c || !out) return TSQ_ERR_INVALID;
  if (!c->kids.empty()) return tsq_measure_pipe_rates(c->kids[0], out);
  memset(out, 0, sizeof *out);
  TSQ_CUDA(c, cudaSetDevice(c->device));
  TSQ_CUDA(c, tsq::dpx_probe(c->sm_count, &out->dpx_per_clk_sm, &out->sm_mhz, c->stream));
  TSQ_CUDA(c, tsq::mix_probe(c->sm_count, &out->issue_per_clk_sm, &out->mix_packed_cells_per_clk_sm, c->stream));
  return TSQ_OK;
}

// Square PHYLIP-style matrix ("n", then "label d d d ..." rows, %.6f) from the packed upper triangle.
// Host only.  std::to_chars(fixed, 6) prints the same correctly rounded digits as printf("%.6f") at a
// fraction of the cost: at n = 1 000 the million fprintf calls of the first version took longer than the
// whole GPU job.
int tsq_write_distmat(const char* path, const char* const* labels, uint32_t n, const double* packed) {
  if (!path || (n > 0 && !labels) || (n > 1 && !packed)) return TSQ_ERR_INVALID;
  FILE* fo = fopen(path, "w");
  if (!fo) return TSQ_ERR_IO;
  std::string line;
  line.reserve((size_t)n * 10 + 64);
  char num[64];
  fprintf(fo, "%u\n", n);
  bool ok = true;
  for (uint64_t i = 0; i < n && ok; i++) {
    line.assign(labels[i] ? labels[i] : "");
    for (uint64_t j = 0; j < n; j++) {
      double v = 0.0;
      if (i != j) v = packed[i < j ? tri(i, j, n) : tri(j, i, n)];
      line.push_back(' ');
      const auto r = std::to_chars(num, num + sizeof num, v, std::chars_format::fixed, 6);
      if (r.ec == std::errc()) line.append(num, r.ptr);
      else line.append("nan");
    }
    line.push_back('\n');
    ok = fwrite(line.data(), 1, line.size(), fo) == line.size();
  }
  if (fclose(fo) != 0) ok = false;
  return ok ? TSQ_OK : TSQ_ERR_IO;
}

// ---- file-level convenience -------------------------------------------------------------------
// FASTA reading follows tweakseq/Core/FASTAFile.cpp:71-147 (state machine: '>' or ';' starts a
// record, further ';' lines directly after a header are skipped, blank lines ignored, lines
// trimmed) and the label rule of parseComment (:177-187: header[1 .. first space)).
int tsq_run_fasta(const char* fin, const char* fout, const tsq_params* params, tsq_log_cb log, void* user,
                  volatile int* cancel) {
  if (!fin || !fout) return TSQ_ERR_INVALID;
  auto say = [&](const std::string& s) { if (log) log(user, s.c_str()); };
  const double t_start = now_ms();
  double t_read = 0, t_create = 0, t_dist = 0, t_files = 0, t_align = 0;   // stage split of the call (last log line)
  // the whole file in one read, then lines by memchr (the state machine is the reference reader's, FASTAFile.cpp:96-140)
  std::string text;
  {
    FILE* f = fopen(fin, "rb");
    if (!f) {
      say(std::string("cannot open ") + fin);
      return TSQ_ERR_IO;
    }
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof chunk, f)) > 0) text.append(chunk, got);
    const bool bad = ferror(f) != 0;
    fclose(f);
    if (bad) {
      say(std::string("cannot read ") + fin);
      return TSQ_ERR_IO;
    }
  }
  std::vector<std::string> labels, headers, seqs;
  int state = 0;  // 0 seeking header, 1 just read header, 2 reading residues
  auto add_header = [&](const char* a, size_t len) {
    // label: what follows the marker up to the first blank from position 1 on (find(' ', 1) of the old reader)
    const char* sp = len > 1 ? static_cast<const char*>(memchr(a + 1, ' ', len - 1)) : nullptr;
    labels.emplace_back(a + 1, sp ? (size_t)(sp - a - 1) : len - 1);
    headers.emplace_back(a, len);
    seqs.emplace_back();
    state = 1;
  };
  for (size_t pos = 0; pos < text.size();) {
    const char* a = text.data() + pos;
    const char* nl = static_cast<const char*>(memchr(a, '\n', text.size() - pos));
    const char* b = nl ? nl : text.data() + text.size();
    pos = (size_t)(b - text.data()) + 1;
    while (a < b && isspace((unsigned char)*a)) a++;           // trim
    while (b > a && isspace((unsigned char)b[-1])) b--;
    if (a == b) continue;
    const size_t len = (size_t)(b - a);
    const char f = *a;
    const bool hdr = (f == '>' || f == ';');
    if (state == 0) {
      if (hdr) add_header(a, len);
    } else if (state == 1) {
      if (f == ';') continue;
      seqs.back().append(a, len);   // whatever follows a header is residues, a second '>' line included (FASTAFile.cpp:117-124)
      state = 2;
    } else {
      if (hdr) add_header(a, len);
      else seqs.back().append(a, len);
    }
  }
  char msg[256];
  snprintf(msg, sizeof msg, "tsq-b200: read %zu sequences from %s", seqs.size(), fin);
  say(msg);
  tsq_params prm;
  tsq_default_params(&prm);
  if (params) {
    if (params->struct_size < 8 || params->struct_size > sizeof(tsq_params)) return TSQ_ERR_INVALID;
    memcpy(&prm, params, params->struct_size);
    prm.struct_size = (uint32_t)sizeof(tsq_params);
  }
  if (prm.alphabet == TSQ_ALPHABET_AUTO) {
    // what clustalo does without --seqtype (a tweakseq project holds either kind: SequenceFile::DNA / ::Proteins)
    std::vector<const char*> rp(seqs.size());
    std::vector<uint32_t> rl(seqs.size());
    for (size_t i = 0; i < seqs.size(); i++) {
      rp[i] = seqs[i].data();
      rl[i] = (uint32_t)seqs[i].size();
    }
    prm.alphabet = tsq_detect_alphabet(rp.data(), rl.data(), (uint32_t)seqs.size());
    say(prm.alphabet == TSQ_NUCLEOTIDE ? "tsq-b200: residues look like nucleotides (ACGTN +5/-4, gap 10/1)"
                                       : "tsq-b200: residues look like protein (BLOSUM62, gap 11/1)");
  }
  params = &prm;
  t_read = now_ms() - t_start;
  tsq_ctx* c = nullptr;
  int rc = tsq_create(&c, params);
  t_create = now_ms() - t_start - t_read;
  if (rc != TSQ_OK) {
    say(std::string("tsq_create failed: ") + tsq_status_string(rc));
    return rc;
  }
  std::vector<const char*> ptrs(seqs.size());
  std::vector<uint32_t> lens(seqs.size());
  for (size_t i = 0; i < seqs.size(); i++) {
    ptrs[i] = seqs[i].data();
    lens[i] = (uint32_t)seqs[i].size();
  }
  const double t_d0 = now_ms();
  rc = tsq_set_sequences(c, ptrs.data(), lens.data(), (uint32_t)seqs.size());
  if (rc == TSQ_OK) rc = tsq_run(c, nullptr, nullptr, cancel);
  if (rc != TSQ_OK) {
    say(std::string("tsq-b200: ") + tsq_last_error(c));
    tsq_destroy(c);
    return rc;
  }
  t_dist = now_ms() - t_d0;
  const double t_f0 = now_ms();
  const double* d = nullptr;
  uint64_t cnt = 0;
  rc = tsq_distances(c, &d, &cnt);
  // TSQ_FLAG_MSA_OUT: fout is the alignment itself (what Project::readNewAlignment ingests,
  // Project.cpp:908-1032); the matrix moves to <fout>.distmat.  Otherwise fout is the matrix.
  const bool msa_out = (params->flags & TSQ_FLAG_MSA_OUT) != 0;
  const bool keep_matrix = !msa_out || (params->flags & TSQ_FLAG_KEEP_DISTMAT);   // n^2 numbers of text: only on request
  const std::string matrix_path = msa_out ? std::string(fout) + ".distmat" : std::string(fout);
  std::vector<const char*> lab(labels.size());
  for (size_t i = 0; i < labels.size(); i++) lab[i] = labels[i].c_str();
  if (rc == TSQ_OK && keep_matrix && tsq_write_distmat(matrix_path.c_str(), lab.data(), (uint32_t)seqs.size(), d) != TSQ_OK) {
    say(std::string("cannot write ") + matrix_path);
    rc = TSQ_ERR_IO;
  }
  if (rc == TSQ_OK) {
    // guide tree for clustalo --guidetree-in, next to the matrix; with the alignment itself as the output
    // the tree file is written on request only (nothing should appear next to tweakseq's temporary file)
    const std::string tree = std::string(fout) + ".dnd";
    if ((!msa_out || (params->flags & TSQ_FLAG_KEEP_TREE)) && tsq_write_newick(c, lab.data(), tree.c_str()) == TSQ_OK)
      say("tsq-b200: wrote guide tree " + tree);
    tsq_stats st;
    tsq_get_stats(c, &st);
    snprintf(msg, sizeof msg, "tsq-b200: %llu pairs, %.3e cells on %d device%s, kernel %.3f ms (%.1f GCUPS)%s%s", (unsigned long long)cnt,
             (double)st.cells, std::max(1, params->n_devices), params->n_devices > 1 ? "s" : "", st.kernel_ms, st.gcups_kernel,
             keep_matrix ? ", wrote " : "", keep_matrix ? matrix_path.c_str() : "");
    say(msg);
  }
  t_files = now_ms() - t_f0;
  const double t_a0 = now_ms();
  if (rc == TSQ_OK && msa_out) {
    if (cancel && *cancel) {
      rc = TSQ_ERR_CANCELLED;
    } else {
      // rows in tree order like the reference's clustalo argv (--output-order=tree-order, ClustalO.cpp:51)
      // unless TSQ_FLAG_INPUT_ORDER asks for the input's order, header lines and residue spelling exactly as read, so readNewAlignment matches every label
      std::vector<const char*> hdr(headers.size()), res(seqs.size());
      for (size_t i = 0; i < headers.size(); i++) hdr[i] = headers[i].c_str();
      for (size_t i = 0; i < seqs.size(); i++) res[i] = seqs[i].data();
      c->msa_cancel = cancel;
      rc = tsq_write_msa_fasta(c, hdr.data(), res.data(), lens.data(), fout, (params->flags & TSQ_FLAG_INPUT_ORDER) ? 0 : 1);
      c->msa_cancel = nullptr;
      if (rc == TSQ_OK) {
        tsq_stats st;
        tsq_get_stats(c, &st);
        uint32_t cols = 0;
        tsq_msa(c, nullptr, nullptr, &cols, nullptr);
        snprintf(msg, sizeof msg, "tsq-b200: progressive alignment, %zu rows x %u columns in %.1f ms, wrote %s", seqs.size(), cols,
                 st.msa_ms, fout);
        say(msg);
      } else {
        say(std::string("tsq-b200: ") + tsq_last_error(c));
      }
    }
  }
  t_align = now_ms() - t_a0;
  if (rc == TSQ_OK) {
    // where the time of the call went, as the editor's user waits for it (bench.py's e2e_plugin parses this line)
    tsq_stats st;
    tsq_get_stats(c, &st);
    const double t_destroy0 = now_ms();
    tsq_destroy(c);
    c = nullptr;
    const double t_end = now_ms();
    snprintf(msg, sizeof msg,
             "tsq-b200: timing ms: read=%.3f context=%.3f distances=%.3f pack_h2d=%.3f kernels=%.3f finalize_d2h=%.3f "
             "files=%.3f tree_kernels=%.3f alignment_write=%.3f alignment=%.3f release=%.3f total=%.3f",
             t_read, t_create, t_dist, st.upload_ms, st.kernel_ms, st.download_ms, t_files, st.tree_ms, t_align, st.msa_ms,
             t_end - t_destroy0, t_end - t_start);
    say(msg);
  }
  if (c) tsq_destroy(c);
  return rc;
}

}  // extern "C"
