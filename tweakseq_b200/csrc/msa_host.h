// msa_host.h -- host-side plan of the progressive alignment (msa.cuh): tree levels, batches, memory
// layout, descriptors.  Pure C++ over a tiny device interface: libtsqb200.so implements it with CUDA
// (tsq_api.cpp: arena of cudaMalloc chunks, stream copies, the msa_* kernels); tests/msa_emul.cpp
// implements it with malloc and runs the kernels' phase functions thread by thread on the CPU, so this
// exact planning code is checked against the oracle without a GPU.  No alignment arithmetic lives here.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>

#include "msa.cuh"

namespace tsq {

class MsaDevice {
 public:
  virtual ~MsaDevice() {}
  // Memory that lives until the job ends (profiles, column maps, results); 256-byte aligned; nullptr = out of memory.
  virtual void* alloc(size_t bytes) = 0;
  // One scratch block of at least `bytes`, reused by every batch (contents do not survive the next call).
  virtual void* scratch(size_t bytes) = 0;
  virtual bool h2d(void* dst, const void* src, size_t bytes) = 0;
  virtual bool d2h(void* dst, const void* src, size_t bytes) = 0;   // complete on return, after all earlier work
  virtual bool fill(void* dst, int byte, size_t bytes) = 0;
  virtual bool launch_leaves(const MsaLeaf* d_leaves, uint32_t n, uint32_t nsym) = 0;
  // smem_bytes: shared memory per CTA for the sweep's edge arrays (merges that need more use their global scratch)
  virtual bool launch_merges(const MsaTask* d_tasks, uint32_t count, uint32_t threads, uint32_t smem_bytes, const MsaConst& k) = 0;
  virtual bool launch_rows(const MsaRows& p) = 0;
};

struct MsaJob {
  uint32_t n = 0;                       // sequences (leaves 0..n-1, submitted order)
  const uint8_t* d_sym = nullptr;       // device: encoded residues of all sequences
  std::vector<uint64_t> sym_off;        // per leaf: offset into d_sym
  std::vector<uint32_t> len;            // per leaf
  std::vector<uint32_t> left, right;    // n-1 merges; node of merge t = n + t
  std::vector<int32_t> smat;            // nsym x nsym
  uint32_t nsym = 0;
  int32_t go = 0, ge = 0;
  const char* letters = "";            // nsym characters
  size_t scratch_budget = (size_t)4 << 30;   // scratch bytes one launch may use (at least one merge always runs)
  bool force_wide = false;              // tests: int64 sweep even where int32 would do
  uint32_t cells_per_thread = 1;        // CTA size = longest diagonal of tiles / this (tuning knob; any value is correct)
  uint32_t device_sms = 148;            // SMs of the device (CTA-size policy only)
  const volatile int* cancel = nullptr; // "Stop" (SeqEditMainWin.cpp:803-812): polled before every launch
};

struct MsaOut {
  uint32_t ncols = 0;
  std::vector<uint8_t> rows;            // n x ncols characters ('-' = gap), submitted order
  std::vector<long long> merge_score;   // n-1
  std::vector<uint32_t> merge_cols;     // n-1: columns after each merge
  std::vector<uint32_t> tree_order;     // leaves left to right (clustalo --output-order=tree-order)
  uint32_t launches = 0, levels = 0;
};

enum { MSA_OK = 0, MSA_NOMEM = 1, MSA_DEVICE = 2, MSA_BAD_TREE = 3, MSA_CANCELLED = 4 };

inline size_t msa_align(size_t x) { return (x + 255) & ~(size_t)255; }
constexpr size_t kMsaSmemLimit = 200 * 1024;   // of the 227 KB a CTA may have on sm_100

inline int msa_progressive(MsaDevice& dev, const MsaJob& job, MsaOut& out) {
  const uint32_t n = job.n, nsym = job.nsym;
  out = MsaOut();
  if (n == 0) return MSA_OK;
  if (n >= (1u << 24) || nsym > (uint32_t)kMsaMaxSym) return MSA_BAD_TREE;   // column counts share a word with their letter (msa_prep_phase)
  const uint32_t nnodes = 2 * n - 1, NONE = 0xffffffffu;
  if (job.left.size() != n - 1 || job.right.size() != n - 1) return MSA_BAD_TREE;

  // ---- tree shape: parents, sizes, levels (a merge may run once both children exist) ----
  std::vector<uint32_t> parent(nnodes, NONE), size(nnodes, 1), level(nnodes, 0);
  for (uint32_t t = 0; t + 1 < n; t++) {
    const uint32_t l = job.left[t], r = job.right[t], z = n + t;
    if (l >= z || r >= z || l == r || parent[l] != NONE || parent[r] != NONE) return MSA_BAD_TREE;
    parent[l] = parent[r] = z;
    size[z] = size[l] + size[r];
    level[z] = 1 + std::max(level[l], level[r]);
  }
  uint32_t nlevels = 0;
  for (uint32_t t = 0; t + 1 < n; t++) nlevels = std::max(nlevels, level[n + t]);
  std::vector<std::vector<uint32_t>> by_level(nlevels + 1);
  for (uint32_t t = 0; t + 1 < n; t++) by_level[level[n + t]].push_back(t);
  out.levels = nlevels;

  // ---- constants and leaf profiles ----
  MsaConst kc{};
  {
    int32_t* d_smat = (int32_t*)dev.alloc((size_t)nsym * nsym * 4);
    if (!d_smat) return MSA_NOMEM;
    if (!dev.h2d(d_smat, job.smat.data(), (size_t)nsym * nsym * 4)) return MSA_DEVICE;
    kc.smat = d_smat; kc.nsym = nsym; kc.go = job.go; kc.ge = job.ge;
  }
  std::vector<uint32_t> ncol(nnodes, 0), cap(nnodes, 0);
  std::vector<uint32_t*> prof(nnodes, nullptr);
  std::vector<const uint32_t*> nodemap(nnodes, nullptr);
  std::vector<MsaLeaf> leaves(n);
  {
    // all leaf profiles in one allocation, each leaf's columns padded to a multiple of 64
    size_t words = 0;
    std::vector<size_t> at(n);
    for (uint32_t r = 0; r < n; r++) {
      cap[r] = (job.len[r] + 63u) & ~63u;
      at[r] = words;
      words += (size_t)cap[r] * nsym;
    }
    uint32_t* base = (uint32_t*)dev.alloc(std::max<size_t>(words, 1) * 4);
    if (!base) return MSA_NOMEM;
    for (uint32_t r = 0; r < n; r++) {
      ncol[r] = job.len[r];
      prof[r] = base + at[r];
      leaves[r].sym = job.d_sym + job.sym_off[r];
      leaves[r].len = job.len[r];
      leaves[r].cap = cap[r];
      leaves[r].c = prof[r];
    }
  }
  MsaLeaf* d_leaves = (MsaLeaf*)dev.alloc((size_t)n * sizeof(MsaLeaf));
  if (!d_leaves) return MSA_NOMEM;
  if (!dev.h2d(d_leaves, leaves.data(), (size_t)n * sizeof(MsaLeaf))) return MSA_DEVICE;
  if (!dev.launch_leaves(d_leaves, n, nsym)) return MSA_DEVICE;
  out.launches++;

  // ---- merges, level by level; a level is cut into batches that fit the scratch budget ----
  out.merge_score.assign(n - 1, 0);
  out.merge_cols.assign(n - 1, 0);
  MsaResult* d_res = nullptr;
  if (n > 1) {
    d_res = (MsaResult*)dev.alloc((size_t)(n - 1) * sizeof(MsaResult));
    if (!d_res) return MSA_NOMEM;
  }
  uint32_t slot = 0;   // results are stored in launch order, so a batch reads back one contiguous range
  std::vector<MsaTask> tasks;
  std::vector<MsaResult> res;
  struct Scratch { size_t diag, pbig, lst, lnz, dir, path, total; };
  auto scratch_of = [&](uint32_t Lx, uint32_t Ly) -> Scratch {
    const size_t mn = std::min(Lx, Ly), mx = std::max<size_t>(std::max(Lx, Ly), 1);
    Scratch q;
    q.diag = msa_align(msa_diag_bytes(Lx, Ly, false));
    q.pbig = msa_align((size_t)nsym * mx * 4);
    q.lst = msa_align((size_t)nsym * mx * 4);
    q.lnz = msa_align(mx * 4);
    q.dir = msa_align(((size_t)Lx + Ly + 1) * (mn + 1));
    q.path = msa_align(std::max<size_t>(2 * ((size_t)Lx + Ly), 1) * 4);
    q.total = q.diag + q.pbig + q.lst + q.lnz + q.dir + q.path;
    return q;
  };
  int32_t max_abs_s = 0;
  for (int32_t v : job.smat) max_abs_s = std::max(max_abs_s, v < 0 ? -v : v);
  auto narrow_of = [&](uint32_t t) -> bool {
    const uint32_t x = job.left[t], y = job.right[t];
    return !job.force_wide && msa_fits_narrow(size[x], size[y], ncol[x], ncol[y], max_abs_s, job.go, job.ge);
  };
  const char* trace_env = getenv("TSQ_MSA_DEBUG");
  const bool trace = trace_env && atoi(trace_env) >= 2;
  for (uint32_t lv = 1; lv <= nlevels; lv++) {
    const std::vector<uint32_t>& ms = by_level[lv];
    size_t b = 0;
    while (b < ms.size()) {
      if (job.cancel && *job.cancel) return MSA_CANCELLED;
      // batch [b, e): as many merges of this level as the scratch budget takes
      size_t e = b, bytes = msa_align(sizeof(MsaTask));
      uint32_t longest = 0;
      size_t smem = 0;
      while (e < ms.size()) {
        const uint32_t t = ms[e], Lx = ncol[job.left[t]], Ly = ncol[job.right[t]];
        const size_t need = scratch_of(Lx, Ly).total + msa_align(sizeof(MsaTask));
        if (e > b && bytes + need > job.scratch_budget) break;
        bytes += need;
        longest = std::max(longest, std::min(Lx, Ly) + 1);
        // shared memory of the launch: the sweep's edge arrays of every merge that fits, and behind them the
        // column-score tables where those fit too (msa.cuh: msa_merge_cta decides per merge with the same sizes)
        const size_t db = msa_round16(msa_diag_bytes(Lx, Ly, narrow_of(t)));
        const size_t tb = msa_table_bytes(Lx, Ly, nsym, std::min(size[job.left[t]], size[job.right[t]]));
        const size_t cb = msa_round16(msa_code_bytes(Lx, Ly));   // 4-bit direction codes, where they fit as well
        if (db + tb + cb <= kMsaSmemLimit) smem = std::max(smem, db + tb + cb);
        else if (db + tb <= kMsaSmemLimit) smem = std::max(smem, db + tb);
        else if (db <= kMsaSmemLimit) smem = std::max(smem, db);
        e++;
      }
      const size_t count = e - b;
      const size_t head = msa_align(count * sizeof(MsaTask));
      bytes = head;
      for (size_t q = b; q < e; q++) bytes += scratch_of(ncol[job.left[ms[q]]], ncol[job.right[ms[q]]]).total;
      char* sc = (char*)dev.scratch(bytes);
      if (!sc) return MSA_NOMEM;
      tasks.assign(count, MsaTask{});
      size_t at = head;
      for (size_t q = b; q < e; q++) {
        const uint32_t t = ms[q], x = job.left[t], y = job.right[t], z = n + t;
        const uint32_t Lx = ncol[x], Ly = ncol[y];
        MsaTask& k = tasks[q - b];
        cap[z] = std::max<uint32_t>((Lx + Ly + 63u) & ~63u, 64u);
        // merged profile and the two column maps: one persistent allocation
        const size_t pw = (size_t)cap[z] * nsym, total = pw + Lx + Ly;
        uint32_t* p = (uint32_t*)dev.alloc(total * 4);
        if (!p) return MSA_NOMEM;
        prof[z] = p;
        k.cx = prof[x]; k.cy = prof[y]; k.cn = p;
        k.capx = cap[x]; k.capy = cap[y]; k.capn = cap[z];
        k.Lx = Lx; k.Ly = Ly; k.nx = size[x]; k.ny = size[y];
        k.narrow = narrow_of(t) ? 1u : 0u;
        k.mapx = p + pw; k.mapy = p + pw + Lx;
        nodemap[x] = k.mapx; nodemap[y] = k.mapy;
        const Scratch sz = scratch_of(Lx, Ly);
        k.diag = (long long*)(sc + at);  at += sz.diag;
        k.pbig = (int32_t*)(sc + at);    at += sz.pbig;
        k.lst = (uint32_t*)(sc + at);    at += sz.lst;
        k.lnz = (uint32_t*)(sc + at);    at += sz.lnz;
        k.dir = (uint8_t*)(sc + at);     at += sz.dir;
        k.path = (int32_t*)(sc + at);    at += sz.path;
        k.res = d_res + slot + (q - b);
      }
      if (!dev.h2d(sc, tasks.data(), count * sizeof(MsaTask))) return MSA_DEVICE;
      // one thread per tile of the longest anti-diagonal of tiles (cells_per_thread = 1), or fewer threads that take
      // several tiles each; never under 128, for the column-parallel phases around the sweep
      const uint32_t per_thread = std::max<uint32_t>(job.cells_per_thread, 1u) * (uint32_t)kMsaTile;
      const uint32_t want = (longest + per_thread - 1) / per_thread;
      // the column-parallel phases around the sweep (letter scores: nsym^2 multiply-adds per column) want many threads;
      // a level with more merges than the device holds CTAs of 512 threads wants many CTAs per SM instead
      const uint32_t floor_threads = count <= job.device_sms ? 512u : count <= 2 * job.device_sms ? 256u : 128u;
      const uint32_t threads = std::min<uint32_t>(512u, std::max<uint32_t>(floor_threads, (want + 31u) & ~31u));
      const auto t_l0 = std::chrono::steady_clock::now();
      if (!dev.launch_merges((const MsaTask*)sc, (uint32_t)count, threads, (uint32_t)smem, kc)) return MSA_DEVICE;
      out.launches++;
      res.resize(count);
      if (!dev.d2h(res.data(), d_res + slot, count * sizeof(MsaResult))) return MSA_DEVICE;
      if (trace)   // TSQ_MSA_DEBUG=2: one line per launch (launch + wait + read-back of the merged lengths)
        fprintf(stderr, "tsq_msa: level %u: %zu merges, longest diagonal %u, %u threads, %zu B shared, %.3f ms\n", lv, count, longest,
                threads, smem, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_l0).count());
      for (size_t q = b; q < e; q++) {
        const uint32_t t = ms[q], z = n + t;
        const MsaResult& r = res[q - b];
        if (r.len > ncol[job.left[t]] + ncol[job.right[t]]) return MSA_DEVICE;   // cannot happen: a path has <= Lx + Ly columns
        ncol[z] = r.len;
        out.merge_score[t] = r.score;
        out.merge_cols[t] = r.len;
      }
      slot += (uint32_t)count;
      b = e;
    }
  }

  // ---- final rows: every residue follows the column maps up to the root ----
  const uint32_t root = n == 1 ? 0 : nnodes - 1;
  out.ncols = ncol[root];
  out.rows.assign((size_t)n * out.ncols, (uint8_t)'-');
  if (out.ncols > 0) {
    const size_t bytes = (size_t)n * out.ncols;
    uint8_t* d_out = (uint8_t*)dev.alloc(bytes);
    uint32_t* d_parent = (uint32_t*)dev.alloc((size_t)nnodes * 4);
    const uint32_t** d_nodemap = (const uint32_t**)dev.alloc((size_t)nnodes * sizeof(uint32_t*));
    if (!d_out || !d_parent || !d_nodemap) return MSA_NOMEM;
    if (!dev.fill(d_out, '-', bytes)) return MSA_DEVICE;
    if (!dev.h2d(d_parent, parent.data(), (size_t)nnodes * 4)) return MSA_DEVICE;
    if (!dev.h2d(d_nodemap, nodemap.data(), (size_t)nnodes * sizeof(uint32_t*))) return MSA_DEVICE;
    MsaRows rp{};
    rp.leaves = d_leaves; rp.parent = d_parent; rp.nodemap = d_nodemap; rp.out = d_out;
    rp.ncols = out.ncols; rp.n = n;
    memset(rp.letters, 0, sizeof rp.letters);
    strncpy(rp.letters, job.letters, sizeof rp.letters - 1);
    if (!dev.launch_rows(rp)) return MSA_DEVICE;
    out.launches++;
    if (!dev.d2h(out.rows.data(), d_out, bytes)) return MSA_DEVICE;
  }

  // ---- leaves left to right (iterative: a caterpillar tree of 10^5 leaves must not recurse) ----
  out.tree_order.reserve(n);
  std::vector<uint32_t> st{root};
  while (!st.empty()) {
    const uint32_t id = st.back();
    st.pop_back();
    if (id < n) { out.tree_order.push_back(id); continue; }
    st.push_back(job.right[id - n]);
    st.push_back(job.left[id - n]);
  }
  return MSA_OK;
}

}  // namespace tsq
