// msa_emul.cpp -- TEST INFRASTRUCTURE: runs the progressive-alignment plan of libtsqb200.so
// (tweakseq_b200/csrc/msa_host.h) and the phase functions of its kernels (msa.cuh) on the CPU, one
// emulated thread at a time, so that the product's own planning and per-thread code are compared with
// the oracle (tsq_oracle_msa) in the CPU test tier.  A phase has no barrier inside, so running its
// threads one after the other -- here in DESCENDING thread order, to expose any accidental dependence
// on thread order -- is a legal schedule.  "Device" memory is malloc'd and filled with garbage so that
// a read of something no phase wrote shows up as a wrong answer.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <vector>

#include "../tweakseq_b200/csrc/msa_host.h"

namespace {

class EmulDevice : public tsq::MsaDevice {
 public:
  std::vector<void*> blocks;
  void* scr = nullptr;
  size_t scr_cap = 0;
  int order = 0;   // 0: threads descending, 1: ascending
  uint32_t force_threads = 0;
  bool no_smem = false;   // force the global-scratch path of the rolling diagonals
  ~EmulDevice() override {
    for (void* p : blocks) free(p);
    free(scr);
  }
  void* alloc(size_t bytes) override {
    void* p = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
    if (p) { memset(p, 0xA7, bytes); blocks.push_back(p); }
    return p;
  }
  void* scratch(size_t bytes) override {
    if (bytes > scr_cap) {
      free(scr);
      scr = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
      scr_cap = scr ? bytes : 0;
    }
    if (scr) memset(scr, 0x5C, bytes);
    return scr;
  }
  bool h2d(void* d, const void* s, size_t b) override { memcpy(d, s, b); return true; }
  bool d2h(void* d, const void* s, size_t b) override { memcpy(d, s, b); return true; }
  bool fill(void* d, int v, size_t b) override { memset(d, v, b); return true; }
  template <class F>
  void each_thread(int nt, F f) {
    if (order) for (int t = 0; t < nt; t++) f(t);
    else for (int t = nt - 1; t >= 0; t--) f(t);
  }
  // the body of msa_merge_cta<T>, barrier by barrier
  template <typename T>
  void sweep(const tsq::MsaTask& t, const tsq::MsaConst& k, void* shared, uint32_t smem_bytes, int nt) {
    // what lives in "shared memory": as msa_merge_cta decides, with the same sizes
    const size_t db = tsq::msa_round16(tsq::msa_diag_bytes(t.Lx, t.Ly, sizeof(T) == 4));
    void* const edge = db <= (size_t)smem_bytes ? shared : (void*)t.diag;
    uint16_t* codes = nullptr;   // 4-bit direction codes behind the edge arrays and the column-score tables
    if (db <= (size_t)smem_bytes) {
      const uint32_t Lb = tsq::msa_big_is_x(t) ? t.Lx : t.Ly, Ls = tsq::msa_big_is_x(t) ? t.Ly : t.Lx;
      const uint32_t nsmall = tsq::msa_big_is_x(t) ? t.ny : t.nx;
      const size_t pb = tsq::msa_round16((size_t)k.nsym * Lb * 4),
                   lb = tsq::msa_round16((size_t)tsq::msa_list_rows(k.nsym, nsmall) * Ls * 4), nb = tsq::msa_round16((size_t)Ls * 4);
      if (db + pb + lb + nb + tsq::msa_round16(tsq::msa_code_bytes(t.Lx, t.Ly)) <= (size_t)smem_bytes)
        codes = reinterpret_cast<uint16_t*>(static_cast<char*>(shared) + db + pb + lb + nb);
    }
    each_thread(nt, [&](int tid) { tsq::msa_prep_phase(t, k, tid, nt); });
    const tsq::MsaSweep<T> sw = tsq::msa_sweep_init<T>(t, k, edge, nullptr, codes);
    // barrier by barrier as msa_sweep_cta: tile threads and column-score workers share an interval
    const int tw = tsq::msa_tile_threads(t.Lx, t.Ly, nt);
    each_thread(nt, [&](int tid) { tsq::msa_edge_phase<T>(sw, tid, nt); tsq::msa_sub_phase<T>(sw, 0, tid, nt); });
    const int steps = tsq::msa_sweep_steps<T>(sw);
    for (int st = 0; st < steps; ++st) {
      each_thread(nt, [&](int tid) {
        if (tid < tw) tsq::msa_tile_phase<T>(sw, st, tid, tw);
        else tsq::msa_sub_phase<T>(sw, st + 1, tid - tw, nt - tw);
      });
      if (tw == nt) each_thread(nt, [&](int tid) { tsq::msa_sub_phase<T>(sw, st + 1, tid, nt); });
    }
    // phase 3 twice: the serial statement, then the device's one-warp walk with its ballots done by a loop over 32
    // lanes (lanes in the emulation's thread order); the two paths must agree entry for entry
    const long long score = tsq::msa_final_score<T>(sw);
    tsq::msa_walk_phase(t, score, codes);
    const uint32_t len1 = t.res->len;
    std::vector<int32_t> path1(t.path, t.path + 2 * (size_t)len1);
    memset(t.path, 0x5C, 2 * (size_t)len1 * sizeof(int32_t));
    t.res->len = 0xdeadbeefu;
    walk_warp(t, score, codes);
    if (t.res->len != len1 || memcmp(path1.data(), t.path, path1.size() * sizeof(int32_t)) != 0 || t.res->score != score) walk_mismatch = true;
  }
  bool walk_mismatch = false;
  // msa_walk_warp (msa.cuh), the collectives emulated: every per-lane piece runs for all 32 lanes before the next one
  void walk_warp(const tsq::MsaTask& t, long long score, const uint16_t* codes) {
    tsq::MsaWalk w{(int)t.Lx, (int)t.Ly, 0, 0u};
    while (w.i > 0 && w.j > 0) {
      uint32_t code[32][2];
      bool valid[32][2];
      each_thread(32, [&](int lane) { tsq::msa_walk_fetch(t, codes, w, lane, code[lane], valid[lane]); });
      bool more = true;
      for (int h = 0; h < 2 && more; ++h) {
        uint32_t stops = 0, turns = 0;
        each_thread(32, [&](int lane) {
          const bool turn = valid[lane][h] && tsq::msa_walk_turn(w.state, code[lane][h]);
          if (!valid[lane][h] || turn) stops |= 1u << lane;
          if (turn) turns |= 1u << lane;
        });
        int s;
        bool s_turn;
        const int count = tsq::msa_walk_count(w.state, stops, turns, &s, &s_turn);
        each_thread(32, [&](int lane) { tsq::msa_walk_store(t, w, lane, count); });
        more = tsq::msa_walk_advance(w, s, s_turn, count, code[s & 31][h]);
      }
    }
    each_thread(32, [&](int lane) { tsq::msa_walk_tails(t, w, lane); });
    tsq::msa_walk_finish(t, w, score);
  }
  bool launch_leaves(const tsq::MsaLeaf* l, uint32_t n, uint32_t nsym) override {
    for (uint32_t r = 0; r < n; r++) each_thread(128, [&](int t) { tsq::msa_leaf_phase(l[r], nsym, t, 128); });
    return true;
  }
  bool launch_merges(const tsq::MsaTask* tasks, uint32_t count, uint32_t threads, uint32_t smem_bytes,
                     const tsq::MsaConst& k) override {
    const int nt = (int)(force_threads ? force_threads : threads);
    if (nt < 1 || nt > 1024 || smem_bytes > 227 * 1024) return false;
    if (no_smem) smem_bytes = 0;
    std::vector<long long> shared(smem_bytes / sizeof(long long) + 1);
    for (uint32_t b = 0; b < count; b++) {
      const tsq::MsaTask t = tasks[b];
      std::fill(shared.begin(), shared.end(), (long long)0x5C5C5C5C5C5C5C5CLL);   // a new CTA: shared memory is garbage
      if (t.narrow) sweep<int32_t>(t, k, shared.data(), smem_bytes, nt);
      else sweep<long long>(t, k, shared.data(), smem_bytes, nt);
      each_thread(nt, [&](int tid) { tsq::msa_build_phase(t, k, tid, nt); });
    }
    return true;
  }
  bool launch_rows(const tsq::MsaRows& p) override {
    for (uint32_t r = 0; r < p.n; r++) each_thread(256, [&](int t) { tsq::msa_rows_phase(p, r, t, 256); });
    return true;
  }
};

}  // namespace

extern "C" int msa_emul(const uint8_t* seqs, const uint64_t* offs, const uint32_t* lens, uint32_t n, const int8_t* mat,
                        int nsym, int go, int ge, const uint32_t* left, const uint32_t* right, const char* letters,
                        uint8_t* rows_out, uint64_t rows_cap, uint32_t* ncols, long long* merge_scores,
                        uint32_t* tree_order, uint64_t scratch_budget, uint32_t force_threads, int ascending,
                        uint32_t* launches, uint32_t* levels) {
  EmulDevice dev;
  dev.order = ascending;
  dev.force_threads = force_threads & 0xffffu;
  dev.no_smem = (force_threads >> 16) & 1u;
  const bool force_wide = (force_threads >> 17) & 1u;
  tsq::MsaJob job;
  job.n = n;
  job.d_sym = seqs;
  job.sym_off.assign(offs, offs + n);
  job.len.assign(lens, lens + n);
  if (n > 1) {
    job.left.assign(left, left + n - 1);
    job.right.assign(right, right + n - 1);
  }
  job.smat.resize((size_t)nsym * nsym);
  for (int i = 0; i < nsym * nsym; i++) job.smat[i] = mat[i];
  job.nsym = (uint32_t)nsym;
  job.go = go;
  job.ge = ge;
  job.letters = letters;
  if (scratch_budget) job.scratch_budget = scratch_budget;
  job.force_wide = force_wide;
  const int raised = (force_threads >> 18) & 1u;   // the cancel flag, already up
  job.cancel = &raised;
  tsq::MsaOut out;
  const int rc = tsq::msa_progressive(dev, job, out);
  if (rc != tsq::MSA_OK) return rc;
  if (dev.walk_mismatch) return -2;   // the one-warp walk-back and the serial one disagree
  if (out.rows.size() > rows_cap) return -1;
  memcpy(rows_out, out.rows.data(), out.rows.size());
  *ncols = out.ncols;
  for (size_t t = 0; t < out.merge_score.size(); t++) merge_scores[t] = out.merge_score[t];
  for (size_t i = 0; i < out.tree_order.size(); i++) tree_order[i] = out.tree_order[i];
  if (launches) *launches = out.launches;
  if (levels) *levels = out.levels;
  return 0;
}
