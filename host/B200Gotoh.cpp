// B200Gotoh.cpp -- see B200Gotoh.h.  Mirrors tweakseq/Core/ClustalO.cpp:48-111.
#include "B200Gotoh.h"

#include <cstring>

#include "../include/tsq_b200.h"

namespace tsqhost {

B200Gotoh::B200Gotoh() { init(); }
B200Gotoh::~B200Gotoh() {}

void B200Gotoh::init() {  // ClustalO.cpp:92-98
  name_ = "b200gotoh";
  version_ = "";
  executable_ = "libtsqb200.so";
}

void B200Gotoh::makeCommand(std::string& fin, std::string& fout, std::string& exec,
                            std::vector<std::string>& arglist) {
  // Informational for an in-process tool: what a clustalo run consuming our matrix looks like
  // (ClustalO.cpp:48-52 plus --distmat-in).
  exec = executable_;
  arglist = {"--force", "-v", "--outfmt=fa", "--output-order=tree-order", "-i", fin, "--distmat-in", fout};
}

void B200Gotoh::writeSettings(SettingsDocument& doc) {  // ClustalO.cpp:54-61
  SettingsElement e;
  e.children.push_back({"name", name()});
  e.children.push_back({"path", executable()});
  e.children.push_back({"preferred", preferred() ? "yes" : "no"});
  e.children.push_back({"gap_open", std::to_string(gapOpen)});
  e.children.push_back({"gap_extend", std::to_string(gapExtend)});
  e.children.push_back({"device", std::to_string(device)});
  e.children.push_back({"devices", std::to_string(devices)});
  e.children.push_back({"alphabet", detectAlphabet ? "auto" : nucleotide ? "nucleotide" : "protein"});
  e.children.push_back({"align_in_process", alignInProcess ? "yes" : "no"});
  doc.alignment_tools.push_back(e);
}

void B200Gotoh::readSettings(SettingsDocument& doc) {  // ClustalO.cpp:63-86
  for (auto& tool : doc.alignment_tools) {
    for (auto& kv : tool.children) {
      if (kv.first == "name" && kv.second != name_) break;
      if (kv.first == "path") executable_ = kv.second;
      if (kv.first == "preferred") setPreferred(kv.second == "yes");
      if (kv.first == "gap_open") gapOpen = std::stoi(kv.second);
      if (kv.first == "gap_extend") gapExtend = std::stoi(kv.second);
      if (kv.first == "device") device = std::stoi(kv.second);
      if (kv.first == "devices") devices = std::stoi(kv.second);
      if (kv.first == "alphabet") {
        detectAlphabet = kv.second == "auto";
        nucleotide = kv.second == "nucleotide";
      }
      if (kv.first == "align_in_process") alignInProcess = kv.second == "yes";
    }
  }
  getVersion();
}

void B200Gotoh::getVersion() {  // ClustalO.cpp:100-111 asks `clustalo --version`
  version_ = tsq_version_string();
}

static void fill(tsq_params& p, const B200Gotoh& t) {
  tsq_default_params(&p);
  p.alphabet = t.nucleotide ? TSQ_NUCLEOTIDE : TSQ_PROTEIN;
  p.gap_open = t.gapOpen;
  p.gap_extend = t.gapExtend;
  p.device = t.device;
  p.n_devices = t.devices;
  if (t.identityDistance) p.flags |= TSQ_FLAG_IDENTITY;
  if (t.alignInProcess) p.flags |= TSQ_FLAG_MSA_OUT;
}

int B200Gotoh::run(const std::string& fin, const std::string& fout, const LogSink& log, CancelFlag* cancel) {
  tsq_params p;
  fill(p, *this);
  if (detectAlphabet) p.alphabet = TSQ_ALPHABET_AUTO;
  struct Ctx { const LogSink* log; } ctx{&log};
  auto cb = [](void* user, const char* line) {
    const LogSink* l = static_cast<Ctx*>(user)->log;
    if (*l) (*l)(line);
  };
  return tsq_run_fasta(fin.c_str(), fout.c_str(), &p, cb, &ctx, cancel);
}

int B200Gotoh::distanceMatrix(const std::vector<std::string>& residues, std::vector<int>& scores,
                              std::vector<double>& distances, std::string* error) {
  tsq_params p;
  fill(p, *this);
  tsq_ctx* c = nullptr;
  int rc = tsq_create(&c, &p);
  if (rc != TSQ_OK) {
    if (error) *error = tsq_status_string(rc);
    return rc;
  }
  std::vector<const char*> ptr(residues.size());
  std::vector<uint32_t> len(residues.size());
  for (size_t i = 0; i < residues.size(); i++) {
    ptr[i] = residues[i].data();
    len[i] = (uint32_t)residues[i].size();
  }
  rc = tsq_set_sequences(c, ptr.data(), len.data(), (uint32_t)residues.size());
  if (rc == TSQ_OK) rc = tsq_run(c, nullptr, nullptr, nullptr);
  const int32_t* s = nullptr;
  const double* d = nullptr;
  uint64_t cnt = 0;
  if (rc == TSQ_OK) rc = tsq_scores(c, &s, &cnt);
  if (rc == TSQ_OK) rc = tsq_distances(c, &d, &cnt);
  if (rc == TSQ_OK) {
    scores.assign(s, s + cnt);
    distances.assign(d, d + cnt);
  } else if (error) {
    *error = tsq_last_error(c);
  }
  tsq_destroy(c);
  return rc;
}

namespace {
int load(tsq_ctx* c, const std::vector<std::string>& residues) {
  std::vector<const char*> ptr(residues.size());
  std::vector<uint32_t> len(residues.size());
  for (size_t i = 0; i < residues.size(); i++) {
    ptr[i] = residues[i].data();
    len[i] = (uint32_t)residues[i].size();
  }
  return tsq_set_sequences(c, ptr.data(), len.data(), (uint32_t)residues.size());
}
}  // namespace

int B200Gotoh::guideTree(const std::vector<std::string>& residues, const std::vector<std::string>& labels,
                         const std::string& newickPath, std::string* error) {
  tsq_params p;
  fill(p, *this);
  tsq_ctx* c = nullptr;
  int rc = tsq_create(&c, &p);
  if (rc != TSQ_OK) {
    if (error) *error = tsq_status_string(rc);
    return rc;
  }
  rc = load(c, residues);
  if (rc == TSQ_OK) rc = tsq_run(c, nullptr, nullptr, nullptr);
  if (rc == TSQ_OK) {
    std::vector<const char*> lab(labels.size());
    for (size_t i = 0; i < labels.size(); i++) lab[i] = labels[i].c_str();
    rc = tsq_write_newick(c, labels.size() == residues.size() ? lab.data() : nullptr, newickPath.c_str());
  }
  if (rc != TSQ_OK && error) *error = tsq_last_error(c);
  tsq_destroy(c);
  return rc;
}

int B200Gotoh::multipleAlignment(const std::vector<std::string>& residues, std::vector<std::string>& rows,
                                 std::vector<unsigned>& treeOrder, std::string* error) {
  tsq_params p;
  fill(p, *this);
  tsq_ctx* c = nullptr;
  int rc = tsq_create(&c, &p);
  if (rc != TSQ_OK) {
    if (error) *error = tsq_status_string(rc);
    return rc;
  }
  rc = load(c, residues);
  if (rc == TSQ_OK) rc = tsq_run(c, nullptr, nullptr, nullptr);
  const char* flat = nullptr;
  const uint32_t* order = nullptr;
  uint32_t n = 0, cols = 0;
  if (rc == TSQ_OK) rc = tsq_msa(c, &flat, &n, &cols, &order);
  if (rc == TSQ_OK) {
    rows.clear();
    for (uint32_t r = 0; r < n; r++) rows.emplace_back(flat + (size_t)r * cols, cols);
    treeOrder.assign(order, order + n);
  } else if (error) {
    *error = tsq_last_error(c);
  }
  tsq_destroy(c);
  return rc;
}

int B200Gotoh::pairwiseAlignment(const std::string& a, const std::string& b, std::string& rowA, std::string& rowB, int& score,
                                 std::string* error) {
  tsq_params p;
  fill(p, *this);
  tsq_ctx* c = nullptr;
  int rc = tsq_create(&c, &p);
  if (rc != TSQ_OK) {
    if (error) *error = tsq_status_string(rc);
    return rc;
  }
  const char* ptr[2] = {a.data(), b.data()};
  const uint32_t len[2] = {(uint32_t)a.size(), (uint32_t)b.size()};
  const uint32_t cap = len[0] + len[1] + 1;
  std::vector<char> ra(cap), rb(cap);
  uint32_t cols = 0;
  int32_t sc = 0;
  rc = tsq_set_sequences(c, ptr, len, 2);
  if (rc == TSQ_OK) rc = tsq_upload(c);
  if (rc == TSQ_OK) rc = tsq_align_pair(c, 0, 1, ra.data(), rb.data(), cap, &cols, &sc);
  if (rc == TSQ_OK) {
    rowA.assign(ra.data(), cols);
    rowB.assign(rb.data(), cols);
    score = sc;
  } else if (error) {
    *error = tsq_last_error(c);
  }
  tsq_destroy(c);
  return rc;
}

int B200Gotoh::consensus(const std::vector<std::string>& rows, double plurality, std::string& out, std::string* error) {
  tsq_params p;
  fill(p, *this);
  tsq_ctx* c = nullptr;
  int rc = tsq_create(&c, &p);
  if (rc != TSQ_OK) {
    if (error) *error = tsq_status_string(rc);
    return rc;
  }
  const uint32_t ncols = rows.empty() ? 0u : (uint32_t)rows[0].size();
  std::vector<const char*> ptr(rows.size());
  for (size_t i = 0; i < rows.size(); i++) {
    if (rows[i].size() != ncols) rc = TSQ_ERR_INVALID;
    ptr[i] = rows[i].data();
  }
  out.assign(ncols, '?');
  if (rc == TSQ_OK) rc = tsq_consensus(c, ptr.data(), (uint32_t)rows.size(), ncols, plurality, ncols ? &out[0] : nullptr);
  if (rc != TSQ_OK && error) *error = rc == TSQ_ERR_INVALID ? "alignment rows differ in length" : tsq_last_error(c);
  tsq_destroy(c);
  return rc;
}

std::string filterCells(const std::vector<unsigned short>& cells, bool applyExclusions) {
  const unsigned short EXCLUDE_CELL = 0x0080, REMOVE_FLAGS = 0x007F;  // Sequence.h:36-39
  std::string r;
  for (unsigned short q : cells) {
    if ((q & EXCLUDE_CELL) && applyExclusions) continue;
    r.push_back((char)(q & REMOVE_FLAGS));
  }
  return r;
}

}  // namespace tsqhost
