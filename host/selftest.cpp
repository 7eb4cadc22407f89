// selftest.cpp -- exercises the C++ adapter.  Without a B200 it checks the interface mirror and
// that run() fails loudly (no fallback); with one (argv[1] = fasta in, argv[2] = matrix out) it
// runs the in-process tool end to end.
#include <cstdio>
#include <cstdlib>

#include "../include/tsq_b200.h"
#include "B200Gotoh.h"

using namespace tsqhost;

#define CHECK(x) do { if (!(x)) { printf("FAIL line %d: %s\n", __LINE__, #x); return 1; } } while (0)

int main(int argc, char** argv) {
  AlignmentTool base;
  CHECK(base.name().empty() && !base.preferred() && !base.usesStdOut() && !base.inProcess());
  B200Gotoh t;
  CHECK(t.name() == "b200gotoh" && t.inProcess() && !t.usesStdOut());
  t.setPreferred(true);
  t.gapOpen = 9;
  t.gapExtend = 2;
  t.alignInProcess = false;
  SettingsDocument doc;
  SettingsElement other;
  other.children = {{"name", "clustalo"}, {"path", "/usr/local/bin/clustalo"}, {"preferred", "no"}};
  doc.alignment_tools.push_back(other);
  t.writeSettings(doc);
  B200Gotoh u;
  CHECK(u.alignInProcess);   // the default: run() writes the alignment the editor ingests
  u.readSettings(doc);
  CHECK(u.preferred() && u.gapOpen == 9 && u.gapExtend == 2 && !u.alignInProcess && u.executable() == t.executable());
  CHECK(u.version().find("tsq-b200") != std::string::npos);
  std::string fin = "in.fa", fout = "out.mat", exec;
  std::vector<std::string> args;
  t.makeCommand(fin, fout, exec, args);
  CHECK(args.size() == 8 && args[6] == "--distmat-in");
  CHECK(filterCells({'A', (unsigned short)('C' | 0x80), '-', (unsigned short)('D' | 0x100)}, true) == "A-D");

  if (tsq_device_count() <= 0) {
    std::vector<int> s;
    std::vector<double> d;
    std::string err;
    CHECK(u.distanceMatrix({"ACD", "ACE"}, s, d, &err) == TSQ_ERR_NO_DEVICE);
    printf("host selftest ok (no B200 here: run() refuses, as it must)\n");
    return 0;
  }
  B200Gotoh g;
  std::vector<int> s;
  std::vector<double> d;
  std::string err;
  CHECK(g.distanceMatrix({"WWWW", "WWWW", "ACDEFG"}, s, d, &err) == TSQ_OK);
  CHECK(s.size() == 3 && s[0] == 44 && d[0] == 0.0);
  {
    std::string cons;
    CHECK(g.consensus({"AWC-a", "AWC-A", "AYD-A", "RWC-x"}, -1.0, cons, &err) == TSQ_OK);
    CHECK(cons == "AWC-?");
    std::string ra, rb;
    int sc = 0;
    CHECK(g.pairwiseAlignment("WWCWW", "WWWW", ra, rb, sc, &err) == TSQ_OK);
    CHECK(ra == "WWCWW" && rb.size() == 5 && sc == 44 - 12);   // one gap of length 1 against C
    std::vector<std::string> rows;
    std::vector<unsigned> order;
    CHECK(g.multipleAlignment({"WWCWW", "WWWW", "WWCWW"}, rows, order, &err) == TSQ_OK);
    CHECK(rows.size() == 3 && rows[0] == "WWCWW" && rows[2] == "WWCWW" && rows[1].size() == 5 && order.size() == 3);
    g.identityDistance = true;
    CHECK(g.distanceMatrix({"WWWW", "WWCWW", "ACDEFG"}, s, d, &err) == TSQ_OK);
    CHECK(d[0] == 0.0);   // 4 identities over the shorter length 4
    g.identityDistance = false;
  }
  if (argc >= 3) {
    g.alignInProcess = false;   // argv[2] = distance matrix out (the clustalo hand-off of SURVEY 8f-1)
    int rc = g.run(argv[1], argv[2], [](const std::string& l) { printf("[log] %s\n", l.c_str()); }, nullptr);
    CHECK(rc == 0);
    if (argc >= 4) {   // argv[3] = alignment out: the tool as a complete in-process aligner
      g.alignInProcess = true;
      rc = g.run(argv[1], argv[3], [](const std::string& l) { printf("[log] %s\n", l.c_str()); }, nullptr);
      CHECK(rc == 0);
    }
  }
  printf("host selftest ok (B200 path)\n");
  return 0;
}
