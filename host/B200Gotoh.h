// B200Gotoh.h -- the in-process alignment tool: all-vs-all Gotoh distance matrix on a B200.
//
// Shaped like tweakseq/Core/ClustalO.{h,cpp} (name/path/preferred settings element, version
// probe, makeCommand), but run() calls libtsqb200.so through the C ABI (include/tsq_b200.h)
// instead of handing an argv to QProcess (SeqEditMainWin.cpp:1654-1660).
#ifndef TSQ_HOST_B200GOTOH_H
#define TSQ_HOST_B200GOTOH_H

#include "AlignmentTool.h"

namespace tsqhost {

class B200Gotoh : public AlignmentTool {
 public:
  B200Gotoh();
  ~B200Gotoh() override;

  void makeCommand(std::string& fin, std::string& fout, std::string& exec,
                   std::vector<std::string>& arglist) override;
  void writeSettings(SettingsDocument& doc) override;
  void readSettings(SettingsDocument& doc) override;
  bool inProcess() override { return true; }
  int run(const std::string& fin, const std::string& fout, const LogSink& log, CancelFlag* cancel) override;

  // In-memory path: residues exactly as Sequence::filter(true) returns them
  // (Sequence.cpp:57-69).  Packed upper-triangle results, submitted order.
  int distanceMatrix(const std::vector<std::string>& residues, std::vector<int>& scores,
                     std::vector<double>& distances, std::string* error = nullptr);

  // Guide tree of the same job: Newick text for clustalo --guidetree-in.
  int guideTree(const std::vector<std::string>& residues, const std::vector<std::string>& labels,
                const std::string& newickPath, std::string* error = nullptr);
  // Distances, guide tree and the progressive alignment along it, in memory: equal-length gapped rows in
  // submitted order (what Project::readNewAlignment, Project.cpp:908-1032, takes from the aligner's
  // output file) and the tree order of the rows.
  int multipleAlignment(const std::vector<std::string>& residues, std::vector<std::string>& rows,
                        std::vector<unsigned>& treeOrder, std::string* error = nullptr);
  // One optimal global alignment of two sequences with its path: two gapped rows.
  int pairwiseAlignment(const std::string& a, const std::string& b, std::string& rowA, std::string& rowB, int& score,
                        std::string* error = nullptr);
  // Consensus annotation of an alignment: Consensus::calculate (Consensus.cpp:80-161) on the GPU.
  int consensus(const std::vector<std::string>& alignedRows, double plurality, std::string& out,
                std::string* error = nullptr);

  int gapOpen = -1, gapExtend = -1, device = 0;  // <0: library defaults (11/1 protein)
  int devices = 1;                               // how many B200s of the box one job uses (tsq_params.n_devices;
                                                 // -1 = all): the pair space is cut into that many slabs
  bool nucleotide = false;
  bool detectAlphabet = false;                   // run(): decide protein / nucleotide from the file's residues
                                                 // (a project holds either kind: SequenceFile::DNA / ::Proteins)
  bool identityDistance = false;                 // ClustalW-style 1 - identities/min(len)
  bool alignInProcess = true;                    // run(): fout = the multiple alignment (FASTA, tree order) that
                                                 // readNewAlignment ingests;
                                                 // false: fout = the distance matrix for clustalo --distmat-in

 private:
  void init();
  void getVersion();
};

// Sequence::filter for 16-bit residue cells carrying tweakseq's flag bits (Sequence.h:36-39).
std::string filterCells(const std::vector<unsigned short>& cells, bool applyExclusions);

}  // namespace tsqhost
#endif
