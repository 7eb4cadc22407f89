// consensus_driver.cpp -- TEST INFRASTRUCTURE: C entry point around the reference's own Consensus class
// (compiled from /root/reference/tweakseq/Core/Annotations/Consensus.cpp by oracle/Makefile into
// oracle/_ref/libref_consensus.so).  cells: nrows x ncols 16-bit residue cells as Sequence::residues holds
// them (flag bits and all); plurality < 0 = the default Consensus::setSequences computes (rows / 2).
// out: ncols characters = Consensus::sequence() & 0xff.
#include "qt_min.h"
#include "Consensus.h"
#include "Sequence.h"
#include "Sequences.h"

extern "C" int tsq_ref_consensus(const unsigned short* cells, unsigned nrows, unsigned ncols, double plurality, char* out) {
  if (nrows == 0) return -1;   // Consensus::calculate reads ss.at(0)
  Sequences seqs;
  std::vector<Sequence> store(nrows);
  for (unsigned r = 0; r < nrows; r++) {
    for (unsigned c = 0; c < ncols; c++) store[r].residues.append(QChar((int)cells[(size_t)r * ncols + c]));
    seqs.sequences().append(&store[r]);
  }
  Consensus cons;
  cons.setSequences(&seqs);
  if (plurality >= 0) cons.setPlurality(plurality);
  cons.calculate();
  QString& s = cons.sequence();
  if ((unsigned)s.length() != ncols) return -2;
  for (unsigned c = 0; c < ncols; c++) out[c] = (char)(((const QString&)s)[(int)c].unicode() & 0xff);
  return 0;
}
