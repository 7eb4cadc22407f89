// sequence_driver.cpp -- TEST INFRASTRUCTURE: C entry point around the reference's own Sequence::filter
// (tweakseq/Core/Sequence.cpp:57-69, compiled where it lies into oracle/_ref/libref_sequence.so): what
// Project::exportFASTA applies to every sequence before the aligner sees it (Project.cpp:870-881).
// cells: n 16-bit residue cells with tweakseq's flag bits; out: up to n cells; returns the filtered length.
#include "qt_min.h"
#include "Sequence.h"

extern "C" int tsq_ref_filter(const unsigned short* cells, unsigned n, int apply_exclusions, unsigned short* out) {
  QString r;
  for (unsigned i = 0; i < n; i++) r.append(QChar((int)cells[i]));
  Sequence s(QString("label"), r);
  const QString f = s.filter(apply_exclusions != 0);
  for (int i = 0; i < f.size(); i++) out[i] = f.at(i).unicode();
  return f.size();
}
