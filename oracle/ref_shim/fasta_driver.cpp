// fasta_driver.cpp -- TEST INFRASTRUCTURE: C entry points around the reference's own FASTAFile class
// (tweakseq/Core/FASTAFile.cpp + SequenceFile.cpp compiled where they lie by oracle/Makefile into
// oracle/_ref/libref_fasta.so): the reader tweakseq applies to the aligner's output
// (Project::readNewAlignment, Project.cpp:908-915) and the writer it applies to the aligner's input
// (Project::exportFASTA, Project.cpp:870-881).  Lists cross the boundary as '\n'-joined text.
#include <cstring>
#include <string>

#include "qt_min.h"
#include "FASTAFile.h"

static std::string join(QStringList& l) {
  std::string s;
  for (int i = 0; i < l.size(); i++) { s += l.at(i).toStd(); s.push_back('\n'); }
  return s;
}

static QStringList split(const char* text) {
  QStringList l;
  std::string cur;
  for (const char* p = text; *p; p++) {
    if (*p == '\n') { l << QString::fromStd(cur); cur.clear(); }
    else cur.push_back(*p);
  }
  return l;
}

// labels / seqs / comments: caller buffers of `cap` bytes each, filled with '\n'-terminated entries.
// Returns the number of labels, or -1 (read failed) / -2 (a buffer is too small).
extern "C" int tsq_ref_fasta_read(const char* path, char* labels, char* seqs, char* comments, unsigned long cap) {
  FASTAFile f(QString::fromStd(path));
  QStringList l, s, c;
  if (!f.read(l, s, c)) return -1;
  const std::string a = join(l), b = join(s), d = join(c);
  if (a.size() + 1 > cap || b.size() + 1 > cap || d.size() + 1 > cap) return -2;
  memcpy(labels, a.c_str(), a.size() + 1);
  memcpy(seqs, b.c_str(), b.size() + 1);
  memcpy(comments, d.c_str(), d.size() + 1);
  return l.size();
}

extern "C" int tsq_ref_fasta_write(const char* path, const char* labels, const char* seqs, const char* comments) {
  FASTAFile f(QString::fromStd(path));
  QStringList l = split(labels), s = split(seqs), c = split(comments);
  return f.write(l, s, c) ? 0 : -1;
}
