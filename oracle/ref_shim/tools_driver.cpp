// tools_driver.cpp -- TEST INFRASTRUCTURE: C entry points around the reference's own tool wrappers
// (tweakseq/Core/ClustalO.cpp, Muscle.cpp, MAFFT.cpp, AlignmentTool.cpp compiled where they lie into
// oracle/_ref/libref_tools.so): the argv tweakseq hands to QProcess::start (SeqEditMainWin.cpp:1654-1660)
// comes out of THEIR makeCommand(), and the version string out of THEIR getVersion() run on a real executable.
#include <cstring>
#include <string>

#include "qt_proc_dom.h"
#include "ClustalO.h"
#include "MAFFT.h"
#include "Muscle.h"
#include "XMLHelper.h"

// writeSettings() is not exercised here; the wrappers only need the symbol to link
QDomElement XMLHelper::addElement(QDomDocument&, QDomElement&, QString, QString) { return QDomElement(); }

static AlignmentTool* make(int tool) {
  if (tool == 0) return new ClustalO();
  if (tool == 1) return new Muscle();
  if (tool == 2) return new MAFFT();
  return nullptr;
}

static int put(const std::string& s, char* out, unsigned long cap) {
  if (s.size() + 1 > cap) return -2;
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}

// tool: 0 ClustalO, 1 MUSCLE, 2 MAFFT.  out: name, executable, then one argv entry per line; *uses_stdout as the wrapper says.
extern "C" int tsq_ref_tool_command(int tool, const char* exe, const char* fin, const char* fout, char* out, unsigned long cap,
                                    int* uses_stdout) {
  AlignmentTool* t = make(tool);
  if (!t) return -1;
  if (exe && *exe) t->setExecutable(QString::fromStd(exe));
  QString qin = QString::fromStd(fin), qout = QString::fromStd(fout), qexec;
  QStringList args;
  t->makeCommand(qin, qout, qexec, args);
  std::string s = t->name().toStd() + "\n" + qexec.toStd() + "\n";
  for (int i = 0; i < args.size(); i++) s += args.at(i).toStd() + "\n";
  *uses_stdout = t->usesStdOut() ? 1 : 0;
  delete t;
  return put(s, out, cap);
}

// The wrapper's own version probe (readSettings() on an empty document ends in getVersion()) run on `exe`.
extern "C" int tsq_ref_tool_version(int tool, const char* exe, char* out, unsigned long cap) {
  AlignmentTool* t = make(tool);
  if (!t) return -1;
  t->setExecutable(QString::fromStd(exe));
  QDomDocument doc;
  t->readSettings(doc);
  const std::string v = t->version().toStd();
  delete t;
  return put(v, out, cap);
}
