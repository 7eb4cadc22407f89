// qt_adapter_driver.cpp -- TEST INFRASTRUCTURE: runs this repository's Qt adapter (host/qt/B200GotohTool.{h,cpp},
// SURVEY 8f-3) -- the REAL file a maintainer drops into tweakseq/Core, not its Qt-free twin -- against the
// reference's own AlignmentTool.h / AlignmentTool.cpp / ClustalO.cpp and the functional Qt stand-ins of
// qt_proc_dom.h: settings written and read in ONE document together with the reference's ClustalO, and the
// QThread worker run through tsq_run_fasta.  Built by oracle/Makefile into oracle/_ref/libref_qt_adapter.so.
#include <cstring>
#include <string>

#include "qt_proc_dom.h"
#include "B200GotohTool.h"
#include "ClustalO.h"
#include "XMLHelper.h"

// glue, not under test: XMLHelper::addElement as tweakseq/Core/XMLHelper.cpp:51-59 has it (element + text child)
QDomElement XMLHelper::addElement(QDomDocument& doc, QDomElement& root, QString tag, QString txt) {
  QDomElement e = doc.createElement(tag);
  root.appendChild(e);
  e.appendChild(doc.createTextNode(txt));
  return e;
}

// what moc would generate for the worker's signals: deliver to the "connected slot" (here: remember the values)
static int g_exit_code = -999, g_exit_status = -999;
void B200GotohWorker::message(const QString&) {}
void B200GotohWorker::finished(int code, int status) { g_exit_code = code; g_exit_status = status; }

static int put(const std::string& s, char* out, unsigned long cap) {
  if (s.size() + 1 > cap) return -2;
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}

// Both tools write their settings into one <settings> element (Project::writeSettings, Project.cpp:853-862), two
// fresh tools read them back (Project.cpp:1161-1182).  out: "key=value" lines for the test to assert on.
extern "C" int tsq_qt_settings_round_trip(char* out, unsigned long cap) {
  QDomDocument doc;
  QDomElement root = doc.createElement("settings");
  doc.appendChild(root);
  ClustalO c;
  c.setExecutable("/opt/somewhere/clustalo");
  c.setPreferred(false);
  B200GotohTool b;
  b.setPreferred(true);
  b.gapOpen = 9; b.gapExtend = 2; b.device = 1; b.alignInProcess = false;
  c.writeSettings(doc, root);
  b.writeSettings(doc, root);
  ClustalO c2;
  B200GotohTool b2;
  c2.readSettings(doc);    // must not pick up anything of the b200gotoh element
  b2.readSettings(doc);    // must not pick up anything of the clustalo element
  std::string s;
  s += "elements=" + std::to_string(doc.elementsByTagName("alignment_tool").count()) + "\n";
  s += "clustalo.path=" + c2.executable().toStd() + "\n";
  s += std::string("clustalo.preferred=") + (c2.preferred() ? "yes" : "no") + "\n";
  s += "b200.name=" + b2.name().toStd() + "\n";
  s += "b200.path=" + b2.executable().toStd() + "\n";
  s += std::string("b200.preferred=") + (b2.preferred() ? "yes" : "no") + "\n";
  s += "b200.gap_open=" + std::to_string(b2.gapOpen) + "\n";
  s += "b200.gap_extend=" + std::to_string(b2.gapExtend) + "\n";
  s += "b200.device=" + std::to_string(b2.device) + "\n";
  s += std::string("b200.align_in_process=") + (b2.alignInProcess ? "yes" : "no") + "\n";
  s += "b200.version=" + b2.version().toStd() + "\n";
  s += std::string("b200.in_process=") + (b2.inProcess() ? "yes" : "no") + "\n";
  QString fin("in.fa"), fout("out.fa"), exec;
  QStringList args;
  b2.makeCommand(fin, fout, exec, args);
  s += "b200.argc=" + std::to_string(args.size()) + "\n";
  return put(s, out, cap);
}

// The worker as SeqEditMainWin::startAlignment would use it: construct, start(), and
// read what it emitted through finished(int, int); log lines arrive as queued addMessage(QString) invocations.
extern "C" int tsq_qt_worker_run(const char* fin, const char* fout, int align_in_process, int* exit_code, int* exit_status,
                                 char* log, unsigned long cap) {
  QObject main_window;
  B200GotohTool tool;
  tool.alignInProcess = align_in_process != 0;
  g_exit_code = g_exit_status = -999;
  qtShimInvocations().clear();
  B200GotohWorker w(&tool, QString::fromStd(fin), QString::fromStd(fout), &main_window);
  w.start();
  w.wait();
  *exit_code = g_exit_code;
  *exit_status = g_exit_status;
  std::string s;
  for (auto& inv : qtShimInvocations())   // every log line leaves as the worker's own message(QString) signal
    if (inv.receiver == &w && inv.member == "message") s += inv.text + "\n";
  return put(s, log, cap);
}

// The in-memory route: labels and residues as startAlignment takes them from the
// model -- label and Sequence::filter(true) -- `n` of each, '\n'-separated; fout = the alignment, headers ">label".
extern "C" int tsq_qt_worker_run_in_memory(const char* labels_nl, const char* residues_nl, const char* fout, int* exit_code,
                                           int* exit_status, char* log, unsigned long cap) {
  auto split = [](const char* s) {
    QStringList out;
    std::string cur;
    for (const char* p = s; *p; ++p) {
      if (*p == '\n') { out << QString::fromStd(cur); cur.clear(); }
      else cur.push_back(*p);
    }
    return out;
  };
  QObject main_window;
  B200GotohTool tool;
  g_exit_code = g_exit_status = -999;
  qtShimInvocations().clear();
  B200GotohWorker w(&tool, split(labels_nl), split(residues_nl), QString::fromStd(fout), &main_window);
  w.start();
  w.wait();
  *exit_code = g_exit_code;
  *exit_status = g_exit_status;
  std::string s;
  for (auto& inv : qtShimInvocations())
    if (inv.receiver == &w && inv.member == "message") s += inv.text + "\n";
  return put(s, log, cap);
}
