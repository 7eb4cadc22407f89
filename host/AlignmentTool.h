// AlignmentTool.h -- Qt-free twin of tweakseq's tool-wrapper interface.
//
// Same members, same defaults as tweakseq/Core/AlignmentTool.h:36-71 and AlignmentTool.cpp:43-63,
// with std::string / std::vector standing in for QString / QStringList and a tiny key/value tree
// standing in for QDomDocument, so the adapter can be compiled and tested where Qt5 is absent
// (this image).  The two members marked [ext] are the minimal extension SURVEY.md section 8b
// proposes to admit an in-process backend; host/qt/B200GotohTool.{h,cpp} is the same class
// written against the real Qt types for a tweakseq maintainer to drop in.
#ifndef TSQ_HOST_ALIGNMENT_TOOL_H
#define TSQ_HOST_ALIGNMENT_TOOL_H

#include <functional>
#include <string>
#include <utility>
#include <vector>

namespace tsqhost {

// Stand-in for the <alignment_tool> elements of a settings document (ClustalO.cpp:54-86).
struct SettingsElement {
  std::vector<std::pair<std::string, std::string>> children;  // tag -> text, in document order
};
struct SettingsDocument {
  std::vector<SettingsElement> alignment_tools;
};

using LogSink = std::function<void(const std::string&)>;  // -> MessageWin::addMessage
using CancelFlag = volatile int;                          // set by alignmentStop()

class AlignmentTool {
 public:
  AlignmentTool() : preferred_(false), usesStdOut_(false) {}
  virtual ~AlignmentTool() {}

  std::string name() { return name_; }
  std::string version() { return version_; }
  std::string executable() { return executable_; }
  void setExecutable(std::string e) { executable_ = e; }
  void setPreferred(bool pref) { preferred_ = pref; }
  bool preferred() { return preferred_; }
  bool usesStdOut() { return usesStdOut_; }  // the alignment is written to stdout

  virtual void makeCommand(std::string& /*fin*/, std::string& /*fout*/, std::string& /*exec*/,
                           std::vector<std::string>& /*arglist*/) {}
  virtual void writeSettings(SettingsDocument&) {}
  virtual void readSettings(SettingsDocument&) {}

  // [ext] in-process tools run inside the editor instead of through QProcess
  virtual bool inProcess() { return false; }
  // [ext] returns the "exit code" alignmentFinished() expects: 0 = success
  virtual int run(const std::string& /*fin*/, const std::string& /*fout*/, const LogSink&, CancelFlag*) {
    return -1;
  }

 protected:
  std::string name_;
  std::string version_;
  std::string executable_;
  bool preferred_;
  bool usesStdOut_;
};

}  // namespace tsqhost
#endif
