/* qt_proc_dom.h -- TEST INFRASTRUCTURE, on top of qt_min.h: what the reference's tool wrappers
 * (tweakseq/Core/ClustalO.cpp, Muscle.cpp, MAFFT.cpp, AlignmentTool.cpp) need beyond strings.
 *   QProcess: start(program, args) really forks and execs the program with its stdout and stderr captured;
 *   waitForStarted / waitForReadyRead / waitForFinished, readAll() (= stdout, Qt's default read channel),
 *   readAllStandardError() -- enough for the wrappers' getVersion() to run a real executable.
 *   QThread::idealThreadCount();  QByteArray -> QString.
 *   QDomDocument / QDomElement / QDomNodeList: an EMPTY document (elementsByTagName finds nothing), so
 *   readSettings() falls through to getVersion(); writeSettings() is not exercised. */
#ifndef TSQ_REF_QT_PROC_DOM_H
#define TSQ_REF_QT_PROC_DOM_H
#include <string>
#include <vector>

#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "qt_min.h"

class QByteArray {
 public:
  QByteArray() {}
  explicit QByteArray(const std::string& s) : s_(s) {}
  std::string s_;
};
inline QString::QString(const QByteArray& b) { for (unsigned char c : b.s_) d_.push_back(c); }

class QProcess {
 public:
  QProcess() : started_(false), done_(false), pid_(-1), out_(-1), err_(-1) {}
  ~QProcess() { finish(); }
  void start(const QString& program, const QStringList& args) {
    int po[2], pe[2];
    if (pipe(po) != 0 || pipe(pe) != 0) return;
    std::vector<std::string> store;
    store.push_back(program.toStd());
    for (int i = 0; i < args.size(); i++) store.push_back(args.at(i).toStd());
    std::vector<char*> argv;
    for (auto& s : store) argv.push_back(const_cast<char*>(s.c_str()));
    argv.push_back(nullptr);
    pid_ = fork();
    if (pid_ == 0) {
      dup2(po[1], 1); dup2(pe[1], 2);
      close(po[0]); close(po[1]); close(pe[0]); close(pe[1]);
      execv(argv[0], argv.data());
      _exit(127);
    }
    close(po[1]); close(pe[1]);
    out_ = po[0]; err_ = pe[0];
    started_ = pid_ > 0;
  }
  bool waitForStarted(int = 30000) { return started_; }
  bool waitForReadyRead(int = 30000) { finish(); return !stdout_.empty() || !stderr_.empty(); }
  bool waitForFinished(int = 30000) { finish(); return started_; }
  QByteArray readAll() { finish(); QByteArray b(stdout_); stdout_.clear(); return b; }
  QByteArray readAllStandardError() { finish(); QByteArray b(stderr_); stderr_.clear(); return b; }
  int exitCode() { finish(); return code_; }
 private:
  void drain(int fd, std::string& into) {
    char buf[4096];
    ssize_t n;
    while (fd >= 0 && (n = read(fd, buf, sizeof buf)) > 0) into.append(buf, (size_t)n);
    if (fd >= 0) close(fd);
  }
  void finish() {
    if (!started_ || done_) return;
    drain(out_, stdout_); drain(err_, stderr_);
    int st = 0;
    waitpid(pid_, &st, 0);
    code_ = WIFEXITED(st) ? WEXITSTATUS(st) : -1;
    done_ = true;
  }
  bool started_, done_;
  pid_t pid_;
  int out_, err_, code_ = -1;
  std::string stdout_, stderr_;
};

class QThread { public: static int idealThreadCount() { return (int)sysconf(_SC_NPROCESSORS_ONLN); } };

class QDomElement;
class QDomNode {
 public:
  bool isNull() const { return true; }
  inline QDomElement firstChildElement() const;
  inline QDomElement nextSiblingElement() const;
  QDomNode appendChild(const QDomNode& n) { return n; }
};
class QDomElement : public QDomNode {
 public:
  QString tagName() const { return QString(); }
  QString text() const { return QString(); }
};
inline QDomElement QDomNode::firstChildElement() const { return QDomElement(); }
inline QDomElement QDomNode::nextSiblingElement() const { return QDomElement(); }
class QDomNodeList {
 public:
  int count() const { return 0; }
  QDomNode item(int) const { return QDomNode(); }
};
class QDomDocument : public QDomNode {
 public:
  QDomElement createElement(const QString&) { return QDomElement(); }
  QDomNodeList elementsByTagName(const QString&) const { return QDomNodeList(); }
};
#endif
