/* qt_proc_dom.h -- TEST INFRASTRUCTURE, on top of qt_min.h: what the reference's tool wrappers
 * (tweakseq/Core/ClustalO.cpp, Muscle.cpp, MAFFT.cpp, AlignmentTool.cpp) need beyond strings.
 *   QProcess: start(program, args) really forks and execs the program with its stdout and stderr captured;
 *   waitForStarted / waitForReadyRead / waitForFinished, readAll() (= stdout, Qt's default read channel),
 *   readAllStandardError() -- enough for the wrappers' getVersion() to run a real executable.
 *   QThread::idealThreadCount();  QByteArray -> QString.
 *   QDomDocument / QDomElement / QDomText / QDomNodeList: a small functional DOM (handle semantics) for
 *   writeSettings() / readSettings().
 *   QObject / QThread / QMetaObject for host/qt/B200GotohTool: start() runs run() in place, queued
 *   invocations are recorded. */
#ifndef TSQ_REF_QT_PROC_DOM_H
#define TSQ_REF_QT_PROC_DOM_H
#include <memory>
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
  const char* constData() const { return s_.c_str(); }
  int size() const { return (int)s_.size(); }
  std::string s_;
};
inline QByteArray QString::toLocal8Bit() const { return QByteArray(toStd()); }
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

// A small functional DOM with Qt's handle semantics (a QDomNode is a shared handle onto a tree node): what
// writeSettings()/readSettings() of the tool wrappers do -- createElement, appendChild, createTextNode,
// elementsByTagName (document order), firstChildElement / nextSiblingElement, tagName, text.
struct QDomNodeData {
  QString tag, value;   // element: tag; text node: value
  bool is_text = false;
  std::vector<std::shared_ptr<QDomNodeData>> kids;
  QDomNodeData* parent = nullptr;
};

class QDomElement;
class QDomNode {
 public:
  QDomNode() {}
  explicit QDomNode(std::shared_ptr<QDomNodeData> d) : d_(d) {}
  bool isNull() const { return !d_; }
  inline QDomElement firstChildElement() const;
  inline QDomElement nextSiblingElement() const;
  QDomNode appendChild(const QDomNode& n) {
    if (d_ && n.d_) { n.d_->parent = d_.get(); d_->kids.push_back(n.d_); }
    return n;
  }
  std::shared_ptr<QDomNodeData> d_;
};
class QDomElement : public QDomNode {
 public:
  QDomElement() {}
  explicit QDomElement(std::shared_ptr<QDomNodeData> d) : QDomNode(d) {}
  QString tagName() const { return d_ ? d_->tag : QString(); }
  QString text() const {   // all text below this element, in document order
    QString t;
    if (d_) collect(d_.get(), t);
    return t;
  }
 private:
  static void collect(const QDomNodeData* n, QString& t) {
    for (auto& k : n->kids) { if (k->is_text) t += k->value; else collect(k.get(), t); }
  }
};
class QDomText : public QDomNode {
 public:
  explicit QDomText(std::shared_ptr<QDomNodeData> d) : QDomNode(d) {}
};
inline QDomElement QDomNode::firstChildElement() const {
  if (d_) for (auto& k : d_->kids) if (!k->is_text) return QDomElement(k);
  return QDomElement();
}
inline QDomElement QDomNode::nextSiblingElement() const {
  if (!d_ || !d_->parent) return QDomElement();
  auto& sib = d_->parent->kids;
  bool seen = false;
  for (auto& k : sib) {
    if (seen && !k->is_text) return QDomElement(k);
    if (k.get() == d_.get()) seen = true;
  }
  return QDomElement();
}
class QDomNodeList {
 public:
  int count() const { return (int)v_.size(); }
  QDomNode item(int i) const { return i >= 0 && (size_t)i < v_.size() ? QDomNode(v_[(size_t)i]) : QDomNode(); }
  std::vector<std::shared_ptr<QDomNodeData>> v_;
};
class QDomDocument : public QDomNode {
 public:
  QDomDocument() : QDomNode(std::make_shared<QDomNodeData>()) {}
  QDomElement createElement(const QString& tag) {
    auto d = std::make_shared<QDomNodeData>();
    d->tag = tag;
    return QDomElement(d);
  }
  QDomText createTextNode(const QString& v) {
    auto d = std::make_shared<QDomNodeData>();
    d->is_text = true;
    d->value = v;
    return QDomText(d);
  }
  QDomNodeList elementsByTagName(const QString& tag) const {
    QDomNodeList l;
    walk(d_, tag, l);
    return l;
  }
 private:
  static void walk(const std::shared_ptr<QDomNodeData>& n, const QString& tag, QDomNodeList& l) {
    for (auto& k : n->kids) {
      if (!k->is_text && k->tag == tag) l.v_.push_back(k);
      walk(k, tag, l);
    }
  }
};

// QObject / QThread / QMetaObject as far as host/qt/B200GotohTool.{h,cpp} uses them.  No event loop: start()
// runs run() on the calling thread, a queued invokeMethod() is recorded for the test to read.
#define Q_OBJECT
#define signals public
#define slots
#define emit
namespace Qt { enum ConnectionType { AutoConnection, DirectConnection, QueuedConnection }; }

class QObject {
 public:
  explicit QObject(QObject* parent = nullptr) : parent_(parent) {}
  virtual ~QObject() {}
  QObject* parent() const { return parent_; }
 private:
  QObject* parent_;
};

class QGenericArgument {
 public:
  QGenericArgument(const char* name = nullptr, const void* data = nullptr) : name_(name), data_(data) {}
  const char* name_;
  const void* data_;
};
#define Q_ARG(type, data) QGenericArgument(#type, static_cast<const void*>(&static_cast<const type&>(data)))

struct QtShimInvocation { QObject* receiver; std::string member, text; };
inline std::vector<QtShimInvocation>& qtShimInvocations() { static std::vector<QtShimInvocation> v; return v; }

class QMetaObject {
 public:
  static bool invokeMethod(QObject* obj, const char* member, Qt::ConnectionType, QGenericArgument a0 = QGenericArgument()) {
    std::string text;
    if (a0.data_ && a0.name_ && std::string(a0.name_) == "QString") text = static_cast<const QString*>(a0.data_)->toStd();
    qtShimInvocations().push_back({obj, member ? member : "", text});
    return true;
  }
};

class QThread : public QObject {
 public:
  explicit QThread(QObject* parent = nullptr) : QObject(parent) {}
  static int idealThreadCount() { return (int)sysconf(_SC_NPROCESSORS_ONLN); }
  void start() { run(); }
  bool wait(unsigned long = ~0ul) { return true; }
 protected:
  virtual void run() {}
};
#endif
