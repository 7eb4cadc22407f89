/* qt_min.h -- TEST INFRASTRUCTURE.  The handful of Qt5 types the reference's Consensus.cpp, FASTAFile.cpp and
 * SequenceFile.cpp touch, with the Qt 5 semantics those files rely on, so that the reference's OWN sources can
 * be compiled where they lie (tweakseq/Core/...) into oracle/_ref/ and run beside this repository's code:
 *   QChar(int), unicode(), toLatin1() (0 beyond Latin-1);
 *   QString: length/size/isEmpty, at, operator[] returning a QCharRef whose assignment GROWS the string (Qt 5
 *   pads with spaces: Consensus::calculate fills an initially empty consensusSequence_ that way), trimmed()
 *   (QChar::isSpace: \t \n \v \f \r and space for Latin-1 text), indexOf(QChar, from), mid(pos, n) with
 *   n = -1 or past the end meaning "the rest", toLower(), + and ==;
 *   QList<T> / QStringList: at, size, append, <<, replace, contains(s, Qt::CaseInsensitive);
 *   QFile / QTextStream: open, atEnd, readLine() without the line terminator ("\n" or "\r\n"), << QString,
 *   << endl;  QFileInfo::suffix();  qDebug() << anything (discarded).
 * ASCII / Latin-1 text only (the tests stay inside it).  Nothing here is used by the product. */
#ifndef TSQ_REF_QT_MIN_H
#define TSQ_REF_QT_MIN_H
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

class QChar {
 public:
  QChar() : u_(0) {}
  QChar(int v) : u_((unsigned short)v) {}
  QChar(char c) : u_((unsigned char)c) {}
  unsigned short unicode() const { return u_; }
  char toLatin1() const { return u_ > 0xff ? 0 : (char)u_; }
 private:
  unsigned short u_;
};

class QString;
class QByteArray;
class QStringList;
class QRegExp {   // only the pattern the reference uses: "\\s+" (Muscle.cpp:108)
 public:
  explicit QRegExp(const char* pattern) : p_(pattern) {}
  std::string p_;
};
class QCharRef {
 public:
  QCharRef(QString& s, int i) : s_(s), i_(i) {}
  inline QCharRef& operator=(QChar c);
  QCharRef& operator=(int v) { return *this = QChar(v); }
  QCharRef& operator=(const QCharRef& o) { return *this = QChar((int)o.unicode()); }
  operator QChar() const { return QChar((int)unicode()); }
  inline unsigned short unicode() const;
 private:
  QString& s_;
  int i_;
};

class QString {
 public:
  QString() {}
  QString(const char* s) { while (s && *s) d_.push_back((unsigned short)(unsigned char)*s++); }
  inline QString(const QByteArray& b);   // defined in qt_proc_dom.h
  static QString fromStd(const std::string& s) { QString q; for (unsigned char c : s) q.d_.push_back(c); return q; }
  std::string toStd() const { std::string s; for (unsigned short u : d_) s.push_back((char)(u & 0xff)); return s; }
  int length() const { return (int)d_.size(); }
  int size() const { return (int)d_.size(); }
  bool isEmpty() const { return d_.empty(); }
  QCharRef operator[](int i) { return QCharRef(*this, i); }
  const QChar operator[](int i) const { return QChar((int)d_[(size_t)i]); }
  const QChar at(int i) const { return QChar((int)d_[(size_t)i]); }
  void append(QChar c) { d_.push_back(c.unicode()); }
  QString& operator+=(const QString& o) { d_.insert(d_.end(), o.d_.begin(), o.d_.end()); return *this; }
  QString trimmed() const {
    auto sp = [](unsigned short u) { return u == ' ' || (u >= 9 && u <= 13); };
    size_t a = 0, b = d_.size();
    while (a < b && sp(d_[a])) a++;
    while (b > a && sp(d_[b - 1])) b--;
    QString q; q.d_.assign(d_.begin() + (long)a, d_.begin() + (long)b); return q;
  }
  int indexOf(QChar c, int from = 0) const {
    if (from < 0) from = 0;
    for (size_t i = (size_t)from; i < d_.size(); i++) if (d_[i] == c.unicode()) return (int)i;
    return -1;
  }
  QString mid(int pos, int n = -1) const {
    QString q;
    if (pos < 0) { if (n >= 0) n += pos; pos = 0; }
    if ((size_t)pos >= d_.size()) return q;
    size_t end = d_.size();
    if (n >= 0 && (size_t)pos + (size_t)n < end) end = (size_t)pos + (size_t)n;
    q.d_.assign(d_.begin() + pos, d_.begin() + (long)end);
    return q;
  }
  QString& remove(int pos, int n) {   // Qt: out-of-range positions do nothing, n is clipped to the end
    if (pos < 0 || (size_t)pos >= d_.size() || n <= 0) return *this;
    const size_t end = (size_t)pos + (size_t)n < d_.size() ? (size_t)pos + (size_t)n : d_.size();
    d_.erase(d_.begin() + pos, d_.begin() + (long)end);
    return *this;
  }
  inline QStringList split(const QRegExp& re) const;   // KeepEmptyParts, as Qt's default
  static QString number(int v) { return fromStd(std::to_string(v)); }
  static QString fromUtf8(const char* s) { return QString(s); }
  int toInt(bool* ok = nullptr) const {
    const std::string t = trimmed().toStd();
    char* end = nullptr;
    const long v = strtol(t.c_str(), &end, 10);
    const bool good = !t.empty() && end && *end == 0;
    if (ok) *ok = good;
    return good ? (int)v : 0;
  }
  inline QByteArray toLocal8Bit() const;   // defined in qt_proc_dom.h
  QString toLower() const { QString q = *this; for (auto& u : q.d_) if (u >= 'A' && u <= 'Z') u = (unsigned short)(u + 32); return q; }
  bool operator==(const QString& o) const { return d_ == o.d_; }
  bool operator!=(const QString& o) const { return d_ != o.d_; }
  bool operator==(const char* s) const { return *this == QString(s); }
  std::vector<unsigned short> d_;
};
inline QString operator+(const QString& a, const QString& b) { QString q = a; q += b; return q; }
inline QString operator+(const char* a, const QString& b) { QString q(a); q += b; return q; }
inline QString operator+(const QString& a, const char* b) { QString q = a; q += QString(b); return q; }

inline QCharRef& QCharRef::operator=(QChar c) {
  if (i_ >= (int)s_.d_.size()) s_.d_.resize((size_t)i_ + 1, (unsigned short)' ');
  s_.d_[(size_t)i_] = c.unicode();
  return *this;
}
inline unsigned short QCharRef::unicode() const { return i_ < (int)s_.d_.size() ? s_.d_[(size_t)i_] : 0; }

namespace Qt { enum CaseSensitivity { CaseInsensitive, CaseSensitive }; }

template <typename T>
class QList {
 public:
  const T& at(int i) const { return v_[(size_t)i]; }
  int size() const { return (int)v_.size(); }
  void append(const T& t) { v_.push_back(t); }
  void replace(int i, const T& t) { v_[(size_t)i] = t; }
  QList<T>& operator<<(const T& t) { v_.push_back(t); return *this; }
 protected:
  std::vector<T> v_;
};

class QStringList : public QList<QString> {
 public:
  QStringList& operator<<(const QString& s) { v_.push_back(s); return *this; }
  QStringList& operator<<(const char* s) { v_.push_back(QString(s)); return *this; }
  bool contains(const QString& s, Qt::CaseSensitivity cs = Qt::CaseSensitive) const {
    for (const QString& x : v_)
      if (cs == Qt::CaseSensitive ? x == s : x.toLower() == s.toLower()) return true;
    return false;
  }
};

inline QStringList QString::split(const QRegExp& re) const {
  QStringList out;
  if (re.p_ != "\\s+") return out;   // nothing else is needed here
  auto sp = [](unsigned short u) { return u == ' ' || (u >= 9 && u <= 13); };
  QString cur;
  size_t i = 0;
  while (i < d_.size()) {
    if (sp(d_[i])) {
      out << cur;
      cur = QString();
      while (i < d_.size() && sp(d_[i])) i++;
    } else {
      cur.append(QChar((int)d_[i++]));
    }
  }
  out << cur;
  return out;
}

class QIODevice {
 public:
  enum OpenModeFlag { ReadOnly = 1, WriteOnly = 2, Text = 16 };
};
inline int operator|(QIODevice::OpenModeFlag a, QIODevice::OpenModeFlag b) { return (int)a | (int)b; }

class QFile : public QIODevice {
 public:
  explicit QFile(const QString& name) : name_(name.toStd()), f_(nullptr) {}
  ~QFile() { close(); }
  bool open(int mode) { f_ = fopen(name_.c_str(), (mode & WriteOnly) ? "w" : "r"); return f_ != nullptr; }
  void close() { if (f_) fclose(f_); f_ = nullptr; }
  FILE* f() { return f_; }
 private:
  std::string name_;
  FILE* f_;
};

class QTextStream {
 public:
  explicit QTextStream(QFile* f) : f_(f) {}
  bool atEnd() {
    const int c = fgetc(f_->f());
    if (c == EOF) return true;
    ungetc(c, f_->f());
    return false;
  }
  QString readLine() {   // one line without its terminator ("\n" or "\r\n")
    std::string s;
    int c;
    while ((c = fgetc(f_->f())) != EOF && c != '\n') s.push_back((char)c);
    if (!s.empty() && s.back() == '\r') s.pop_back();
    return QString::fromStd(s);
  }
  QTextStream& operator<<(const QString& s) { const std::string t = s.toStd(); fwrite(t.data(), 1, t.size(), f_->f()); return *this; }
  QTextStream& operator<<(QTextStream& (*m)(QTextStream&)) { return m(*this); }
  void put(char c) { fputc(c, f_->f()); }
 private:
  QFile* f_;
};
inline QTextStream& endl(QTextStream& s) { s.put('\n'); return s; }

class QFileInfo {
 public:
  explicit QFileInfo(const QString& path) : p_(path.toStd()) {}
  QString suffix() const {   // after the last '.' of the file name
    const size_t slash = p_.find_last_of('/');
    const std::string file = slash == std::string::npos ? p_ : p_.substr(slash + 1);
    const size_t dot = file.find_last_of('.');
    return dot == std::string::npos ? QString() : QString::fromStd(file.substr(dot + 1));
  }
 private:
  std::string p_;
};

struct QDebugSink {
  template <typename T> QDebugSink& operator<<(const T&) { return *this; }
};
inline QDebugSink qDebug() { return QDebugSink(); }
#endif
