/* qt_min.h -- TEST INFRASTRUCTURE.  The handful of Qt5 types the reference's Consensus.cpp touches, with
 * the Qt 5 semantics that file relies on, so that the reference's OWN source can be compiled where it lies
 * (tweakseq/Core/Annotations/Consensus.cpp) into oracle/_ref/ and run beside the oracle's restatement:
 *   QChar(int), unicode(), toLatin1() (0 beyond Latin-1);
 *   QString::length(), operator[] returning a QCharRef whose assignment GROWS the string (Qt 5 pads with
 *   spaces: Consensus::calculate fills an initially empty consensusSequence_ that way);
 *   QList<T>::at / size / append;  qDebug() << anything (discarded).
 * Nothing here is used by the product. */
#ifndef TSQ_REF_QT_MIN_H
#define TSQ_REF_QT_MIN_H
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
class QCharRef {
 public:
  QCharRef(QString& s, int i) : s_(s), i_(i) {}
  inline QCharRef& operator=(QChar c);
  inline unsigned short unicode() const;
 private:
  QString& s_;
  int i_;
};

class QString {
 public:
  QString() {}
  QString(const char* s) { while (s && *s) d_.push_back((unsigned short)(unsigned char)*s++); }
  int length() const { return (int)d_.size(); }
  int size() const { return (int)d_.size(); }
  QCharRef operator[](int i) { return QCharRef(*this, i); }
  const QChar operator[](int i) const { return QChar((int)d_[(size_t)i]); }
  void append(QChar c) { d_.push_back(c.unicode()); }
  std::vector<unsigned short> d_;
};

inline QCharRef& QCharRef::operator=(QChar c) {
  if (i_ >= (int)s_.d_.size()) s_.d_.resize((size_t)i_ + 1, (unsigned short)' ');
  s_.d_[(size_t)i_] = c.unicode();
  return *this;
}
inline unsigned short QCharRef::unicode() const { return i_ < (int)s_.d_.size() ? s_.d_[(size_t)i_] : 0; }

template <typename T>
class QList {
 public:
  const T& at(int i) const { return v_[(size_t)i]; }
  int size() const { return (int)v_.size(); }
  void append(const T& t) { v_.push_back(t); }
 private:
  std::vector<T> v_;
};

struct QDebugSink {
  template <typename T> QDebugSink& operator<<(const T&) { return *this; }
};
inline QDebugSink qDebug() { return QDebugSink(); }
#endif
