// qtstub_core.h -- TEST INFRASTRUCTURE: the few Qt5 declarations host/qt/B200GotohTool.{h,cpp} and the
// reference's Core/AlignmentTool.h + Core/XMLHelper.h touch, so that tests/test_qt_adapter_syntax.py can
// type-check the Qt adapter against the REAL reference headers with `g++ -fsyntax-only` in an image that
// has no Qt.  Declarations only (nothing is linked or run); signatures follow the Qt 5 documentation.
#ifndef TSQ_QTSTUB_CORE_H
#define TSQ_QTSTUB_CORE_H
#include <cstddef>
#include <string>
#include <vector>

#define Q_OBJECT
#define signals public
#define slots
#define emit

class QByteArray {
 public:
  const char *constData() const;
  int size() const;
};

class QString {
 public:
  QString();
  QString(const char *);
  QString(const QString &);
  QString &operator=(const QString &);
  bool operator==(const QString &) const;
  bool operator!=(const QString &) const;
  bool operator==(const char *) const;
  bool operator!=(const char *) const;
  int toInt(bool *ok = nullptr, int base = 10) const;
  QByteArray toLocal8Bit() const;
  QByteArray toLatin1() const;
  static QString number(int, int base = 10);
  static QString fromUtf8(const char *, int size = -1);
  friend const QString operator+(const QString &, const QString &);
};
const QString operator+(const QString &, const QString &);

template <typename T>
class QList {
 public:
  QList<T> &operator<<(const T &);
  int size() const;
  const T &at(int) const;
};

class QStringList : public QList<QString> {
 public:
  QStringList &operator<<(const QString &);
  QStringList &operator<<(const char *);
};

namespace Qt {
enum ConnectionType { AutoConnection, DirectConnection, QueuedConnection };
}

class QGenericArgument {
 public:
  QGenericArgument(const char *name = nullptr, const void *data = nullptr);
};
#define Q_ARG(type, data) QGenericArgument(#type, static_cast<const void *>(&static_cast<const type &>(data)))

class QObject {
 public:
  explicit QObject(QObject *parent = nullptr);
  virtual ~QObject();
  QObject *parent() const;
};

class QMetaObject {
 public:
  static bool invokeMethod(QObject *obj, const char *member, Qt::ConnectionType type,
                           QGenericArgument val0 = QGenericArgument(), QGenericArgument val1 = QGenericArgument());
};

class QThread : public QObject {
 public:
  explicit QThread(QObject *parent = nullptr);
  void start();
  bool wait(unsigned long msecs = ~0ul);

 protected:
  virtual void run();
};

class QDomNode {
 public:
  bool isNull() const;
  class QDomElement firstChildElement(const QString &tagName = QString()) const;
  class QDomElement nextSiblingElement(const QString &tagName = QString()) const;
  QDomNode appendChild(const QDomNode &newChild);
};

class QDomElement : public QDomNode {
 public:
  QString tagName() const;
  QString text() const;
};

class QDomNodeList {
 public:
  int count() const;
  QDomNode item(int index) const;
};

class QDomDocument : public QDomNode {
 public:
  QDomElement createElement(const QString &tagName);
  QDomNodeList elementsByTagName(const QString &tagname) const;
};
#endif
