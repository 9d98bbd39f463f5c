// Header-only stand-in for the handful of Qt types the reference's hot-path
// sources name (containers, debug sink, file + data stream, QObject macros).
// TEST INFRASTRUCTURE ONLY: it exists so that the reference's own
// chessboard.cpp / chessai.cpp / dqn.cpp (and dqn.cu) compile unmodified from
// /root/reference into oracle/_ref/.  Nothing in the product links this.
#ifndef XQ_QTSHIM_H
#define XQ_QTSHIM_H
#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <random>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

typedef uint64_t quint64;
typedef int64_t qint64;
typedef int32_t qint32;
typedef uint32_t quint32;

template <class A, class B> using QPair = std::pair<A, B>;
template <class A, class B> inline std::pair<A, B> qMakePair(const A& a, const B& b) { return std::pair<A, B>(a, b); }

template <class T> class QVector : public std::vector<T> {
public:
    using std::vector<T>::vector;
    QVector() = default;
    QVector(std::initializer_list<T> il) : std::vector<T>(il) {}
    void append(const T& v) { this->push_back(v); }
    bool isEmpty() const { return this->empty(); }
    QVector<T>& fill(const T& v) { std::fill(this->begin(), this->end(), v); return *this; }
};

class QString {
    std::string s_;
public:
    QString() = default;
    QString(const char* c) : s_(c ? c : "") {}
    QString(const std::string& s) : s_(s) {}
    static QString fromStdString(const std::string& s) { return QString(s); }
    std::string toStdString() const { return s_; }
    // %N placeholder substitution, lowest-numbered marker first (Qt semantics)
    QString arg(const QString& a) const {
        int best = -1; size_t pos = std::string::npos, len = 0;
        for (size_t i = 0; i + 1 < s_.size(); ++i) {
            if (s_[i] == '%' && s_[i + 1] >= '0' && s_[i + 1] <= '9') {
                size_t j = i + 1; int n = 0;
                while (j < s_.size() && j < i + 3 && s_[j] >= '0' && s_[j] <= '9') { n = n * 10 + (s_[j] - '0'); ++j; }
                if (best < 0 || n < best) { best = n; pos = i; len = j - i; }
            }
        }
        if (best < 0) return *this;
        std::string marker = s_.substr(pos, len), out = s_;
        for (size_t p = out.find(marker); p != std::string::npos; p = out.find(marker, p + a.s_.size()))
            out.replace(p, marker.size(), a.s_);
        return QString(out);
    }
    QString arg(int v) const { return arg(QString(std::to_string(v))); }
    QString arg(long long v) const { return arg(QString(std::to_string(v))); }
    const std::string& str() const { return s_; }
};

// qDebug()/qWarning(): collected into a per-thread line buffer, printed only
// when XQ_REF_VERBOSE is set (the reference prints one line per game).
class QDebug {      // no iostreams: the library carries a private libstdc++ whose locale facets are not initialised under dlopen
    std::string os_;
    bool first_ = true;
    void sep() { if (!first_) os_ += ' '; first_ = false; }
    template <class T> QDebug& num(const char* fmt, T v) { char b[48]; std::snprintf(b, sizeof b, fmt, v); sep(); os_ += b; return *this; }
public:
    QDebug() = default;
    QDebug(const QDebug&) {}
    ~QDebug() { if (std::getenv("XQ_REF_VERBOSE")) std::fprintf(stderr, "%s\n", os_.c_str()); }
    QDebug& operator<<(const char* s) { sep(); os_ += (s ? s : ""); return *this; }
    QDebug& operator<<(const std::string& s) { sep(); os_ += s; return *this; }
    QDebug& operator<<(const QString& s) { sep(); os_ += '"'; os_ += s.str(); os_ += '"'; return *this; }
    QDebug& operator<<(int v) { return num("%d", v); }
    QDebug& operator<<(unsigned v) { return num("%u", v); }
    QDebug& operator<<(long v) { return num("%ld", v); }
    QDebug& operator<<(unsigned long v) { return num("%lu", v); }
    QDebug& operator<<(long long v) { return num("%lld", v); }
    QDebug& operator<<(double v) { return num("%g", v); }
    QDebug& operator<<(bool v) { sep(); os_ += (v ? "true" : "false"); return *this; }
};
inline QDebug qDebug() { return QDebug(); }
inline QDebug qWarning() { return QDebug(); }

class QIODevice {
public:
    enum OpenModeFlag { NotOpen = 0, ReadOnly = 1, WriteOnly = 2, ReadWrite = 3, Append = 4, Truncate = 8, Text = 16 };
};
inline QIODevice::OpenModeFlag operator|(QIODevice::OpenModeFlag a, QIODevice::OpenModeFlag b) {
    return static_cast<QIODevice::OpenModeFlag>(static_cast<int>(a) | static_cast<int>(b));
}

// Relative file names are placed under $XQ_REF_OUTDIR (default /tmp/xq_ref_out)
// so that the reference's game_log.txt / autosave files never land in the repo.
class QFile : public QIODevice {
    std::string name_;
    FILE* f_ = nullptr;
    std::string err_;
    static std::string resolve(const std::string& n) {
        if (!n.empty() && n[0] == '/') return n;
        const char* d = std::getenv("XQ_REF_OUTDIR");
        std::string dir = d ? d : "/tmp/xq_ref_out";
        std::string cmd = "mkdir -p '" + dir + "'";
        if (std::system(cmd.c_str()) != 0) {}
        return dir + "/" + n;
    }
public:
    QFile() = default;
    explicit QFile(const QString& n) : name_(n.toStdString()) {}
    ~QFile() { close(); }
    QFile(const QFile&) = delete;
    void setFileName(const QString& n) { name_ = n.toStdString(); }
    bool open(OpenModeFlag m) {
        const char* mode = (m & Append) ? "ab" : (m & WriteOnly) ? "wb" : "rb";
        f_ = std::fopen(resolve(name_).c_str(), mode);
        if (!f_) err_ = std::strerror(errno);
        return f_ != nullptr;
    }
    bool isOpen() const { return f_ != nullptr; }
    void close() { if (f_) { std::fclose(f_); f_ = nullptr; } }
    void flush() { if (f_) std::fflush(f_); }
    QString errorString() const { return QString(err_); }
    FILE* handle() { return f_; }
};

class QTextStream {
    QFile* f_;
public:
    explicit QTextStream(QFile* f) : f_(f) {}
    QTextStream& operator<<(const QString& s) { if (f_ && f_->handle()) std::fputs(s.str().c_str(), f_->handle()); return *this; }
    QTextStream& operator<<(const char* s) { if (f_ && f_->handle()) std::fputs(s, f_->handle()); return *this; }
};

// QDataStream: integers big-endian (Qt default byte order), raw blocks verbatim.
class QDataStream {
    QFile* f_;
    int status_ = 0;
public:
    enum Status { Ok = 0, ReadPastEnd = 1, ReadCorruptData = 2, WriteFailed = 3 };
    enum Version { Qt_6_6 = 21 };
    explicit QDataStream(QFile* f) : f_(f) {}
    void setVersion(int) {}
    Status status() const { return static_cast<Status>(status_); }
    int writeRawData(const char* p, size_t n) {
        if (!f_->handle() || std::fwrite(p, 1, n, f_->handle()) != n) { status_ = WriteFailed; return -1; }
        return static_cast<int>(n);
    }
    int readRawData(char* p, size_t n) {
        if (!f_->handle() || std::fread(p, 1, n, f_->handle()) != n) { status_ = ReadPastEnd; return -1; }
        return static_cast<int>(n);
    }
    template <class I> void put_be(I v) { unsigned char b[sizeof(I)]; for (size_t i = 0; i < sizeof(I); ++i) b[i] = static_cast<unsigned char>(static_cast<uint64_t>(v) >> (8 * (sizeof(I) - 1 - i))); writeRawData(reinterpret_cast<char*>(b), sizeof(I)); }
    template <class I> void get_be(I& v) { unsigned char b[sizeof(I)] = {0}; readRawData(reinterpret_cast<char*>(b), sizeof(I)); uint64_t x = 0; for (size_t i = 0; i < sizeof(I); ++i) x = (x << 8) | b[i]; v = static_cast<I>(x); }
    QDataStream& operator<<(quint64 v) { put_be<quint64>(v); return *this; }
    QDataStream& operator<<(qint32 v) { put_be<quint32>(static_cast<quint32>(v)); return *this; }
    QDataStream& operator>>(quint64& v) { get_be<quint64>(v); return *this; }
    QDataStream& operator>>(qint32& v) { quint32 u = 0; get_be<quint32>(u); v = static_cast<qint32>(u); return *this; }
};

class QRandomGenerator {
    std::mt19937 g_{12345u};
public:
    static QRandomGenerator* global() { static QRandomGenerator r; return &r; }
    int bounded(int hi) { return hi > 0 ? static_cast<int>(g_() % static_cast<unsigned>(hi)) : 0; }
};

class QVariant {};

class QObject {
public:
    QObject() = default;
    virtual ~QObject() = default;
    QObject(const QObject&) = delete;
    QObject& operator=(const QObject&) = delete;
};
#define Q_OBJECT
#define slots
#define signals public
#define emit

#endif
