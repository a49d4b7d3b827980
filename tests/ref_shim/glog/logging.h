// stand-in for glog: LOG(severity) << ... is swallowed (the reference logs, it never branches on it)
#ifndef MINI_GLOG_H
#define MINI_GLOG_H
#include <ostream>
namespace mini_glog {
struct Sink {
    template <class T> Sink& operator<<(const T&) { return *this; }
    Sink& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}
#define LOG(severity) ::mini_glog::Sink()
#endif
