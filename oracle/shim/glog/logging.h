// Stand-in for <glog/logging.h> so the UNMODIFIED reference sources under
// /root/reference compile in an image without glog (test infrastructure only;
// used solely by the oracle/_ref build recipe in oracle/Makefile).
// The reference uses: CHECK, LOG, LOG_IF, LOG_EVERY_N, severities, four FLAGS_*
// and InitGoogleLogging (reference include/util/io.h:26-39, util/debug.h:27-38).
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

namespace google {
typedef int LogSeverity;
const int INFO = 0, WARNING = 1, ERROR = 2, FATAL = 3;
inline void InitGoogleLogging(const char *) {}
}  // namespace google

static int FLAGS_minloglevel = 0;
static bool FLAGS_logtostderr = true;
static std::string FLAGS_log_dir;
static bool FLAGS_log_prefix = false;

namespace gv_shim {
class Message {
public:
    Message(int severity, bool enabled = true) : severity_(severity), enabled_(enabled) {}
    ~Message() {
        if (enabled_ && severity_ >= FLAGS_minloglevel)
            std::cerr << stream_.str() << std::endl;
        if (enabled_ && severity_ == google::FATAL)
            std::abort();
    }
    Message &ref() { return *this; }
    template<class T>
    Message &operator<<(const T &value) {
        if (enabled_)
            stream_ << value;
        return *this;
    }
    Message &operator<<(std::ostream &(*manip)(std::ostream &)) {
        if (enabled_)
            stream_ << manip;
        return *this;
    }
private:
    int severity_;
    bool enabled_;
    std::ostringstream stream_;
};
}  // namespace gv_shim

#define LOG(severity) gv_shim::Message(google::severity).ref()
#define LOG_IF(severity, condition) gv_shim::Message(google::severity, (condition)).ref()
#define LOG_EVERY_N(severity, n) gv_shim::Message(google::severity, false).ref()
#define CHECK(condition) gv_shim::Message(google::FATAL, !(condition)).ref()
