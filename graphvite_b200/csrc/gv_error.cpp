// Error slot and version string of libgv_b200 (see include/gv_b200.h, "Error model").
#include <string>

#include "gv_common.h"

namespace gv {

static thread_local std::string g_last_error;

void set_error(const std::string &message) {
    g_last_error = message;
}

int fail(const std::string &message) {
    set_error(message);
    return -1;
}

}  // namespace gv

extern "C" {

const char *gv_last_error(void) {
    return gv::g_last_error.c_str();
}

const char *gv_version(void) {
#ifdef GV_EMULATE_BUILD  // tests/emu/Makefile: the host build on top of the CUDA emulation, never shipped
    return GV_VERSION " (CUDA emulation -- test build, not a product)";
#else
    return GV_VERSION;
#endif
}

}  // extern "C"
