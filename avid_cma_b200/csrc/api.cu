// Library-level entry points of include/avid_b200.h: version, error string, launch counter.
#include <stdarg.h>
#include "common.cuh"

namespace avid {

static thread_local char g_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

}  // namespace avid

extern "C" {

int avid_version(void) { return AVID_ABI_VERSION; }
const char* avid_last_error(void) { return avid::g_error; }
uint64_t avid_launch_count(void) { return avid::g_launches.load(); }
void avid_reset_launch_count(void) { avid::g_launches.store(0); }

}  // extern "C"
