// Version / error plumbing of the C ABI (include/tstereo.h).
#include "common.cuh"
#include <atomic>
#include <cstring>

namespace tstereo {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace tstereo

extern "C" {
int tstereo_version(void) { return TSTEREO_VERSION; }
const char* tstereo_last_error(void) { return tstereo::g_err; }
long long tstereo_launch_count(void) { return tstereo::g_launches.load(std::memory_order_relaxed); }
}
