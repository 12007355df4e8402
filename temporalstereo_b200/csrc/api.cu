// Version / error plumbing of the C ABI (include/tstereo.h).
#include "common.cuh"
#include <atomic>
#include <cstdlib>
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
bool pdl_enabled() {           // read at every launch: tests and benches flip TSTEREO_PDL inside one process
    const char* e = getenv("TSTEREO_PDL");
    return !(e && *e == '0');
}
}  // namespace tstereo

extern "C" {
int tstereo_version(void) { return TSTEREO_VERSION; }
const char* tstereo_last_error(void) { return tstereo::g_err; }
#ifndef TSTEREO_BUILD_ID
#define TSTEREO_BUILD_ID "unknown"
#endif
// the "TSTEREO_BUILD_ID=" prefix lets the loader read the id from the file without dlopen()ing a stale library
static const char g_build_id[] = "TSTEREO_BUILD_ID=" TSTEREO_BUILD_ID;
const char* tstereo_build_id(void) { return g_build_id + 17; }
long long tstereo_launch_count(void) { return tstereo::g_launches.load(std::memory_order_relaxed); }
}
