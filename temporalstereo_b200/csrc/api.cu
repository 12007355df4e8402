// Version / error plumbing of the C ABI (include/tstereo.h).
#include "common.cuh"
#include <cstring>

namespace tstereo {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace tstereo

extern "C" {
int tstereo_version(void) { return TSTEREO_VERSION; }
const char* tstereo_last_error(void) { return tstereo::g_err; }
}
