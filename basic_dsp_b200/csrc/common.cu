#include <stdarg.h>
#include <atomic>
#include <string.h>

#include "common.cuh"

namespace bdsp {
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_last_error() { return g_err; }
static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(); }
}  // namespace bdsp
