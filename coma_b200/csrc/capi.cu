// Library-level entry points: version, thread-local error string, launch counter.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace coma {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static thread_local const char *g_last_kernel = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void note_kernel(const char *name) { g_last_kernel = name; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace coma

extern "C" {
int coma_b200_version(void) { return 100; }
const char *coma_b200_last_error(void) { return coma::g_err; }
int64_t coma_b200_launch_count(void) { return coma::g_launches.load(std::memory_order_relaxed); }
const char *coma_b200_last_kernel(void) { return coma::g_last_kernel; }
}
