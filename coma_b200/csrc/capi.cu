// Library-level entry points: version, thread-local error string, launch counter.
#include <algorithm>
#include <atomic>
#include <stdarg.h>
#include <thread>
#include <vector>

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
// Host-side staging helper (no device work): dst[i, r, c] = (float)(src[i][(row0 + r) * 3 + c] - (sub ? sub[i][c] : 0.0)) for n
// samples of `rows` rows — the fp64 -> fp32 rounding of utils/misc.py:47-54 (and the fp64 subtraction of utils/coma_occupancy.py:287)
// written straight into the pinned staging buffer, one call per chunk instead of one numpy call per sample. `equal_to` (or NULL):
// 3 doubles every sub[i] is compared with bit for bit; *first_mismatch receives the first differing sample index or -1.
int coma_host_stage_rows_f64_f32(const double *const *src, const double *const *sub, int64_t n, int64_t row0, int64_t rows, float *dst,
                                 const double *equal_to, int64_t *first_mismatch) {
    if (!src || !dst || n < 0 || rows < 0 || row0 < 0) {
        coma::set_error("coma_host_stage_rows_f64_f32: bad arguments");
        return COMA_E_BADARG;
    }
    int64_t mism = -1;
    if (sub && equal_to)
        for (int64_t i = 0; i < n && mism < 0; ++i)
            if (!(sub[i][0] == equal_to[0] && sub[i][1] == equal_to[1] && sub[i][2] == equal_to[2])) mism = i;
    auto work = [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) {
            const double *s = src[i] + row0 * 3;
            float *d = dst + i * rows * 3;
            if (sub) {
                const double o0 = sub[i][0], o1 = sub[i][1], o2 = sub[i][2];
                for (int64_t r = 0; r < rows; ++r) {
                    d[3 * r + 0] = (float)(s[3 * r + 0] - o0);
                    d[3 * r + 1] = (float)(s[3 * r + 1] - o1);
                    d[3 * r + 2] = (float)(s[3 * r + 2] - o2);
                }
            } else {
                for (int64_t e = 0; e < rows * 3; ++e) d[e] = (float)s[e];
            }
        }
    };
    // samples are independent: a few host threads for big chunks (the first chunk of a job is not hidden behind any kernel)
    const int64_t bytes = n * rows * 3 * 8;
    int nt = (int)std::min<int64_t>(8, std::min<int64_t>(n, bytes >> 20));   // >= 1 MiB of input per thread
    const unsigned hw = std::thread::hardware_concurrency();
    if (hw > 0 && nt > (int)hw) nt = (int)hw;
    if (nt <= 1) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
        for (auto &x : th) x.join();
    }
    if (first_mismatch) *first_mismatch = mism;
    return 0;
}
// *first_mismatch = first i whose rows[i][0..count) differs bitwise-as-values from ref[0..count), or -1 (host helper, no device work).
int coma_host_rows_equal_f64(const double *const *rows, int64_t n, const double *ref, int64_t count, int64_t *first_mismatch) {
    if (!rows || !ref || !first_mismatch || n < 0 || count < 0) {
        coma::set_error("coma_host_rows_equal_f64: bad arguments");
        return COMA_E_BADARG;
    }
    *first_mismatch = -1;
    for (int64_t i = 0; i < n; ++i)
        for (int64_t c = 0; c < count; ++c)
            if (!(rows[i][c] == ref[c])) {
                *first_mismatch = i;
                return 0;
            }
    return 0;
}
int coma_b200_version(void) { return 100; }
const char *coma_b200_last_error(void) { return coma::g_err; }
int64_t coma_b200_launch_count(void) { return coma::g_launches.load(std::memory_order_relaxed); }
const char *coma_b200_last_kernel(void) { return coma::g_last_kernel; }
}
