// Error state, version and launch accounting of libpp_b200.
#include <stdarg.h>

#include <atomic>

#include "pp_common.cuh"

namespace pp {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace pp

extern "C" int pp_version(void) { return 100; }
extern "C" const char *pp_last_error(void) { return pp::g_err; }
extern "C" int64_t pp_launch_count(void) { return pp::g_launches.load(std::memory_order_relaxed); }
