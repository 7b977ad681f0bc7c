// Error state, version, launch accounting and the opt-in per-kernel event profiler of libpp_b200.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pp_common.cuh"

namespace pp {

static thread_local char g_err[512] = "";
static thread_local cudaStream_t g_stream = nullptr;
static std::atomic<int64_t> g_launches{0};

struct Mark {
    const char *name;
    cudaEvent_t ev;
};
static std::mutex g_prof_mu;
static bool g_prof = false;
static std::vector<Mark> g_marks;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void prof_mark(const char *name)
{
    if (!g_prof) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, g_stream);
    g_marks.push_back({name, e});
}

void enter(cudaStream_t st)
{
    g_stream = st;
    prof_mark("@enter");
}

}  // namespace pp

extern "C" int pp_version(void) { return 100; }
extern "C" const char *pp_last_error(void) { return pp::g_err; }
extern "C" int64_t pp_launch_count(void) { return pp::g_launches.load(std::memory_order_relaxed); }

extern "C" int pp_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(pp::g_prof_mu);
    if (on) {
        for (auto &m : pp::g_marks) cudaEventDestroy(m.ev);
        pp::g_marks.clear();
    }
    pp::g_prof = on != 0;
    return PP_OK;
}

extern "C" int pp_profile_report(char *buf, size_t buf_bytes)
{
    std::lock_guard<std::mutex> lk(pp::g_prof_mu);
    if (!buf || buf_bytes == 0) return PP_ERR_INVALID;
    buf[0] = 0;
    if (pp::g_marks.empty()) return PP_OK;
    if (cudaEventSynchronize(pp::g_marks.back().ev) != cudaSuccess) return PP_ERR_CUDA;
    std::map<std::string, std::pair<int64_t, double>> acc;
    std::vector<std::string> order;
    for (size_t i = 1; i < pp::g_marks.size(); ++i) {
        const char *name = pp::g_marks[i].name;
        if (name[0] == '@') continue;          // time between calls belongs to nobody
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, pp::g_marks[i - 1].ev, pp::g_marks[i].ev) != cudaSuccess) continue;
        auto it = acc.find(name);
        if (it == acc.end()) {
            order.push_back(name);
            acc[name] = {1, (double)ms};
        } else {
            it->second.first += 1;
            it->second.second += ms;
        }
    }
    size_t off = 0;
    for (auto &n : order) {
        int w = snprintf(buf + off, buf_bytes - off, "%s %lld %.6f\n", n.c_str(), (long long)acc[n].first, acc[n].second);
        if (w < 0 || (size_t)w >= buf_bytes - off) break;
        off += (size_t)w;
    }
    return PP_OK;
}
