#include "prof.cuh"

#include <mutex>
#include <vector>

#include "common.cuh"

namespace ucod {

namespace {
struct Rec {
    int cls;
    cudaEvent_t a, b;
};
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
double g_work[KC_COUNT] = {0};
long long g_launches[KC_COUNT] = {0};
long long g_total = 0;
cudaEvent_t g_pending_start = nullptr;

cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_on = on != 0;
}

void prof_pre(int cls, cudaStream_t s, double work) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (cls < 0 || cls >= KC_COUNT) cls = KC_OTHER;
    ++g_total;
    ++g_launches[cls];
    g_work[cls] += work;
    if (g_on) {
        g_pending_start = get_event();
        cudaEventRecord(g_pending_start, s);
    }
}

void prof_post(int cls, cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_on && g_pending_start != nullptr) {
        cudaEvent_t b = get_event();
        cudaEventRecord(b, s);
        g_recs.push_back(Rec{cls, g_pending_start, b});
        g_pending_start = nullptr;
    }
}

int prof_collect(double* ms, double* work, long long* launches) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < KC_COUNT; ++i) {
        if (ms) ms[i] = 0.0;
        if (work) work[i] = g_work[i];
        if (launches) launches[i] = g_launches[i];
        g_work[i] = 0.0;
        g_launches[i] = 0;
    }
    for (const Rec& r : g_recs) {
        cudaError_t e = cudaEventSynchronize(r.b);
        if (e != cudaSuccess) {
            set_last_error("prof_collect: %s", cudaGetErrorString(e));
            return 2;
        }
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        if (ms) ms[r.cls] += t;
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    g_recs.clear();
    return 0;
}

long long launch_count_total() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_total;
}

}  // namespace ucod
