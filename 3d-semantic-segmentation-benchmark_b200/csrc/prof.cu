// prof.cu -- optional per-kernel timing inside libpcnbr (bench.py's roofline and kernel sweep).
//
// When enabled, every kernel launch site (PCNBR_TIMED in common.cuh) is bracketed by two CUDA events
// recorded on the stream the kernel is launched on, together with the ALGORITHMIC work of that launch
// (bytes that must cross HBM and flops that must be executed) as stated by the launch site.
// pcnbr_prof_collect() synchronises the events and returns one text line per launch.  Disabled (the
// default) the scope is a single predictable branch; it must stay disabled under CUDA-graph capture.
#include "common.cuh"
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace pcnbr {

struct ProfRecord {
    const char* name;
    cudaEvent_t e0, e1;
    double bytes, flops;
};

static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfRecord> g_prof_records;
static std::vector<cudaEvent_t> g_prof_pool;

bool prof_enabled() { return g_prof_on; }

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(const char* name, cudaStream_t s, double bytes, double flops) : slot_(-1), stream_(s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRecord r{name, prof_event(), prof_event(), bytes, flops};
    cudaEventRecord(r.e0, s);
    slot_ = (int)g_prof_records.size();
    g_prof_records.push_back(r);
}

ProfScope::~ProfScope() {
    if (slot_ < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot_ < (int)g_prof_records.size()) cudaEventRecord(g_prof_records[slot_].e1, stream_);
}

}  // namespace pcnbr

using namespace pcnbr;

extern "C" void pcnbr_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = on != 0;
}

// Writes "kernel\tms\talgorithmic_bytes\talgorithmic_flops\n" per recorded launch (in launch order) into out
// (NUL-terminated, truncated at whole lines), clears the records and returns the number of launches recorded.
extern "C" int pcnbr_prof_collect(char* out, size_t cap) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::string text;
    char line[256];
    int n = 0;
    for (ProfRecord& r : g_prof_records) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            snprintf(line, sizeof line, "%s\t%.6f\t%.0f\t%.0f\n", r.name, ms, r.bytes, r.flops);
            if (out && text.size() + strlen(line) + 1 <= cap) text += line;
            ++n;
        }
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_records.clear();
    if (out && cap) memcpy(out, text.c_str(), text.size() + 1);
    return n;
}
