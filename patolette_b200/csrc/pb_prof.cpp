#include "pb_prof.h"

#include <atomic>
#include <map>
#include <mutex>
#include <sstream>
#include <vector>

namespace {
std::atomic<long> g_launches{0};
std::atomic<bool> g_on{false};
double g_next_bytes = 0;
struct Pending {
    const char *name;
    cudaEvent_t a, b;
    double bytes;
    cudaStream_t st;
};
std::atomic<bool> g_timeline{false};
std::mutex g_mu;
std::vector<Pending> g_pending;
struct Acc {
    long launches = 0;
    double ms = 0, bytes = 0;
    std::vector<float> each; // individual launch durations, in launch order
};
std::map<std::string, Acc> g_acc;
} // namespace

void pb_prof_enable(bool on) { g_on = on; }
void pb_prof_set_timeline(bool on) { g_timeline = on; }
bool pb_prof_timeline() { return g_on && g_timeline; }
bool pb_prof_enabled() { return g_on; }
long pb_prof_launch_count() { return g_launches; }
void pb_prof_next_bytes(double bytes) { g_next_bytes = bytes; }

void pb_prof_reset() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &p : g_pending) {
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_pending.clear();
    g_acc.clear();
}

PbProfScope::PbProfScope(const char *kernel_name, cudaStream_t stream, bool takes_bytes)
    : st(stream), name(kernel_name) {
    g_launches++;
    if (takes_bytes) {
        bytes = g_next_bytes;
        g_next_bytes = 0;
    }
    if (g_on) {
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, st);
    }
}

PbProfScope::~PbProfScope() {
    if (a) {
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_mu);
        g_pending.push_back(Pending{name, a, b, bytes, st});
    }
}

// "name stream start_ms end_ms" per recorded launch, times relative to the first one (events of different streams of
// one device are comparable).  Leaves the pending list intact: call before pb_prof_json().
std::string pb_prof_timeline_text() {
    std::lock_guard<std::mutex> lk(g_mu);
    std::ostringstream os;
    if (g_pending.empty()) return os.str();
    std::map<cudaStream_t, int> ids;
    for (auto &p : g_pending) cudaEventSynchronize(p.b);
    const cudaEvent_t base = g_pending.front().a;
    for (auto &p : g_pending) {
        float t0 = 0, t1 = 0;
        cudaEventElapsedTime(&t0, base, p.a);
        cudaEventElapsedTime(&t1, base, p.b);
        const int id = ids.emplace(p.st, (int)ids.size()).first->second;
        os << p.name << " " << id << " " << t0 << " " << t1 << "\n";
    }
    return os.str();
}

std::string pb_prof_json() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &p : g_pending) {
        cudaEventSynchronize(p.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, p.a, p.b);
        Acc &acc = g_acc[p.name];
        acc.launches++;
        acc.ms += ms;
        acc.bytes += p.bytes;
        acc.each.push_back(ms);
        cudaEventDestroy(p.a);
        cudaEventDestroy(p.b);
    }
    g_pending.clear();
    std::ostringstream os;
    os << "{";
    bool first = true;
    for (auto &kv : g_acc) {
        if (!first) os << ", ";
        first = false;
        os << "\"" << kv.first << "\": {\"launches\": " << kv.second.launches << ", \"ms\": " << kv.second.ms
           << ", \"bytes\": " << kv.second.bytes << ", \"each\": [";
        for (size_t i = 0; i < kv.second.each.size() && i < 64; i++) os << (i ? "," : "") << kv.second.each[i];
        os << "]}";
    }
    os << "}";
    return os.str();
}
