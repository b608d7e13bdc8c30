// pb_prof.h - launch counter and optional per-kernel CUDA-event profiler.
//
// Every kernel launch in the library goes through a PbProfScope: it always bumps the
// launch counter (reported as gpu_launches by bench.py) and, when profiling is enabled
// (patolette_b200_profile_enable), brackets the launch with CUDA events on the launching
// stream so bench.py can report per-kernel durations and algorithmic bytes.
#pragma once
#include <cuda_runtime.h>

#include <string>

void pb_prof_enable(bool on);
bool pb_prof_enabled();
long pb_prof_launch_count();
// Algorithmic bytes to attribute to the NEXT recorded launch (set by the pipeline).
void pb_prof_next_bytes(double bytes);
void pb_prof_reset();
// Resolve all pending events (synchronises them) and return a JSON object
// {"kernel": {"launches": n, "ms": total, "bytes": total}, ...}.
std::string pb_prof_json();
// Timeline mode: the split loop keeps its two streams while profiling (per-kernel durations then include waiting
// for the other stream's CTAs) and pb_prof_timeline_text() lists "name stream start_ms end_ms" of every launch.
void pb_prof_set_timeline(bool on);
bool pb_prof_timeline();
std::string pb_prof_timeline_text();

struct PbProfScope {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    const char *name;
    double bytes = 0;
    PbProfScope(const char *kernel_name, cudaStream_t stream, bool takes_bytes = true);
    ~PbProfScope();
};
