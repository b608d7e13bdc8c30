// pb_hostpool.h - a few parked host threads for the O(512^2) host-side loops of the pipeline
// (the Wu dynamic programme of GQ, pb_pipeline.cu).  run(T, fn) calls fn(tid) for tid = 0..T-1, tid 0 on the
// calling thread, and returns when all are done.  Workers are created on first use and parked on a condition
// variable in between; they never touch CUDA.
#pragma once
#include <functional>

int pb_hostpool_default_threads();
void pb_hostpool_run(int threads, const std::function<void(int)> &fn);
