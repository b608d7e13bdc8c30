// Microbenchmark: dependent-chain latency of DADD / DFMA / IADD64 / LDS+DADD on this GPU (cycles per op).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -o gpurun_out/dadd_latency tools/ubench/dadd_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dadd(const double *in, double *out, long long *cyc, int n, int mode) {
    __shared__ double sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    double s = in[0];
    long long acc = (long long)in[1];
    const long long t0 = clock64();
    if (mode == 0) {
#pragma unroll 16
        for (int i = 0; i < n; i++) s = __dadd_rn(s, 1.0000001);
    } else if (mode == 1) {
#pragma unroll 16
        for (int i = 0; i < n; i++) s = __fma_rn(s, 1.0000001, 0.5);
    } else if (mode == 2) {
#pragma unroll 8
        for (int i = 0; i < n; i++) s = __dadd_rn(s, sm[i & 4095]);
    } else if (mode == 3) {
#pragma unroll 16
        for (int i = 0; i < n; i++) acc = acc * 3 + (acc >> 7);
    } else if (mode == 4) {
        float f = (float)s;
#pragma unroll 16
        for (int i = 0; i < n; i++) f = __fadd_rn(f, 1.0000001f);
        s = f;
    } else if (mode == 5) { // warp shuffle 64-bit dependent
#pragma unroll 16
        for (int i = 0; i < n; i++) s = __shfl_xor_sync(0xffffffffu, s, 1);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)acc;
}

int main() {
    double *in, *out;
    long long *cyc;
    cudaMalloc(&in, 4096 * 8);
    cudaMalloc(&out, 1 << 22);
    cudaMalloc(&cyc, 8);
    double h[4096];
    for (int i = 0; i < 4096; i++) h[i] = 1.0 + i * 1e-9;
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    const char *names[] = {"DADD chain", "DFMA chain", "LDS+DADD chain (unroll 8)", "IMAD64+shift chain", "FADD chain", "SHFL.64 chain"};
    const int n = 1 << 16;
    for (int mode = 0; mode < 6; mode++)
        for (int cfg = 0; cfg < 3; cfg++) {
            const int blocks = cfg == 0 ? 1 : 148 * 4, threads = cfg == 2 ? 224 : 32;
            k_dadd<<<blocks, threads>>>(in, out, cyc, n, mode);
            k_dadd<<<blocks, threads>>>(in, out, cyc, n, mode);
            long long c;
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("%-28s blocks %4d x %3d threads: %.2f cycles/op\n", names[mode], blocks, threads, (double)c / n);
        }
    return 0;
}
