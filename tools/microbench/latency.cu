// Latencies (SM clocks) of the building blocks of the serial proposal phase of a small ensemble's half-step, one warp, one thread active:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency latency.cu && ./latency
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lightcurve_fitting_b200/csrc/lcf_device.cuh"
using namespace lcf;

__global__ void k_lat(double *buf, long long *out, double x0, double y0, unsigned int *bar) {
    double x = x0, acc = 0.;
    long long t0, t1;
    const int tid = threadIdx.x;
    // dependent DFMA chain
    if (tid == 0) { t0 = clock64(); for (int i = 0; i < 64; ++i) x = fma(x, 1.0000001, 1e-9); t1 = clock64(); out[0] = (t1 - t0) / 64; acc += x; }
    // pow_fast, libm pow, libm log, log2_fast, exp2_core
    if (tid == 0) { x = x0; t0 = clock64(); for (int i = 0; i < 8; ++i) x = pow_fast(x, y0) + 1.5; t1 = clock64(); out[1] = (t1 - t0) / 8; acc += x; }
    if (tid == 0) { x = x0; t0 = clock64(); for (int i = 0; i < 8; ++i) x = pow(x, y0) + 1.5; t1 = clock64(); out[2] = (t1 - t0) / 8; acc += x; }
    if (tid == 0) { x = x0; t0 = clock64(); for (int i = 0; i < 8; ++i) x = log(x) + 3.; t1 = clock64(); out[3] = (t1 - t0) / 8; acc += x; }
    if (tid == 0) { x = x0; t0 = clock64(); for (int i = 0; i < 8; ++i) x = log2_fast(x) + 3.; t1 = clock64(); out[4] = (t1 - t0) / 8; acc += x; }
    if (tid == 0) { x = 0.3; t0 = clock64(); for (int i = 0; i < 8; ++i) x = exp2_core(x) - 0.9; t1 = clock64(); out[5] = (t1 - t0) / 8; acc += x; }
    if (tid == 0) { x = x0; t0 = clock64(); for (int i = 0; i < 8; ++i) x = 1. / x + 0.7; t1 = clock64(); out[6] = (t1 - t0) / 8; acc += x; }
    if (tid == 0) { x = x0; t0 = clock64(); for (int i = 0; i < 8; ++i) x = sqrt(x) + 0.7; t1 = clock64(); out[7] = (t1 - t0) / 8; acc += x; }
    // philox
    if (tid == 0) { uint32_t r[4] = {1, 2, 3, 4}; t0 = clock64(); for (int i = 0; i < 8; ++i) philox4x32_10(r[0], r[1], r[2], r[3], 5u, 6u, r); t1 = clock64(); out[8] = (t1 - t0) / 8; acc += r[0]; }
    // dependent global loads (L2 hits: the buffer was written by another launch), pointer chasing through indices stored as doubles
    if (tid == 0) { long long j = 0; t0 = clock64(); for (int i = 0; i < 16; ++i) j = (long long)__ldcg(buf + j); t1 = clock64(); out[9] = (t1 - t0) / 16; acc += (double)j; }
    // CTA barrier round trip
    t0 = clock64(); for (int i = 0; i < 16; ++i) __syncthreads(); t1 = clock64(); if (tid == 0) out[10] = (t1 - t0) / 16;
    // global atomic with return + fence (the grid barrier's arrival)
    if (tid == 0) { t0 = clock64(); unsigned int v = 0; for (int i = 0; i < 8; ++i) { __threadfence(); v += atomicAdd(bar, 1u); } t1 = clock64(); out[11] = (t1 - t0) / 8; acc += v; }
    if (tid == 0) { t0 = clock64(); unsigned int v = 0; for (int i = 0; i < 8; ++i) v += ld_acquire_gpu(bar); t1 = clock64(); out[12] = (t1 - t0) / 8; acc += v; }
    if (acc == 12345.678) buf[0] = acc;
}

int main() {
    double *buf; long long *out; unsigned int *bar;
    cudaMalloc(&buf, 1 << 20); cudaMalloc(&out, 16 * 8); cudaMalloc(&bar, 8);
    cudaMemset(bar, 0, 8);
    double h[1 << 14];
    for (int i = 0; i < (1 << 14); ++i) h[i] = (double)((i * 977 + 131) % (1 << 14));   // a pointer chain with 8 KB+ strides
    cudaMemcpy(buf, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 2; ++rep) k_lat<<<1, 512>>>(buf, out, 1.7, 0.37, bar);
    long long o[16];
    cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost);
    const char *names[13] = {"dependent DFMA", "pow_fast", "libm pow", "libm log", "log2_fast", "exp2_core", "1/x (double)", "sqrt (double)", "philox4x32-10",
                             "dependent global load (L2)", "__syncthreads (512 threads)", "fence + atomicAdd(return)", "ld.acquire.gpu"};
    for (int i = 0; i < 13; ++i) printf("%-32s %6lld clk\n", names[i], o[i]);
    return 0;
}
