// Register-only pipe-rate microbenchmarks for B200 (sm_100a): FFMA, FFMA2 (packed fp32x2), MUFU.EX2, MUFU.RCP,
// DFMA.  Prints lane-ops per clock per SM; used for the roofline denominators in DESIGN.md / bench.py.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(1024) k(float *out, float seed, long long *cycles) {
    float a[ILP], b = seed, c = seed * 0.5f;
    float2 p[ILP];
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; p[i] = make_float2(a[i], a[i] + 1.f); d[i] = a[i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) a[i] = fmaf(a[i], b, c);
            if (OP == 1) p[i] = __ffma2_rn(p[i], make_float2(b, b), make_float2(c, c));
            if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 4) d[i] = fma(d[i], (double)b, (double)c);
            if (OP == 5) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); p[i] = __ffma2_rn(p[i], make_float2(b, b), make_float2(c, c)); p[i] = __ffma2_rn(p[i], make_float2(c, b), make_float2(b, c)); }
            if (OP == 6) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], c, b); a[i] = fmaf(a[i], b, c); a[i] = fmaf(a[i], c, b);}
            if (OP == 7) p[i] = __fmul2_rn(p[i], make_float2(b, c));
            if (OP == 8) p[i] = __fadd2_rn(p[i], make_float2(b, c));
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i] + p[i].x + p[i].y + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, double ops_per_iter_lane) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 1024, blocks = sms * 2;
    float *out; long long *cyc;
    cudaMalloc(&out, sizeof(float) * threads * blocks);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    k<OP><<<blocks, threads>>>(out, 1.0001f, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 1.0001f, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024]; cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    // two resident blocks of 1024 threads per SM -> lanes per SM = 2048
    double lane_ops = (double)ITERS * ILP * ops_per_iter_lane * 2048.0;
    printf("%-28s %8.3f ms  %10.0f clk/block  -> %7.2f lane-ops/clk/SM  (%.2f Tops/s chip)\n", name, ms, avg,
           lane_ops / avg, (double)ITERS * ILP * ops_per_iter_lane * threads * blocks / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA (scalar fp32)", 1);
    run<1>("FFMA2 (packed, per fp32 op)", 2);
    run<7>("FMUL2 (packed, per fp32 op)", 2);
    run<8>("FADD2 (packed, per fp32 op)", 2);
    run<2>("MUFU.EX2", 1);
    run<3>("MUFU.RCP", 1);
    run<4>("DFMA", 1);
    run<5>("EX2 + 2 FFMA2 (per EX2)", 1);
    run<6>("EX2 + 4 FFMA  (per EX2)", 1);
    return 0;
}
