// Inner-loop formulations of the Planck x transmission sum, timed in isolation on B200 (sm_100a):
// 296 CTAs x 512 threads (the shipped kernel's occupancy), every warp sweeps a 24-record (48-sample) bank in shared
// memory REP times for two blackbodies per lane.  Prints Planck samples per clock per SM for each formulation, and
// the raw MUFU.EX2 / MUFU.RCP / MUFU.LG2 rates with the same harness.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o loops loops.cu && ./loops
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lightcurve_fitting_b200/csrc/lcf_device.cuh"

using namespace lcf;

#define K2 24
#define REP 400

__device__ __forceinline__ float2 ex2m1_pair_(float2 x) { return __fadd2_rn(make_float2(Mth<float>::ex2(x.x), Mth<float>::ex2(x.y)), make_float2(-1.f, -1.f)); }
__device__ __forceinline__ float ex2f_(float x) { return Mth<float>::ex2(x); }
__device__ __forceinline__ float rcpf_(float x) { return Mth<float>::rcp(x); }

// reciprocal of a positive normal float on the FMA/ALU pipes: bit-trick seed (12 % error) + three Newton steps (4e-8)
__device__ __forceinline__ float rcp_nr(float x) {
    float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
    float e = fmaf(-x, r, 1.f); r = fmaf(r, e, r);
    e = fmaf(-x, r, 1.f); r = fmaf(r, e, r);
    e = fmaf(-x, r, 1.f); r = fmaf(r, e, r);
    return r;
}
// seed + one cubic + one quadratic step (5 FMA-pipe ops, error ~3e-6 before the last step -> 1e-11)
__device__ __forceinline__ float rcp_nr2(float x) {
    float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
    float e = fmaf(-x, r, 1.f);
    float t = fmaf(e, e, e);
    r = fmaf(r, t, r);                                   // error e^3 = 1.7e-3
    e = fmaf(-x, r, 1.f);
    t = fmaf(e, e, e);
    r = fmaf(r, t, r);                                   // 5e-9
    return r;
}

// quad + Newton reciprocal
template <int NR>
__device__ __forceinline__ void quad_nr(const float4 *__restrict__ b4, int k2, float iA, float iB, float &SA, float &SB) {
    const float2 iA2 = make_float2(iA, iA), iB2 = make_float2(iB, iB);
    float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int k = 0; k < k2; ++k) {
        const float4 s = b4[k];
        const float2 a = make_float2(s.x, s.y), w = make_float2(s.z, s.w);
        const float2 dA = ex2m1_pair_(__fmul2_rn(a, iA2));
        const float2 dB = ex2m1_pair_(__fmul2_rn(a, iB2));
        const float2 p = __fmul2_rn(dA, dB);
        const float r = NR == 1 ? rcp_nr(p.x * p.y) : rcp_nr2(p.x * p.y);
        const float2 t = __fmul2_rn(w, make_float2(r * p.y, r * p.x));
        accA = __ffma2_rn(t, dB, accA);
        accB = __ffma2_rn(t, dA, accB);
    }
    SA = accA.x + accA.y;
    SB = accB.x + accB.y;
}

// quad + Newton, unrolled 4 records deep
__device__ __forceinline__ void quad_nr_u4(const float4 *__restrict__ b4, int k2, float iA, float iB, float &SA, float &SB) {
    const float2 iA2 = make_float2(iA, iA), iB2 = make_float2(iB, iB);
    float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < k2; ++k) {
        const float4 s = b4[k];
        const float2 a = make_float2(s.x, s.y), w = make_float2(s.z, s.w);
        const float2 dA = ex2m1_pair_(__fmul2_rn(a, iA2));
        const float2 dB = ex2m1_pair_(__fmul2_rn(a, iB2));
        const float2 p = __fmul2_rn(dA, dB);
        const float r = rcp_nr(p.x * p.y);
        const float2 t = __fmul2_rn(w, make_float2(r * p.y, r * p.x));
        accA = __ffma2_rn(t, dB, accA);
        accB = __ffma2_rn(t, dA, accB);
    }
    SA = accA.x + accA.y;
    SB = accB.x + accB.y;
}
// shipped quad, unrolled 4
__device__ __forceinline__ void quad_u4(const float4 *__restrict__ b4, int k2, float iA, float iB, float &SA, float &SB) {
    const float2 iA2 = make_float2(iA, iA), iB2 = make_float2(iB, iB);
    float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < k2; ++k) {
        const float4 s = b4[k];
        const float2 a = make_float2(s.x, s.y), w = make_float2(s.z, s.w);
        const float2 dA = ex2m1_pair_(__fmul2_rn(a, iA2));
        const float2 dB = ex2m1_pair_(__fmul2_rn(a, iB2));
        const float2 p = __fmul2_rn(dA, dB);
        const float r = rcpf_(p.x * p.y);
        const float2 t = __fmul2_rn(w, make_float2(r * p.y, r * p.x));
        accA = __ffma2_rn(t, dB, accA);
        accB = __ffma2_rn(t, dA, accB);
    }
    SA = accA.x + accA.y;
    SB = accB.x + accB.y;
}

// packed Newton reciprocal of two positive normal floats (FFMA2: one issue slot per step for both)
__device__ __forceinline__ float2 rcp_nr_x2(float2 x) {
    float2 r = make_float2(__int_as_float(0x7EF311C7 - __float_as_int(x.x)), __int_as_float(0x7EF311C7 - __float_as_int(x.y)));
    const float2 one = make_float2(1.f, 1.f), nx = make_float2(-x.x, -x.y);
    float2 e = __ffma2_rn(nx, r, one); r = __ffma2_rn(r, e, r);
    e = __ffma2_rn(nx, r, one); r = __ffma2_rn(r, e, r);
    e = __ffma2_rn(nx, r, one); r = __ffma2_rn(r, e, r);
    return r;
}
// quad, two records per iteration, packed Newton
template <bool TAB>
__device__ __forceinline__ void quad_nr_x2(const float4 *__restrict__ b4, int k2, float iA, float iB, const float2 *__restrict__ tab, int ts, float &SA, float &SB) {
    const float2 iA2 = make_float2(iA, iA), iB2 = make_float2(iB, iB);
    float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
    int k = 0;
    for (; k + 1 < k2; k += 2) {
        float2 a0, w0, a1, w1;
        if (TAB) {
            a0 = *reinterpret_cast<const float2 *>(b4 + k); a1 = *reinterpret_cast<const float2 *>(b4 + k + 1);
            w0 = tab[0]; w1 = tab[ts]; tab += 2 * ts;
        } else {
            const float4 s0 = b4[k], s1 = b4[k + 1];
            a0 = make_float2(s0.x, s0.y); w0 = make_float2(s0.z, s0.w); a1 = make_float2(s1.x, s1.y); w1 = make_float2(s1.z, s1.w);
        }
        const float2 dA0 = ex2m1_pair_(__fmul2_rn(a0, iA2)), dB0 = ex2m1_pair_(__fmul2_rn(a0, iB2));
        const float2 dA1 = ex2m1_pair_(__fmul2_rn(a1, iA2)), dB1 = ex2m1_pair_(__fmul2_rn(a1, iB2));
        const float2 p0 = __fmul2_rn(dA0, dB0), p1 = __fmul2_rn(dA1, dB1);
        const float2 r = rcp_nr_x2(make_float2(p0.x * p0.y, p1.x * p1.y));
        const float2 t0 = __fmul2_rn(w0, __fmul2_rn(make_float2(r.x, r.x), make_float2(p0.y, p0.x)));
        const float2 t1 = __fmul2_rn(w1, __fmul2_rn(make_float2(r.y, r.y), make_float2(p1.y, p1.x)));
        accA = __ffma2_rn(t0, dB0, accA); accB = __ffma2_rn(t0, dA0, accB);
        accA = __ffma2_rn(t1, dB1, accA); accB = __ffma2_rn(t1, dA1, accB);
    }
    if (k < k2) {
        float2 a, w;
        if (TAB) { a = *reinterpret_cast<const float2 *>(b4 + k); w = tab[0]; }
        else { const float4 s = b4[k]; a = make_float2(s.x, s.y); w = make_float2(s.z, s.w); }
        const float2 dA = ex2m1_pair_(__fmul2_rn(a, iA2)), dB = ex2m1_pair_(__fmul2_rn(a, iB2));
        const float2 p = __fmul2_rn(dA, dB);
        const float r = rcp_nr(p.x * p.y);
        const float2 t = __fmul2_rn(w, make_float2(r * p.y, r * p.x));
        accA = __ffma2_rn(t, dB, accA); accB = __ffma2_rn(t, dA, accB);
    }
    SA = accA.x + accA.y;
    SB = accB.x + accB.y;
}
// SC4 quad (A, A*, B, B*) per sample, both samples of the record in the packed lanes, packed Newton
__device__ __forceinline__ void sc4_nr(const float4 *__restrict__ b4, int k2, float iA, float iB, float &SA, float &SAs, float &SB, float &SBs) {
    const float c = (float)(1. / 0.74);
    const float2 iA2 = make_float2(iA, iA), iB2 = make_float2(iB, iB), iAs2 = make_float2(iA * c, iA * c), iBs2 = make_float2(iB * c, iB * c);
    float2 a = make_float2(0.f, 0.f), as = a, bb = a, bs = a;
#pragma unroll 2
    for (int k = 0; k < k2; ++k) {
        const float4 s = b4[k];
        const float2 x = make_float2(s.x, s.y), w = make_float2(s.z, s.w);
        const float2 dA = ex2m1_pair_(__fmul2_rn(x, iA2)), dAs = ex2m1_pair_(__fmul2_rn(x, iAs2));
        const float2 dB = ex2m1_pair_(__fmul2_rn(x, iB2)), dBs = ex2m1_pair_(__fmul2_rn(x, iBs2));
        const float2 pA = __fmul2_rn(dA, dAs), pB = __fmul2_rn(dB, dBs);
        const float2 r = rcp_nr_x2(__fmul2_rn(pA, pB));
        const float2 tA = __fmul2_rn(w, __fmul2_rn(r, pB)), tB = __fmul2_rn(w, __fmul2_rn(r, pA));
        a = __ffma2_rn(tA, dAs, a);   as = __ffma2_rn(tA, dA, as);
        bb = __ffma2_rn(tB, dBs, bb); bs = __ffma2_rn(tB, dB, bs);
    }
    SA = a.x + a.y; SAs = as.x + as.y; SB = bb.x + bb.y; SBs = bs.x + bs.y;
}

// oct: two points x four samples (two pair records) share one reciprocal; e/(1-e) form (no overflow, e = 2^-x)
template <int NR>
__device__ __forceinline__ void oct_em(const float4 *__restrict__ b4, int k2, float iA, float iB, float &SA, float &SB) {
    const float2 nA = make_float2(-iA, -iA), nB = make_float2(-iB, -iB), one = make_float2(1.f, 1.f);
    float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
    for (int k = 0; k + 1 < k2; k += 2) {
        const float4 s0 = b4[k], s1 = b4[k + 1];
        const float2 a0 = make_float2(s0.x, s0.y), a1 = make_float2(s1.x, s1.y);
        const float2 xA0 = __fmul2_rn(a0, nA), xB0 = __fmul2_rn(a0, nB), xA1 = __fmul2_rn(a1, nA), xB1 = __fmul2_rn(a1, nB);
        const float2 eA0 = make_float2(ex2f_(xA0.x), ex2f_(xA0.y)), eB0 = make_float2(ex2f_(xB0.x), ex2f_(xB0.y));
        const float2 eA1 = make_float2(ex2f_(xA1.x), ex2f_(xA1.y)), eB1 = make_float2(ex2f_(xB1.x), ex2f_(xB1.y));
        const float2 mA0 = __fadd2_rn(one, make_float2(-eA0.x, -eA0.y)), mB0 = __fadd2_rn(one, make_float2(-eB0.x, -eB0.y));
        const float2 mA1 = __fadd2_rn(one, make_float2(-eA1.x, -eA1.y)), mB1 = __fadd2_rn(one, make_float2(-eB1.x, -eB1.y));
        const float2 p0 = __fmul2_rn(mA0, mB0), p1 = __fmul2_rn(mA1, mB1);      // per sample: (1-eA)(1-eB)
        const float2 pp = __fmul2_rn(p0, p1);                                    // (s0 s2, s1 s3) products of 4
        const float P = pp.x * pp.y;
        const float r = NR == 0 ? rcpf_(P) : (NR == 1 ? rcp_nr(P) : rcp_nr2(P));
        const float2 rq = make_float2(r * pp.y, r * pp.x);                       // 1/(pp.x), 1/(pp.y)
        const float2 r0 = __fmul2_rn(rq, p1), r1 = __fmul2_rn(rq, p0);           // 1/p0, 1/p1 (per sample)
        const float2 t0 = __fmul2_rn(make_float2(s0.z, s0.w), r0), t1 = __fmul2_rn(make_float2(s1.z, s1.w), r1);   // w/(mA mB)
        accA = __ffma2_rn(__fmul2_rn(t0, eA0), mB0, accA);                       // w eA/(1-eA)
        accB = __ffma2_rn(__fmul2_rn(t0, eB0), mA0, accB);
        accA = __ffma2_rn(__fmul2_rn(t1, eA1), mB1, accA);
        accB = __ffma2_rn(__fmul2_rn(t1, eB1), mA1, accB);
    }
    SA = accA.x + accA.y;
    SB = accB.x + accB.y;
}

template <int V, int T>
__global__ void __launch_bounds__(T, 1024 / T) kloop(float *out, const float4 *bank_g, float i0) {
    __shared__ float4 bank[K2];
    __shared__ float2 tab[K2 * 32];
    if (threadIdx.x < K2) bank[threadIdx.x] = bank_g[threadIdx.x];
    for (int i = threadIdx.x; i < K2 * 32; i += blockDim.x) tab[i] = make_float2(bank_g[i / 32].z, bank_g[i / 32].w);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float iA = i0 * (1.f + 0.003f * lane), iB = iA * 1.07f;
    float accA = 0.f, accB = 0.f;
    for (int r = 0; r < REP; ++r) {
        float SA = 0, SB = 0, SAs = 0, SBs = 0;
        if (V == 0) planck_quad_f32<false, false, 32>(bank, K2, iA, iB, nullptr, 32, SA, SB);
        if (V == 1) planck_quad_f32<true, false, 32>(bank, K2, iA, iB, tab + lane, 32, SA, SB);
        if (V == 2) quad_nr<1>(bank, K2, iA, iB, SA, SB);
        if (V == 3) quad_nr<2>(bank, K2, iA, iB, SA, SB);
        if (V == 4) oct_em<0>(bank, K2, iA, iB, SA, SB);
        if (V == 5) oct_em<1>(bank, K2, iA, iB, SA, SB);
        if (V == 6) oct_em<2>(bank, K2, iA, iB, SA, SB);
        if (V == 8) quad_nr_u4(bank, K2, iA, iB, SA, SB);
        if (V == 9) quad_u4(bank, K2, iA, iB, SA, SB);
        if (V == 10) quad_nr_x2<false>(bank, K2, iA, iB, nullptr, 0, SA, SB);
        if (V == 11) quad_nr_x2<true>(bank, K2, iA, iB, tab + lane, 32, SA, SB);
        if (V == 12) { sc4_nr(bank, K2, iA, iB, SA, SAs, SB, SBs); SA += SAs; SB += SBs; }
        if (V == 13) planck_quad_f32<true, true, 32>(bank, K2, iA, iB, tab + lane, 32, SA, SB);
        if (V == 7) { planck_quad_sc4_f32<false>(bank, K2, iA, iB, SA, SAs, SB, SBs); SA += SAs; SB += SBs; }
        accA += SA; accB += SB;
        iA += 1e-6f; iB += 1e-6f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = accA + accB;
}

// raw MUFU rates, same occupancy: OP 0 ex2, 1 rcp, 2 lg2, 3 four ex2 + one rcp, 4 rsqrt
template <int OP>
__global__ void __launch_bounds__(512, 2) kmufu(float *out, float seed) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + 0.01f * i + 1e-5f * threadIdx.x;
    for (int it = 0; it < 2048; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 2) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) { if (i % 5 == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i])); else asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
            if (OP == 4) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return best;
}

int main() {
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    const int blocks = sms * 2, threads = 512;
    float *out; float4 *bank;
    cudaMalloc(&out, sizeof(float) * blocks * threads);
    float4 h[K2];
    for (int k = 0; k < K2; ++k) h[k] = make_float4(3.f + 0.05f * (2 * k), 3.f + 0.05f * (2 * k + 1), 1.f / (k + 1), 0.5f / (k + 1));
    cudaMalloc(&bank, sizeof(h)); cudaMemcpy(bank, h, sizeof(h), cudaMemcpyHostToDevice);
    printf("SM clock %.0f MHz (nominal max), %d SMs\n", clk / 1e6, sms);
    const char *mn[5] = {"MUFU.EX2", "MUFU.RCP", "MUFU.LG2", "4 EX2 : 1 RCP", "MUFU.RSQ"};
    float ms[5];
    ms[0] = time_ms([&] { kmufu<0><<<blocks, threads>>>(out, 1.0f); });
    ms[1] = time_ms([&] { kmufu<1><<<blocks, threads>>>(out, 1.0f); });
    ms[2] = time_ms([&] { kmufu<2><<<blocks, threads>>>(out, 1.0f); });
    ms[3] = time_ms([&] { kmufu<3><<<blocks, threads>>>(out, 1.0f); });
    ms[4] = time_ms([&] { kmufu<4><<<blocks, threads>>>(out, 1.0f); });
    for (int i = 0; i < 5; ++i)
        printf("%-16s %8.3f ms -> %6.2f lane-ops/clk/SM\n", mn[i], ms[i], 2048.0 * 8 * blocks * threads / (ms[i] * 1e-3) / clk / sms);
    const char *vn[14] = {"quad MUFU.RCP (shipped)", "quad MUFU.RCP + per-lane weight table", "quad Newton rcp (3 quadratic)", "quad Newton rcp (2 cubic)",
                         "oct e/(1-e) MUFU.RCP", "oct e/(1-e) Newton (3 quadratic)", "oct e/(1-e) Newton (2 cubic)", "SC4 quad (T, 0.74T) MUFU.RCP",
                         "quad Newton rcp, unroll 4", "quad MUFU.RCP, unroll 4",
                         "quad packed Newton, 2 records/iter", "quad packed Newton, 2 records/iter + table", "SC4 quad packed Newton", "shipped, Wien form w e/(1-e)"};
    auto report = [&](int i, float ms, int nb, int nt) {
        const double samples = (double)nb * nt * REP * K2 * 4 * ((i == 7 || i == 12) ? 2 : 1);
        printf("%-40s %4d thr x %4d CTAs %8.3f ms -> %6.2f Planck samples/clk/SM\n", vn[i], nt, nb, ms, samples / (ms * 1e-3) / clk / sms);
    };
#define RUN(V, T, NB) report(V, time_ms([&] { kloop<V, T><<<NB, T>>>(out, bank, 1.0f); }), NB, T)
    RUN(0, 512, blocks); RUN(1, 512, blocks); RUN(2, 512, blocks); RUN(3, 512, blocks); RUN(4, 512, blocks);
    RUN(5, 512, blocks); RUN(7, 512, blocks); RUN(8, 512, blocks); RUN(9, 512, blocks);
    RUN(10, 512, blocks); RUN(11, 512, blocks); RUN(12, 512, blocks); RUN(13, 512, blocks);
    printf("-- warps per SM sweep (2 CTAs per SM; 1 CTA per SM for the last) --\n");
    RUN(0, 384, blocks); RUN(0, 256, blocks); RUN(0, 128, blocks); RUN(0, 512, sms);
    RUN(2, 384, blocks); RUN(2, 256, blocks); RUN(2, 128, blocks); RUN(2, 512, sms);
    RUN(8, 384, blocks); RUN(8, 256, blocks); RUN(9, 384, blocks); RUN(9, 256, blocks);
    // accuracy of the formulations against double precision for one lane
    return 0;
}
