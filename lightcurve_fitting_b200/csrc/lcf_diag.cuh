// lcf_diag.cuh -- convergence diagnostics on the HBM-resident chain: per-series moments, the
// walker-averaged normalised autocorrelation function (what emcee.autocorr.integrated_time builds with FFTs, here by
// direct lag products: the chain never leaves the device and n_t * max_lag * W * D multiply-adds are a few ms), split
// halves for the Gelman-Rubin statistic.  HBM-bound streaming kernels, FP64.
#pragma once
#include <cuda_runtime.h>

namespace lcf {

// chain layout: x[t][w][d], t in [t0, t0 + n), nwalkers W, ndim D.
// moments per series (w, d): [0] mean, [1] sum (x - mean)^2, [2] mean of first half, [3] sum sq dev of first half,
//                            [4] mean of second half, [5] sum sq dev of second half   (halves of floor(n / 2) steps)
__global__ void __launch_bounds__(256) k_series_moments(const double *__restrict__ x, long long t0, long long n, long long W, int D,
                                                        double *__restrict__ mom) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // series index w * D + d: consecutive threads, consecutive addresses
    if (s >= W * D) return;
    const long long stride = W * D;
    const double *p = x + t0 * stride + s;
    const long long h = n / 2;
    double m = 0., m0 = 0., m1 = 0.;
    const double first = p[0];
    bool moved = false;                                  // a walker that never moved has no autocorrelation function
    for (long long t = 0; t < n; ++t) {
        const double v = p[t * stride];
        moved = moved || (v != first);
        m += v;
        if (t < h) m0 += v; else if (t < 2 * h) m1 += v;
    }
    m /= (double)n;
    m0 = h ? m0 / (double)h : 0.;
    m1 = h ? m1 / (double)h : 0.;
    double q = 0., q0 = 0., q1 = 0.;
    for (long long t = 0; t < n; ++t) {
        const double v = p[t * stride];
        q += (v - m) * (v - m);
        if (t < h) q0 += (v - m0) * (v - m0); else if (t < 2 * h) q1 += (v - m1) * (v - m1);
    }
    double *o = mom + s * 6;
    o[0] = m; o[1] = moved ? q : 0.; o[2] = m0; o[3] = q0; o[4] = m1; o[5] = q1;
}

// f[d][k] = (1 / W) sum_w acf_w(k) / acf_w(0),  acf_w(k) = sum_t (x_t - mean_w)(x_{t+k} - mean_w),  k in [0, nlag)
// grid (ceil(nlag / kLagBlock), D); each thread strides over walkers and keeps kLagBlock accumulators in registers.
constexpr int kLagBlock = 8;
__global__ void __launch_bounds__(256) k_mean_acf(const double *__restrict__ x, long long t0, long long n, long long W, int D,
                                                  const double *__restrict__ mom, long long nlag, double *__restrict__ f) {
    const int d = blockIdx.y;
    const long long k0 = (long long)blockIdx.x * kLagBlock;
    const long long stride = W * D;
    double part[kLagBlock];
#pragma unroll
    for (int j = 0; j < kLagBlock; ++j) part[j] = 0.;
    for (long long w = threadIdx.x; w < W; w += blockDim.x) {
        const double *p = x + t0 * stride + w * D + d;
        const double mean = mom[(w * D + d) * 6], a0 = mom[(w * D + d) * 6 + 1];
        if (!(a0 > 0.)) continue;                        // stuck walker (exactly constant series): emcee would return NaN; left out
        double acc[kLagBlock];
#pragma unroll
        for (int j = 0; j < kLagBlock; ++j) acc[j] = 0.;
        for (long long t = 0; t + k0 < n; ++t) {
            const double v = p[t * stride] - mean;
#pragma unroll
            for (int j = 0; j < kLagBlock; ++j)
                if (t + k0 + j < n) acc[j] = fma(v, p[(t + k0 + j) * stride] - mean, acc[j]);
        }
        const double inv = 1. / a0;
#pragma unroll
        for (int j = 0; j < kLagBlock; ++j) part[j] += acc[j] * inv;
    }
    __shared__ double red[kLagBlock][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kLagBlock; ++j) {
        double v = part[j];
        for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[j][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLagBlock && k0 + threadIdx.x < nlag) {
        double v = 0.;
        for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) v += red[threadIdx.x][w2];
        f[(long long)d * nlag + k0 + threadIdx.x] = v / (double)W;
    }
}

}  // namespace lcf
