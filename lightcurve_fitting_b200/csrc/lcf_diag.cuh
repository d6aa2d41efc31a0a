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

// ---- bolometric post-processing of a batch of SED chains (bolometric.py:792-798, 422-480) --------------------------------------
// One CTA per epoch.  For the four per-sample quantities T, R, L_bol = 4 pi R^2 sigma_SB T^4 (stefan_boltzmann) and the
// pseudo-bolometric luminosity (computed beforehand by the Planck-sum pass on the 1-THz comb, `lpseudo`), the samples of the epoch's
// flat chain are sorted in shared memory (bitonic network, padded with +inf) and the percentiles 50 -+ perc/2 are read off with
// numpy's linear interpolation; out[epoch][quantity] = (median, median - lower, upper - median) = median_and_unc.
__device__ __forceinline__ double percentile_sorted(const double *a, long long n, double q) {
    const double pos = q * 0.01 * (double)(n - 1);                     // numpy 'linear': virtual index q/100 (n-1)
    long long lo = (long long)floor(pos);
    if (lo < 0) lo = 0;
    if (lo > n - 1) lo = n - 1;
    const long long hi = lo + 1 < n ? lo + 1 : n - 1;
    const double t = pos - (double)lo, x = a[lo], y = a[hi];
    return t < 0.5 ? x + (y - x) * t : y - (y - x) * (1. - t);            // numpy's _lerp
}
__global__ void __launch_bounds__(256) k_batch_summary(const double *__restrict__ chain, const double *__restrict__ lpseudo, long long n, int D,
                                                        int npad, double sigma_sb, double perc, double *__restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *a = reinterpret_cast<double *>(smem_raw);
    const long long e = blockIdx.x;
    const double *c = chain + e * n * D;
    const double *lp = lpseudo + e * n;
    for (int q = 0; q < 4; ++q) {
        for (int i = threadIdx.x; i < npad; i += blockDim.x) {
            double v = __longlong_as_double(0x7ff0000000000000LL);
            if (i < n) {
                const double T = c[(long long)i * D], R = c[(long long)i * D + 1];
                if (q == 0) v = T;
                else if (q == 1) v = R;
                else if (q == 2) { const double T2 = T * T; v = 4. * 3.14159265358979323846 * (R * R) * sigma_sb * (T2 * T2); }   // bolometric.py:447
                else v = lp[i];
                if (v != v) v = __longlong_as_double(0x7ff0000000000000LL);     // NaN sorts last (numpy would return NaN: flagged by the status)
            }
            a[i] = v;
        }
        __syncthreads();
        for (int k = 2; k <= npad; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                    const int l = i ^ j;
                    if (l > i) {
                        const double x = a[i], y = a[l];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { a[i] = y; a[l] = x; }
                    }
                }
                __syncthreads();
            }
        if (threadIdx.x == 0) {
            const double lo = percentile_sorted(a, n, 50. - 0.5 * perc), md = percentile_sorted(a, n, 50.), hi = percentile_sorted(a, n, 50. + 0.5 * perc);
            double *o = out + (e * 4 + q) * 3;
            o[0] = md; o[1] = md - lo; o[2] = hi - md;
        }
        __syncthreads();
    }
}

// Shared ensembles: a rank receives only ITS walkers (logical [first, first + count)) from the host ...
__global__ void __launch_bounds__(256) k_set_state_slice(const double *__restrict__ in, long long first, long long count, int D, long long n0,
                                                         double *__restrict__ coords, int *__restrict__ flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * D) return;
    const long long j = first + idx / D;
    const int d = (int)(idx % D);
    const long long r = (j & 1) ? n0 + (j >> 1) : (j >> 1);
    const double v = in[idx];
    if (isinf(v)) atomicOr(flags, 1);
    if (isnan(v)) atomicOr(flags, 2);
    coords[r * D + d] = v;
}
// ... and, once their log-probabilities are known, stores both straight into every peer's replica (NVLink peer memory): the same
// posted stores the accept epilogue of k_pass uses, so no rank ever uploads or gathers the full start array.
struct PublishDev {
    int npeers;
    double *peer_coords[kMaxPeers];
    double *peer_logp[kMaxPeers];
};
__global__ void __launch_bounds__(256) k_publish_rows(const double *__restrict__ coords, const double *__restrict__ logp, long long row0,
                                                      long long nrows, int D, PublishDev P) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nrows * (D + 1)) return;
    if (idx < nrows * D) {
        const double v = coords[row0 * D + idx];
        for (int p = 0; p < P.npeers; ++p) P.peer_coords[p][row0 * D + idx] = v;
    } else {
        const long long r = row0 + (idx - nrows * D);
        const double v = logp[r];
        for (int p = 0; p < P.npeers; ++p) P.peer_logp[p][r] = v;
    }
}

}  // namespace lcf

// ---------------------------------------------------------------------------------------
// Batched bounded least-squares blackbody fits: the reference fits every epoch's SED with
// scipy.optimize.curve_fit (planck_fast at the effective frequencies, unweighted, box bounds on T and R;
// bolometric.py:483-531), one Python call per epoch.  Here: one thread per epoch, Levenberg-Marquardt with Marquardt
// scaling and an active set for the bounds, iterated to machine precision, then scipy's covariance
// (J^T J)^-1 * RSS / (n - 2).  Agrees with curve_fit's trust-region-reflective result to its own tolerance (1e-8
// requested there; 7e-7 observed on 400 random SEDs, bound-hitting ones included).
// ---------------------------------------------------------------------------------------
namespace lcf {

__device__ inline double bb_residuals(const double *nu, const double *y, int n, double T, double R, double c1, double c2, double cutoff,
                                      double *A, double *g) {
    // cost = 1/2 sum r^2; A = J^T J (A[0] TT, A[1] TR, A[2] RR); g = J^T r
    double cost = 0., a0 = 0., a1 = 0., a2 = 0., g0 = 0., g1 = 0.;
    for (int i = 0; i < n; ++i) {
        const double x = c1 * nu[i] / T;
        const double em1 = expm1(x);
        const double f = R * R * c2 * nu[i] * nu[i] * nu[i] * fmin(1., cutoff / nu[i]) / em1;
        const double dT = f * (x / T) * (em1 + 1.) / em1, dR = 2. * f / R;
        const double r = f - y[i];
        cost += 0.5 * r * r;
        a0 += dT * dT; a1 += dT * dR; a2 += dR * dR;
        g0 += dT * r; g1 += dR * r;
    }
    A[0] = a0; A[1] = a1; A[2] = a2;
    g[0] = g0; g[1] = g1;
    return cost;
}

__global__ void k_bb_lstsq(long long nepochs, const int *__restrict__ off, const double *__restrict__ nu, const double *__restrict__ lum,
                           double c1, double c2, double cutoff, double T0, double R0, double Tlo, double Thi, double Rlo, double Rhi,
                           double *__restrict__ popt, double *__restrict__ pcov, int *__restrict__ status) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nepochs) return;
    const double *x = nu + off[e], *y = lum + off[e];
    const int n = off[e + 1] - off[e];
    const double lo[2] = {Tlo, Rlo}, hi[2] = {Thi, Rhi};
    double p[2] = {fmin(fmax(T0, Tlo), Thi), fmin(fmax(R0, Rlo), Rhi)};
    double A[3], g[2];
    double cost = bb_residuals(x, y, n, p[0], p[1], c1, c2, cutoff, A, g);
    double lam = 1e-3;
    int st = 1;
    for (int it = 0; it < 200 && st; ++it) {
        bool fr[2];
        for (int d = 0; d < 2; ++d) fr[d] = !((p[d] <= lo[d] && g[d] > 0.) || (p[d] >= hi[d] && g[d] < 0.));
        if (!fr[0] && !fr[1]) { st = 0; break; }
        bool improved = false;
        double dp = 0., dc = 0.;
        for (int tr = 0; tr < 30; ++tr) {
            const double m0 = A[0] * (1. + lam), m2 = A[2] * (1. + lam), m1 = A[1];
            double s[2] = {0., 0.};
            if (fr[0] && fr[1]) {
                const double det = m0 * m2 - m1 * m1;
                s[0] = -(m2 * g[0] - m1 * g[1]) / det;
                s[1] = -(m0 * g[1] - m1 * g[0]) / det;
            } else if (fr[0]) s[0] = -g[0] / m0;
            else s[1] = -g[1] / m2;
            double q[2];
            for (int d = 0; d < 2; ++d) q[d] = fmin(fmax(p[d] + s[d], lo[d]), hi[d]);
            double An[3], gn[2];
            const double cn = bb_residuals(x, y, n, q[0], q[1], c1, c2, cutoff, An, gn);
            if (cn < cost) {
                improved = true;
                dp = fmax(fabs(q[0] - p[0]) / fabs(p[0]), fabs(q[1] - p[1]) / fabs(p[1]));
                dc = cost > 0. ? (cost - cn) / cost : 0.;
                p[0] = q[0]; p[1] = q[1];
                A[0] = An[0]; A[1] = An[1]; A[2] = An[2];
                g[0] = gn[0]; g[1] = gn[1];
                cost = cn;
                lam = fmax(lam * 0.1, 1e-12);
                break;
            }
            lam *= 10.;
            if (lam > 1e12) break;
        }
        if (!improved || dp < 1e-13 || dc < 1e-16) st = 0;             // converged (no further decrease is possible / needed)
    }
    popt[2 * e] = p[0];
    popt[2 * e + 1] = p[1];
    double *cv = pcov + 4 * e;
    if (n > 2) {
        const double det = A[0] * A[2] - A[1] * A[1], s2 = 2. * cost / (double)(n - 2);
        cv[0] = A[2] / det * s2; cv[1] = cv[2] = -A[1] / det * s2; cv[3] = A[0] / det * s2;
    } else {
        const double inf = __longlong_as_double(0x7ff0000000000000LL);   // scipy: covariance could not be estimated
        cv[0] = cv[1] = cv[2] = cv[3] = inf;
    }
    status[e] = st;
}

// Start positions arrive in emcee's logical walker order; the ensemble lives colour-major (even walkers, then odd).  The
// permutation and emcee's initial-state checks (no NaN, no infinity) run on the device instead of a host loop over W x D.
__global__ void __launch_bounds__(256) k_set_state(const double *__restrict__ in, const double *__restrict__ lp_in, long long W, int D,
                                                   long long n0, double *__restrict__ coords, double *__restrict__ logp,
                                                   int *__restrict__ flags) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= W * D) return;
    const long long j = idx / D;
    const int d = (int)(idx - j * D);
    const long long r = (j & 1) ? n0 + (j >> 1) : (j >> 1);
    const double v = in[idx];
    if (isinf(v)) atomicOr(flags, 1);
    if (isnan(v)) atomicOr(flags, 2);
    coords[r * D + d] = v;
    if (lp_in && d == 0) logp[r] = lp_in[j];
}

}  // namespace lcf
