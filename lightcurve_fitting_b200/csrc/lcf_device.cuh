// lcf_device.cuh -- device code of the fused log-posterior + stretch-move kernels (sm_100a).
//
// Work decomposition:
//   CTA      = one group of WPB walkers of the active half-ensemble (WPB = 2^k <= 32; the chain kernel also runs
//              "wide" groups of 64..256 walkers: WPB/32 walker columns of warps); a thread-block cluster of S CTAs
//              can share one group and split its light curve (partials to rank 0 through DSMEM)
//   lane     = (walker-in-group wl = lane % WPB, point slot = lane / WPB); a lane serves the
//              SAME walker for the whole kernel, so the per-walker model constants live in
//              registers and the chi-square partial sums need no atomics
//   tile     = 2 * 32/WPB photometry points that share one filter (points are grouped by filter
//              on the host), two per lane, so every lane of a warp walks the same transmission curve and the
//              shared-memory reads of (a_k, w_k) are pure broadcasts
//   warp     = strides over the tiles of the light curve
// The packed filter bank is staged into shared memory once per CTA with a 1-D TMA bulk copy
// (cp.async.bulk + mbarrier).  Inner loop per quad of Planck samples (two points x two samples, FP32 mode):
//   FMUL2 x = a_k invT ; MUFU.EX2 x4 ; FADD2 -1 ; products ; Newton reciprocal on the FMA pipe (packed) ; FFMA2 sums
// -- one MUFU per sample, everything else off the XU pipe (section "FP32 fast paths" below).
// Multi-GPU: the accept epilogue stores accepted walkers into the peers' replicas (NVLink peer memory) and half-steps
// are ordered by device-side flags (peers_wait / peers_publish).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

namespace lcf {

#ifdef LCF_X_TIMING   // experiment builds: per-phase SM clocks summed over CTAs (thread 0), see tools/microbench
__device__ unsigned long long g_phase_clk[10];
__device__ unsigned long long g_cta_log[3 * 4096];    // per CTA of the last k_pass launch: start, end (globaltimer ns), SM id
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int smid() { unsigned int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define LCF_TICK(i) do { if (threadIdx.x == 0) { long long _t = clock64(); atomicAdd(&g_phase_clk[i], (unsigned long long)(_t - _t0)); _t0 = _t; } } while (0)
#define LCF_TICK_INIT long long _t0 = clock64()
#else
#define LCF_TICK(i) do {} while (0)
#define LCF_TICK_INIT do {} while (0)
#endif

constexpr int kMaxDim = 12;
constexpr int kMaxPeers = 7;   // one 8-GPU NVSwitch box
constexpr int kNumWC = 10; // per-walker model constants held in registers (wc7 = sigma^2 when use_sigma)

enum Mode : int { MODE_MOVE = 0, MODE_LOGPOST = 1, MODE_LOGLIKE = 2, MODE_MODEL = 3 };

struct PriorDev {
    int kind[kMaxDim];
    double pmin[kMaxDim], pmax[kMaxDim], mean[kMaxDim], std[kMaxDim];
};

// Device view of one problem.  `real` arrays are float (FP32 mode) or double (FP64 mode).
struct ProblemDev {
    int model, ndim, nmodel, use_sigma;
    int npoints, nfilters, nsamples, spl_nint;
    const void *bank;      // real4[nsamples/2] pair records (a0, a1, w0, w1), a = alpha*log2(e), w = w/scale
    const void *kappa;     // real[nsamples]: 0.4*log2(10)*kappa_k           (ShockCooling3)
    const int4 *finfo;     // [nfilters]: (first pair record, pair records, bits of float min a, bits of float max a)
    const int *frole;      // [nfilters]
    const void *spl;       // real4[nfilters][spl_nint]                       (CompanionShocking*)
    const double *t;       // [npoints]
    const float2 *t32;     // [npoints] t - tref as an unevaluated float sum hi + lo   (FP32 mode)
    const int *pfilt;      // [npoints] filter of each point
    const void *obs;       // real4[npoints]: (y/scale, use_sigma ? (dy/scale)^2 : scale/dy, (sigma_units/scale)^2, 0)
    double spl_x0, spl_dx;
    double tref;           // reference epoch of t32
    double const_term;     // sum_i log(2 pi dy_i^2)  or  N*(log 2 pi + 2 log scale)
    double scale;          // FP32 unit scale (1 in FP64 mode)
    double kB, c3sq;       // models.py:10-11 constants (host computed)
    double mc[16];         // model constants
    double dk[4];          // derived exponents, FP64 kernels:  SW family: eps_T, eps_L - 4 eps_T, alpha;  SC4: A, alpha
    float fk[4];           // the same as floats (no F2F.F32.F64 on the hot path: conversions share the XU pipe)
    PriorDev prior;
};

struct TileDev {           // one tile table per WPB
    const int4 *tiles;     // (first point, count, filter, 0)
    int ntiles;
};

// byte offsets of the shared-memory carve-up (SmemLayout below); k_pass receives them from the host in its parameter block
// (MoveDev::lay), so that the per-tile code re-reads an offset from the constant bank instead of recomputing the layout arithmetic
// (the hot kernels have no register to spare)
struct SmemOffsets {
    unsigned int off_e2t, off_bank, off_spl, off_tab, off_foff, off_wc, off_t, off_q, off_lp, off_z, off_old, off_term, off_part, off_cpart, off_flag, off_bar, total;
};

struct MoveDev {
    SmemOffsets lay;                 // filled by launch_pass for the chosen shape
    double *coords;                  // [W][D] colour-major rows
    double *logp;                    // [W]
    unsigned long long *accepted;    // [W] indexed by logical walker
    int *nanflag;
    long long W, n0;                 // n0 = rows of colour 0
    int mode, wpb_log2;
    int ks;                          // log2 of the sample chunks a (walker, point pair) is split into (split-K tiles, small problems)
    // Structured chi-square sums (nq > 1): the tile rows of a light curve are dealt to nq "units" (row r belongs to unit r % nq), a
    // walker's chi-square is  sum_q ( sum_warps P(q, warp) )  in that order whichever CTA computed which unit.  A launch may then give
    // every CTA the same number of units (a flat split of the group x unit space: no partial last wave); a group shared by several
    // CTAs is finished by the one that arrives last (partial sums in split_part, arrival counter in split_tick).
    int nq;
    double *split_part;              // [groups][nq][walkers per CTA] unit sums of the groups shared by several CTAs
    unsigned int *split_tick;        // [groups] units delivered so far (zero between launches)
    // active set: physical row = act_rows ? act_rows[i] : act_base + i,  i in [0, Ns)
    long long Ns, act_base;
    const int *act_rows;
    long long Nc, comp_base;
    const int *comp_rows;
    // injected draws (replay mode), indexed by i; NULL -> Philox
    const double *zin;
    const int *rin;
    const double *luin;
    unsigned long long seed;
    unsigned int ctr;                // 2*iteration + half
    // evaluation modes
    const double *qin;               // [Ns][D]  (nmodel columns in MODE_MODEL)
    int qstride;                     // row stride of qin in doubles (0: the column count) -- lets a pass read a stored chain in place
    double *out;                     // [Ns] or [Ns][npoints]
    // fused multi-GPU exchange (npeers = 0: single GPU, or the host exchanges with NCCL): every rank holds a full replica;
    // the accept epilogue stores accepted walkers into the peers' replicas over NVLink (peer memory mapped with cudaIpc),
    // the last CTA publishes "half-step done" into the peers' flag arrays, the next launch waits on its own flags.
    int npeers, myrank;
    double *peer_coords[kMaxPeers];
    double *peer_logp[kMaxPeers];
    unsigned int *peer_flags[kMaxPeers];   // flag array of peer p (we write slot `myrank`)
    int peer_rank[kMaxPeers];
    unsigned int *flags;                   // ours: flags[r] = half-steps rank r has completed and published
    unsigned int *done_count;              // CTAs of this launch that have finished
    unsigned int epoch;                    // half-steps completed before this launch
    int *xstatus;                          // set to 1 when the wait below times out (a peer died)
    // chain write-back for this iteration (NULL = not stored), logical walker order
    double *chain_step;              // [W][D]
    double *lnp_step;                // [W]
};

// ---------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: UBLKCP), MUFU approximations
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

template <typename R> struct Mth;
template <> struct Mth<float> {
    static __device__ __forceinline__ float ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fffffff); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
};
template <> struct Mth<double> {
    static __device__ __forceinline__ double ex2(double x) { return exp2(x); }
    static __device__ __forceinline__ double lg2(double x) { return log2(x); }
    static __device__ __forceinline__ double rcp(double x) { return 1.0 / x; }
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
};
template <typename R> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };
template <typename R> struct Vec4;
template <> struct Vec4<float> { typedef float4 type; };
template <> struct Vec4<double> { typedef double4 type; };

constexpr int kTabStride = 32;      // ShockCooling3 weight table [pair record][walker column]: 32 columns at compile time in the
                                    // 32-walker instantiation, min(wpb, 32) at run time otherwise (a big bank still fits at small wpb)
constexpr double kLog2e = 1.4426950408889634074;
constexpr double kLn2 = 0.69314718055994530942;

// ---------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = (walker, ctr, 0, 0), key = seed
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// The draws of one stretch move (emcee's StretchMove with a = 2): z = ((a - 1) u + 1)^2 / a, the partner's index in the complementary
// half-ensemble of Nc walkers, and the uniform number of the accept test; keyed by (seed, 2 iteration + half, logical walker).
__device__ __forceinline__ void stretch_draw(unsigned long long seed, unsigned int ctr, long long j, long long Nc, double &z, long long &pr, double &uacc) {
    uint32_t r[4];
    philox4x32_10((uint32_t)j, ctr, (uint32_t)(j >> 32), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    double u = ((double)(r[0] >> 5) * 67108864.0 + (double)(r[1] >> 6)) * (1.0 / 9007199254740992.0);
    double a = u + 1.;                   // ((a-1) u + 1)^2 / a with a = 2
    z = a * a * 0.5;
    pr = (long long)(((unsigned long long)r[2] * (unsigned long long)Nc) >> 32);
    uacc = ((double)r[3] + 0.5) * (1.0 / 4294967296.0);
}
// accept test of the stretch move: (D - 1) ln z + ln p' - ln p > ln u
__device__ __forceinline__ bool accept_test(int D, double lnz, double nlp, double old, double logu) {
    const double lnpdiff = (double)(D - 1) * lnz + nlp - old;
    return lnpdiff > logu;
}

// ---------------------------------------------------------------------------------------
// Lean FP64 pow for the per-walker model constants.  libm's pow is a ~300-instruction dependent chain (it carries
// log2(x) in extended precision to be correctly rounded for any exponent); the constants here need ~1e-12 (the FP64
// parity tolerance is 1e-9 on the log-posterior) and have |y log2 x| < 400, so  x^y = 2^(y log2 x)  with a plain
// double log2 is good to ~1e-13: log2 by the atanh series on the mantissa in [sqrt(1/2), sqrt 2), 2^t by range
// reduction and a degree-12 polynomial.  ~65 dependent operations.  Anything unusual (non-positive, subnormal or
// non-finite base, huge result) goes to libm, so the special-case semantics are libm's.
// ---------------------------------------------------------------------------------------
// 1/x for a positive normal double: MUFU.RCP64H seed (20 bits) + two Newton steps; no special-case handling
__device__ __forceinline__ double rcp_f64(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double exp2_core(double t) {               // |t| < 1000
    const double magic = 6755399441055744.0;                           // 1.5 * 2^52
    const double r = t + magic;
    const int n = __double2loint(r);
    const double f = t - (r - magic);
    double p = 2.56784359934881958e-11;
    p = fma(p, f, 4.44553827187081007e-10);
    p = fma(p, f, 7.05491162080112088e-09);
    p = fma(p, f, 1.01780860092396960e-07);
    p = fma(p, f, 1.32154867901443053e-06);
    p = fma(p, f, 1.52527338040598377e-05);
    p = fma(p, f, 1.54035303933816061e-04);
    p = fma(p, f, 1.33335581464284411e-03);
    p = fma(p, f, 9.61812910762847688e-03);
    p = fma(p, f, 5.55041086648215762e-02);
    p = fma(p, f, 2.40226506959100694e-01);
    p = fma(p, f, 6.93147180559945286e-01);
    p = fma(p, f, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
__device__ __forceinline__ double log2_normal(double x) {             // x positive, finite, normal
    int hi = __double2hiint(x);
    int e = (hi >> 20) - 1023;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));   // [1, 2)
    if (m > 1.4142135623730951) { m *= 0.5; e += 1; }                   // [sqrt(1/2), sqrt 2)
    // (m + 1 is in [1.7, 2.42]: MUFU.RCP64H + two Newton steps instead of an IEEE division -- 5 dependent operations for ~25 on the
    //  critical chain of every per-walker constant; the quotient is good to 2e-16)
    const double s = (m - 1.0) * rcp_f64(m + 1.0), s2 = s * s;           // |s| <= 0.1716
    double p = 1. / 23.;
    p = fma(p, s2, 1. / 21.);
    p = fma(p, s2, 1. / 19.);
    p = fma(p, s2, 1. / 17.);
    p = fma(p, s2, 1. / 15.);
    p = fma(p, s2, 1. / 13.);
    p = fma(p, s2, 1. / 11.);
    p = fma(p, s2, 1. / 9.);
    p = fma(p, s2, 1. / 7.);
    p = fma(p, s2, 1. / 5.);
    p = fma(p, s2, 1. / 3.);
    p = fma(p, s2, 1.);
    return fma(2.8853900817779268 * s, p, (double)e);                   // e + (2 / ln 2) atanh(s)
}
__device__ __forceinline__ double pow_fast(double b, double e) {
    const int hi = __double2hiint(b);
    if (hi >= 0x00100000 && hi < 0x7ff00000) {                          // positive, normal, finite
        const double t = e * log2_normal(b);
        if (fabs(t) < 1000.) return exp2_core(t);
    }
    return pow(b, e);
}
__device__ __forceinline__ double log2_fast(double x) {
    const int hi = __double2hiint(x);
    return (hi >= 0x00100000 && hi < 0x7ff00000) ? log2_normal(x) : log2(x);
}

// power() of models.py:42-48: zero for any non-positive (or NaN) base
__device__ __forceinline__ double pw(double b, double e) { return b > 0. ? pow_fast(b, e) : 0.; }

// log-prior of one parameter (models.py:1055-1098): strict bounds, -inf outside
__device__ __forceinline__ double prior_logp(const PriorDev &pr, int d, double p) {
    if (!(pr.pmin[d] < p && p < pr.pmax[d])) return -Mth<double>::inf();
    switch (pr.kind[d]) {
        case 1: return -log(p);
        case 2: { double u = (p - pr.mean[d]) / pr.std[d]; return -0.5 * (u * u); }
        default: return 0.;
    }
}

// ---------------------------------------------------------------------------------------
// per-walker constants (FP64, one thread per walker) -- the model front ends
//   wc[] meaning per model family is documented next to point_model() below
// ---------------------------------------------------------------------------------------
struct WalkerSetup {
    double wc[kNumWC];
    double t0;     // explosion epoch (t_exp)
    double t1;     // SiFTO t_peak (CompanionShocking*)
};

// The FP64 transcendentals of a walker's constants are independent of each other, so the proposal phase computes
// them one per THREAD (walker_term, k < kTerms) instead of one walker per thread (a double-precision pow is a
// ~300-instruction dependent chain: ten of them back to back used to be the serial prologue of every CTA), and
// setup_walker then only combines them with a few multiplications.
constexpr int kMaxTerms = 9;
template <int MODEL> struct ModelTerms { static constexpr int value = (MODEL == 1 || MODEL == 3) ? 4 : (MODEL == 4 ? 9 : ((MODEL >= 5 && MODEL <= 7) ? 3 : 0)); };

// term k of a walker = pow(base, expo) [or power(): 0 for a non-positive base, models.py:42-48] [or cos(base)]:
// ONE pow call per thread whatever k is, so the lanes of a warp never diverge into different pow instances.
template <int MODEL>
__device__ inline double walker_term(const ProblemDev &P, const double *p, int k) {
    const double *mc = P.mc;
    double base = 1., expo = 1.;
    bool guarded = false;                                  // power() semantics
    if (MODEL == 1 || MODEL == 3) {
        // BaseShockCooling.temperature_radius, models.py:260-269 (kappa = 1); mc: A, a, alpha, eps1, eps2, L_0, T_0, Tph_to_Tcol
        const double v = p[0], Menv = p[1], f = p[2], Rr = p[3];
        if (k == 0) { base = v * v / f; expo = mc[3]; guarded = true; }
        else if (k == 1) { base = Rr; expo = 0.25; }
        else if (k == 2) { base = v / f; expo = -mc[4]; guarded = true; }
        else { base = Menv / v; expo = 0.5; }
    } else if (MODEL == 4) {
        // ShockCooling4.temperature_radius, models.py:584-587 (kappa = 1)
        const double v = p[0], f = p[2], Rr = p[3];
        const double ex[8] = {1.26, -1.13, -0.13, 0.78, 2.11, 0.11, -0.32, 0.03};
        // Terms 7 and 8 are whole dependent chains of the constants, walked by their own warps NEXT TO the other powers instead of
        // after them in setup_walker (where they were 1 700 of its 2 660 clocks on the 100-walker ensembles): the right-associative
        // tower v_s ** (0.58 ** (f_rho_M ** 0.03)) of models.py:586, and log2(a / t_tr) with t_tr = t_tr_0 sqrt(M_env / v_s).
        if (k == 7) return pow_fast(v, pow_fast(0.58, pow_fast(f, 0.03)));
        if (k == 8) { const double b = mc[1] / (mc[6] * sqrt(p[1] / v)); return (b > 0.) ? log2_fast(b) : -Mth<double>::inf(); }
        const int which = (0x20210210 >> (4 * k)) & 3;       // k -> 0: R, 1: v, 2: f   (R v f R v f R f)
        base = which == 0 ? Rr : (which == 1 ? v : f);
        expo = ex[k];
    } else if (MODEL >= 5 && MODEL <= 7) {
        // BaseCompanionShocking.temperature_radius, models.py:752-754 (kappa = 1)
        const double Mv7 = (MODEL == 7) ? 1. : p[2];
        if (k == 0) { base = p[1]; expo = 36.; }
        else if (k == 1) { base = Mv7; expo = 1. / 9.; guarded = true; }
        else return (MODEL == 7) ? cos(p[2] * (3.14159265358979323846 / 180.)) : 0.;
    } else {
        return 0.;
    }
    if (guarded && !(base > 0.)) return 0.;
    return pow_fast(base, expo);
}

template <int MODEL>
__device__ inline void setup_walker(const ProblemDev &P, const double *p, const double *term, WalkerSetup &s) {
    const double *mc = P.mc;
    for (int i = 0; i < kNumWC; ++i) s.wc[i] = 0.;
    s.t0 = 0.; s.t1 = 0.;
    const double c3sq = P.c3sq, k_kB = P.kB;
    if (MODEL == 1 || MODEL == 3) {
        double v = p[0], Rr = p[3];
        double texp = (MODEL == 3) ? p[6] : p[4];
        double KT = mc[6] * term[0] * term[1] * mc[7] / k_kB;
        double KL = c3sq * mc[5] * term[2] * v * v * Rr * mc[0];
        if (MODEL == 3) KL = KL / (p[4] * p[4]);          // flux = c4*lum/dist^2 (c4 folded into the bank)
        double ttr = 19.5 * term[3];
        double base = mc[1] / ttr;                        // (a t / t_tr) > 0  <=>  a/t_tr > 0 for t > 0
        s.wc[0] = KT; s.wc[1] = KL;
        s.wc[2] = (base > 0.) ? log2_fast(base) : -Mth<double>::inf();
        if (MODEL == 3) s.wc[3] = p[5];                   // E(B-V)
        s.wc[4] = (KT > 0.) ? 1. / KT : 0.;               // 1/T = wc4 t^-eps_T
        s.wc[5] = (KT > 0.) ? KL / (KT * KT * KT * KT) : 0.;   // L/T^4 = wc5 t^(eps_L - 4 eps_T) exp(-(a t/t_tr)^alpha)
        s.t0 = texp;
    } else if (MODEL == 2) {
        // ShockCooling2.evaluate, models.py:403-407
        double base = mc[1] / p[2];
        s.wc[0] = p[0];
        s.wc[1] = c3sq * p[1] * 1e42;
        s.wc[2] = (base > 0.) ? log2(base) : -Mth<double>::inf();
        s.wc[4] = (p[0] > 0.) ? 1. / p[0] : 0.;
        s.wc[5] = (p[0] > 0.) ? s.wc[1] / (p[0] * p[0] * p[0] * p[0]) : 0.;
        s.t0 = p[3];
    } else if (MODEL == 4) {
        // mc: A, a, alpha, L_br_0, T_col_br_0, t_br_0, t_tr_0
        double t_br = mc[5] * term[0] * term[1] * term[2];
        double L_br = mc[3] * term[3] * term[4] * term[5];
        // models.py:586 as written: v_s ** 0.58 ** f_rho_M ** 0.03 is right-associative (term 7, walker_term)
        double T_br = mc[4] * term[6] * term[7];
        s.wc[0] = T_br / k_kB;
        s.wc[1] = c3sq * L_br;
        s.wc[2] = (t_br > 0.) ? -log2_fast(t_br) : Mth<double>::nan();   // log2(ttilde) = log2(t) + wc2
        s.wc[3] = term[8];                                               // log2(a / t_tr) | -inf
        s.wc[4] = 1. / (0.97 * s.wc[0]);                    // 1/T on the early branch: wc4 ttilde^(1/3)
        s.wc[5] = 1. / s.wc[0];                             // 1/T on the late branch:  wc5 ttilde^0.45
        s.t0 = p[4];
    } else if (MODEL == 5 || MODEL == 6 || MODEL == 7) {
        double Mv7 = (MODEL == 7) ? 1. : p[2];
        double cK = term[0] * Mv7;
        s.wc[0] = 25. * pw(cK, 1. / 144.);                 // T = wc0 * t^(-74/144)
        double rk = 2.7 * term[1];
        s.wc[1] = rk * rk;                                  // R^2 = wc1 * t^(14/9)
        s.wc[2] = 1.;
        if (MODEL == 7) {                                   // models.py:1042-1043
            double th = p[2] * (3.14159265358979323846 / 180.);
            s.wc[2] = (0.5 * term[2] + 0.5) * (0.14 * (th * th) - 0.4 * th + 1.);
        }
        s.wc[3] = p[4];                                     // stretch
        s.wc[9] = 1. / p[4];
        s.wc[8] = (s.wc[0] > 0.) ? 1. / s.wc[0] : 0.;       // 1/T = wc8 t^(74/144)
        if (MODEL == 5) { s.wc[4] = p[5]; s.wc[5] = p[6]; s.wc[6] = p[7]; }   // r_r, r_i, r_U
        else            { s.wc[4] = p[5]; s.wc[5] = p[6]; }                   // dt_U, dt_i
        s.t0 = p[0];
        s.t1 = p[3];
    } else if (MODEL == 8) {
        // planck_fast(nu, T, R), models.py:1127-1128
        s.wc[0] = (p[0] > 0.) ? 1. / p[0] : 0.;             // power(T, -1)
        s.wc[1] = p[1] * p[1];
    }
}

// ---------------------------------------------------------------------------------------
// Planck x transmission sums:  S(invT) = sum_k w_k / (2^(a_k invT) - 1)
//
// Bank layout: every filter is padded to an even number of samples (pad: a = last a, w = 0); the bank is an array of
// PAIR records (a0, a1, w0, w1), 16 bytes in FP32: one LDS.128 broadcast feeds four Planck samples (two points).
// ShockCooling3's per-walker reddened weights live in a pair table tab2[pair][walker] (one conflict-free LDS.64 per
// lane per pair); the kernel then reads only the (a0, a1) half of the record.
// ---------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ R planck_term_safe(R a, R invT) {
    // power(exp(x) - 1, -1) semantics (models.py:1128): 0 when exp(x)-1 is 0 or inf
    if (sizeof(R) == 4) {
        float x = (float)(a * invT);
        float d = (x < 0.0625f) ? expm1f(x * (float)kLn2) : (Mth<float>::ex2(x) - 1.f);
        return (d > 0.f) ? (R)(1.f / d) : (R)0;
    } else {
        double d = exp2((double)(a * invT)) - 1.0;
        return (d > 0.) ? (R)(1.0 / d) : (R)0;
    }
}

// Careful path (any precision): one blackbody, exact guards.  K2 = number of sample pairs.
template <typename R, bool TAB>
__device__ __forceinline__ R planck_sum_safe(const typename Vec4<R>::type *__restrict__ b, int K2, R invT,
                                             const typename Vec2<R>::type *__restrict__ tab, int ts) {
    R acc0 = 0, acc1 = 0;
    for (int k = 0; k < K2; ++k) {
        const typename Vec4<R>::type s = b[k];
        R w0 = s.z, w1 = s.w;
        if (TAB) { typename Vec2<R>::type t = tab[k * ts]; w0 = t.x; w1 = t.y; }
        acc0 = fma(w0, planck_term_safe<R>(s.x, invT), acc0);
        acc1 = fma(w1, planck_term_safe<R>(s.y, invT), acc1);
    }
    return acc0 + acc1;
}

// FP32 fast paths (sm_100a).  The loop is bound by the XU pipe (MUFU: 16 lanes/clk/SM, 8 clk per warp instruction),
// so everything except the exponential itself is kept OFF that pipe:
//  * four Planck denominators share ONE reciprocal:  r = 1/(d0 d1 d2 d3);  1/(d0 d1) = r (d2 d3);  1/d0 = d1/(d0 d1);
//  * that reciprocal is computed on the FMA/ALU pipes: integer-subtract seed (12 % error) + three Newton steps
//    (4e-8, the accuracy of MUFU.RCP), two reciprocals at a time in the packed lanes;
//  * the FP32 arithmetic uses Blackwell's packed FMUL2 / FADD2 / FFMA2 (one issue slot for two lanes of work).
// => 1 MUFU.EX2 per Planck sample and nothing else on the XU pipe.  Measured in isolation (tools/microbench/loops.cu,
// B200, 32 warps/SM): 14.6 samples/clk/SM against 12.2 for the MUFU.RCP version and 15.7 for a bare EX2 stream.
// Callers guarantee that every exponent is >= 1/16 (no cancellation in 2^x - 1) and either that the four exponents
// sum to <= 124 (the product of four denominators stays a normal float) or they ask for the WIEN variant.
// Two variants of every loop.  Plain: d = 2^x - 1, valid while the four exponents of a quad sum to <= 124 (their product
// stays a normal float).  WIEN (cold blackbody / blue filter, any exponent): the same quad on m = 1 - 2^-x, which lives in
// (0.04, 1], and  w/(2^x - 1) = w e/(1 - e)  with e = 2^-x: no overflow at all and the exact Wien limit w 2^-x when e
// underflows, for one more multiplication per sample.
template <bool WIEN>
__device__ __forceinline__ void ex2_terms(float2 a, float2 i2, float2 &d, float2 &e) {
    // plain: d = 2^(a i) - 1 (e unused);  WIEN (i2 holds -i): e = 2^(-a i), d = 1 - e
    const float2 x = __fmul2_rn(a, i2);
    e = make_float2(Mth<float>::ex2(x.x), Mth<float>::ex2(x.y));
    d = WIEN ? __fadd2_rn(make_float2(-e.x, -e.y), make_float2(1.f, 1.f)) : __fadd2_rn(e, make_float2(-1.f, -1.f));
}
#ifndef LCF_RCP_ORDER
#define LCF_RCP_ORDER 6
#endif
// 1/x for a positive normal float WITHOUT the XU pipe: integer-subtract seed (relative error e0 within +-5.05 %), then
//   order 6 (default): r (1 + e0)(1 + e0^2 + e0^4) = r (1 + e0 + ... + e0^5): five FMA-pipe operations, dependency depth 4,
//                      truncation 1.7e-8, max error with rounding 1.4e-7 (1 ulp, the accuracy of MUFU.RCP);
//   order 8: three Newton steps, six dependent operations, 6e-8.
__device__ __forceinline__ float rcp_newton(float x) {                   // x positive and normal
    float r = __int_as_float(0x7EF311C7 - __float_as_int(x));
    float e = fmaf(-x, r, 1.f);
#if LCF_RCP_ORDER == 6
    const float r1 = fmaf(r, e, r), e2 = e * e;
    return fmaf(r1, fmaf(e2, e2, e2), r1);
#else
    r = fmaf(r, e, r);
    e = fmaf(-x, r, 1.f); r = fmaf(r, e, r);
    e = fmaf(-x, r, 1.f); r = fmaf(r, e, r);
    return r;
#endif
}
__device__ __forceinline__ float2 rcp_newton2(float2 x) {
    float2 r = make_float2(__int_as_float(0x7EF311C7 - __float_as_int(x.x)), __int_as_float(0x7EF311C7 - __float_as_int(x.y)));
    const float2 one = make_float2(1.f, 1.f), nx = make_float2(-x.x, -x.y);
    float2 e = __ffma2_rn(nx, r, one);
#if LCF_RCP_ORDER == 6
    const float2 r1 = __ffma2_rn(r, e, r), e2 = __fmul2_rn(e, e);
    return __ffma2_rn(r1, __ffma2_rn(e2, e2, e2), r1);
#else
    r = __ffma2_rn(r, e, r);
    e = __ffma2_rn(nx, r, one); r = __ffma2_rn(r, e, r);
    e = __ffma2_rn(nx, r, one); r = __ffma2_rn(r, e, r);
    return r;
#endif
}

// 2^x on the FMA pipe (no MUFU): Cody-Waite split x = n + f, |f| <= 1/2 (magic-number rounding), degree-5 minimax
// polynomial of 2^f (relative error 1.8e-7 in FP32 Horner form: the accuracy of MUFU.EX2), n added into the exponent field
// on the ALU pipe.  8 FMA-pipe operations + 2 ALU operations.  Valid for |x| <= 126 (callers clamp where x can leave it).
#ifndef LCF_LOOP_POLY
#define LCF_LOOP_POLY 0        // inner loops: 0 = every exponential on MUFU; 8 / 4 = one in eight / four on the FMA pipe
#endif
#ifndef LCF_FE_POLY
#define LCF_FE_POLY 0          // front end: 1 = its exponentials on the FMA pipe (lg2 stays on MUFU)
#endif
__device__ __forceinline__ float ex2_fma(float x) {
    const float magic = 12582912.f;                                      // 1.5 * 2^23
    const float r = __fadd_rn(x, magic);
    const float f = __fsub_rn(x, __fsub_rn(r, magic));
    float p = 0.0013292921939864755f;
    p = fmaf(p, f, 0.009671508334577084f);
    p = fmaf(p, f, 0.05550636723637581f);
    p = fmaf(p, f, 0.24022242426872253f);
    p = fmaf(p, f, 0.6931470632553101f);
    p = fmaf(p, f, 1.f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
__device__ __forceinline__ float2 ex2_fma2(float2 x) {
    const float2 magic = make_float2(12582912.f, 12582912.f), nmagic = make_float2(-12582912.f, -12582912.f);
    const float2 r = __fadd2_rn(x, magic);
    const float2 n = __fadd2_rn(r, nmagic);
    const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
    float2 p = make_float2(0.0013292921939864755f, 0.0013292921939864755f);
    p = __ffma2_rn(p, f, make_float2(0.009671508334577084f, 0.009671508334577084f));
    p = __ffma2_rn(p, f, make_float2(0.05550636723637581f, 0.05550636723637581f));
    p = __ffma2_rn(p, f, make_float2(0.24022242426872253f, 0.24022242426872253f));
    p = __ffma2_rn(p, f, make_float2(0.6931470632553101f, 0.6931470632553101f));
    p = __ffma2_rn(p, f, make_float2(1.f, 1.f));
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(r.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(r.y) << 23)));
}
// the same terms as ex2_terms with part of the exponentials on the FMA pipe: HOW = 1: lane y, HOW = 2: both lanes
template <bool WIEN, int HOW>
__device__ __forceinline__ void ex2_terms_mix(float2 a, float2 i2, float2 &d, float2 &e) {
    float2 x = __fmul2_rn(a, i2);
    if (WIEN) x = make_float2(fmaxf(x.x, -125.f), fmaxf(x.y, -125.f));   // plain: callers keep every exponent in [1/16, 124]
    if (HOW == 2) e = ex2_fma2(x);
    else e = make_float2(Mth<float>::ex2(x.x), ex2_fma(x.y));
    d = WIEN ? __fadd2_rn(make_float2(-e.x, -e.y), make_float2(1.f, 1.f)) : __fadd2_rn(e, make_float2(-1.f, -1.f));
}
template <bool WIEN>
__device__ __forceinline__ void ex2_terms_last(float2 a, float2 i2, float2 &d, float2 &e) {   // the pair that may leave the XU pipe
#if LCF_LOOP_POLY == 8
    ex2_terms_mix<WIEN, 1>(a, i2, d, e);
#elif LCF_LOOP_POLY == 4
    ex2_terms_mix<WIEN, 2>(a, i2, d, e);
#else
    ex2_terms<WIEN>(a, i2, d, e);
#endif
}

// (a) two blackbodies (points A, B of one walker) x the two samples of a pair record; two records per iteration.
//     Pointer-bumped with compile-time strides so that the loop bookkeeping is 2 adds + compare + branch (integer
//     multiply-adds would land on the FMA pipe, which is the second-busiest one here).
template <bool TAB, bool WIEN, int TS>
__device__ __forceinline__ void planck_quad_f32(const float4 *__restrict__ b4, int K2, float iA, float iB,
                                                const float2 *__restrict__ tab, int ts_rt, float &SA, float &SB) {
    const int ts = TS > 0 ? TS : ts_rt;
    const float sA = WIEN ? -iA : iA, sB = WIEN ? -iB : iB;
    const float2 iA2 = make_float2(sA, sA), iB2 = make_float2(sB, sB);
    float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);
    const float4 *pb = b4, *const pe = b4 + (K2 & ~1);
    for (; pb < pe; pb += 2) {
        float2 a0, w0, a1, w1;
        if (TAB) {
            a0 = *reinterpret_cast<const float2 *>(pb);
            a1 = *reinterpret_cast<const float2 *>(pb + 1);
            w0 = tab[0];
            w1 = tab[ts];
            tab += 2 * ts;
        } else {
            const float4 s0 = pb[0], s1 = pb[1];
            a0 = make_float2(s0.x, s0.y); w0 = make_float2(s0.z, s0.w);
            a1 = make_float2(s1.x, s1.y); w1 = make_float2(s1.z, s1.w);
        }
        float2 dA0, dB0, dA1, dB1, eA0, eB0, eA1, eB1;            // d: denominators of (sample k0, sample k1); e: WIEN only
        ex2_terms<WIEN>(a0, iA2, dA0, eA0); ex2_terms<WIEN>(a0, iB2, dB0, eB0);
        ex2_terms<WIEN>(a1, iA2, dA1, eA1); ex2_terms_last<WIEN>(a1, iB2, dB1, eB1);
        const float2 p0 = __fmul2_rn(dA0, dB0), p1 = __fmul2_rn(dA1, dB1);                          // per sample: dA dB
        const float2 r = rcp_newton2(make_float2(p0.x * p0.y, p1.x * p1.y));                         // one per record
        const float2 t0 = __fmul2_rn(w0, __fmul2_rn(make_float2(r.x, r.x), make_float2(p0.y, p0.x)));  // w/(dA dB)
        const float2 t1 = __fmul2_rn(w1, __fmul2_rn(make_float2(r.y, r.y), make_float2(p1.y, p1.x)));
        accA = __ffma2_rn(WIEN ? __fmul2_rn(t0, eA0) : t0, dB0, accA);                               // += w [eA]/dA
        accB = __ffma2_rn(WIEN ? __fmul2_rn(t0, eB0) : t0, dA0, accB);                               // += w [eB]/dB
        accA = __ffma2_rn(WIEN ? __fmul2_rn(t1, eA1) : t1, dB1, accA);
        accB = __ffma2_rn(WIEN ? __fmul2_rn(t1, eB1) : t1, dA1, accB);
    }
    if (K2 & 1) {
        float2 a, w;
        if (TAB) { a = *reinterpret_cast<const float2 *>(pb); w = tab[0]; }
        else { const float4 s = pb[0]; a = make_float2(s.x, s.y); w = make_float2(s.z, s.w); }
        float2 dA, dB, eA, eB;
        ex2_terms<WIEN>(a, iA2, dA, eA); ex2_terms<WIEN>(a, iB2, dB, eB);
        const float2 p = __fmul2_rn(dA, dB);
        const float r = rcp_newton(p.x * p.y);
        const float2 t = __fmul2_rn(w, make_float2(r * p.y, r * p.x));
        accA = __ffma2_rn(WIEN ? __fmul2_rn(t, eA) : t, dB, accA);
        accB = __ffma2_rn(WIEN ? __fmul2_rn(t, eB) : t, dA, accB);
    }
    SA = accA.x + accA.y;
    SB = accB.x + accB.y;
}

// (b) ShockCooling4: two points x (T, 0.74 T) (models.py:629-630); each quad = (A, A*, B, B*) at ONE sample, the two
//     samples of the pair record are processed side by side in the packed lanes (two reciprocals per record).
template <bool WIEN>
__device__ __forceinline__ void planck_quad_sc4_f32(const float4 *__restrict__ b4, int K2, float iA, float iB, float &SA,
                                                    float &SAs, float &SB, float &SBs) {
    const float c = (float)(1. / 0.74), sA = WIEN ? -iA : iA, sB = WIEN ? -iB : iB;
    const float2 iA2 = make_float2(sA, sA), iB2 = make_float2(sB, sB), iAs2 = make_float2(sA * c, sA * c), iBs2 = make_float2(sB * c, sB * c);
    float2 a = make_float2(0.f, 0.f), as = a, bb = a, bs = a;
#pragma unroll 2
    for (const float4 *pb = b4, *const pe = b4 + K2; pb < pe; ++pb) {
        const float4 s = *pb;
        const float2 x = make_float2(s.x, s.y), w = make_float2(s.z, s.w);
        float2 dA, dAs, dB, dBs, eA, eAs, eB, eBs;
        ex2_terms<WIEN>(x, iA2, dA, eA); ex2_terms<WIEN>(x, iAs2, dAs, eAs);
        ex2_terms<WIEN>(x, iB2, dB, eB); ex2_terms_last<WIEN>(x, iBs2, dBs, eBs);
        const float2 pA = __fmul2_rn(dA, dAs), pB = __fmul2_rn(dB, dBs);
        const float2 r = rcp_newton2(__fmul2_rn(pA, pB));
        const float2 tA = __fmul2_rn(w, __fmul2_rn(r, pB)), tB = __fmul2_rn(w, __fmul2_rn(r, pA));
        a = __ffma2_rn(WIEN ? __fmul2_rn(tA, eA) : tA, dAs, a);
        as = __ffma2_rn(WIEN ? __fmul2_rn(tA, eAs) : tA, dA, as);
        bb = __ffma2_rn(WIEN ? __fmul2_rn(tB, eB) : tB, dBs, bb);
        bs = __ffma2_rn(WIEN ? __fmul2_rn(tB, eBs) : tB, dB, bs);
    }
    SA = a.x + a.y; SAs = as.x + as.y; SB = bb.x + bb.y; SBs = bs.x + bs.y;
}

// (c) four points of one walker (A, B, C, D) at ONE sample per quad, the two samples of a pair record side by side in the packed lanes:
//     the ShockCooling4 loop with four independent temperatures and, for ShockCooling3, the per-walker weight table.  Used by the
//     32-walker plain kernels of the one-blackbody models (PPL = 4): per-tile work (descriptor, guards, loop set-up) is paid once per
//     four points instead of once per two.
template <bool TAB, bool WIEN, int TS>
__device__ __forceinline__ void planck_quad4_f32(const float4 *__restrict__ b4, int K2, float iA, float iB, float iC, float iD,
                                                 const float2 *__restrict__ tab, int ts_rt, float &SA, float &SB, float &SC, float &SD) {
    const int ts = TS > 0 ? TS : ts_rt;
    const float sA = WIEN ? -iA : iA, sB = WIEN ? -iB : iB, sC = WIEN ? -iC : iC, sD = WIEN ? -iD : iD;
    const float2 iA2 = make_float2(sA, sA), iB2 = make_float2(sB, sB), iC2 = make_float2(sC, sC), iD2 = make_float2(sD, sD);
    float2 a = make_float2(0.f, 0.f), b = a, c = a, d = a;
#pragma unroll 2
    for (const float4 *pb = b4, *const pe = b4 + K2; pb < pe; ++pb) {
        float2 x, w;
        if (TAB) { x = *reinterpret_cast<const float2 *>(pb); w = *tab; tab += ts; }
        else { const float4 s = *pb; x = make_float2(s.x, s.y); w = make_float2(s.z, s.w); }
        float2 dA, dB, dC, dD, eA, eB, eC, eD;
        ex2_terms<WIEN>(x, iA2, dA, eA); ex2_terms<WIEN>(x, iB2, dB, eB);
        ex2_terms<WIEN>(x, iC2, dC, eC); ex2_terms<WIEN>(x, iD2, dD, eD);
        const float2 pAB = __fmul2_rn(dA, dB), pCD = __fmul2_rn(dC, dD);
        const float2 r = rcp_newton2(__fmul2_rn(pAB, pCD));
        const float2 tAB = __fmul2_rn(w, __fmul2_rn(r, pCD)), tCD = __fmul2_rn(w, __fmul2_rn(r, pAB));   // w/(dA dB), w/(dC dD)
        a = __ffma2_rn(WIEN ? __fmul2_rn(tAB, eA) : tAB, dB, a);
        b = __ffma2_rn(WIEN ? __fmul2_rn(tAB, eB) : tAB, dA, b);
        c = __ffma2_rn(WIEN ? __fmul2_rn(tCD, eC) : tCD, dD, c);
        d = __ffma2_rn(WIEN ? __fmul2_rn(tCD, eD) : tCD, dC, d);
    }
    SA = a.x + a.y; SB = b.x + b.y; SC = c.x + c.y; SD = d.x + d.y;
}

// FP64 fast path: the same quad structure in double precision.  2^x by range reduction (x = n + f, |f| <= 1/2) and a
// degree-12 Taylor polynomial of exp(f ln 2) (truncation 2e-16), n added to the exponent field; exponents capped at 250
// (h nu / k T = 173: such a sample is 1e-75 of its weight) so that four denominators multiply to < 2^1000; one division per
// four samples.  ~21 FP64 pipe operations per Planck sample instead of ~55 (libm exp2 + one division each).
// Callers guarantee every exponent >= 2^-10 (2^x - 1 then keeps 1e-13 relative accuracy).
// 2^x - 1 for the FP64 loop, x in [0, 256): x = n + j/4096 + f with |f| <= 2^-13, 2^(j/4096) from a 4096-entry (32 KB) shared-memory
// table (filled from a device-global copy the host writes once per device), 2^f by its degree-2 Taylor polynomial (truncation
// f^3 ln2^3 / 6 <= 1.0e-13): 7 FP64-pipe operations per exponential, 86 per loop iteration of 8 samples (SASS count).  Measured on cfg2
// (profiles/round2_kernel_variants.jsonl): 16-entry table + degree 6 (round 1, 11 operations) 7.86 M walker-steps/s, 256 entries +
// degree 4 8.65 M, 1024 + degree 3 9.17 M (9.55 M with the flat split), 4096 + degree 2 9.97 M, 8192 + degree 2 9.94 M.
#ifndef LCF_E2T_BITS
#define LCF_E2T_BITS 12
#endif
constexpr int kE2TabBits = LCF_E2T_BITS, kE2TabSize = 1 << kE2TabBits;
__device__ double g_e2tab[kE2TabSize];                                    // 2^(j / kE2TabSize), written by the host (lcf_api.cu)
__device__ __forceinline__ void stage_e2tab(double *dst) {               // every thread of the CTA; the caller synchronises
    for (int j = threadIdx.x; j < kE2TabSize; j += blockDim.x) dst[j] = g_e2tab[j];
}
__device__ __forceinline__ double ex2m1_f64(double x, const double *__restrict__ e2t) {
    const double magic = 6755399441055744.0 / (double)kE2TabSize;        // 1.5 * 2^(52 - bits): adding it rounds x to a multiple of 2^-bits
    const double r = x + magic;
    const int k = __double2loint(r);                                     // 2^bits n + j
    const double f = x - (r - magic);
#if LCF_E2T_BITS >= 12
    double p = 2.40226506959100712e-01;                                  // ln2^2/2  (12 bits: |f| <= 2^-13, truncation f^3 ln2^3/6 = 1.0e-13; 13 bits: 1.3e-14)
    p = fma(p, f, 6.93147180559945309e-01);                              // ln2
#elif LCF_E2T_BITS >= 10
    double p = 5.55041086648215800e-02;                                  // ln2^3/6  (|f| <= 2^-11: truncation f^4 ln2^4/24 = 5e-16)
    p = fma(p, f, 2.40226506959100712e-01);                              // ln2^2/2
    p = fma(p, f, 6.93147180559945309e-01);                              // ln2
#elif LCF_E2T_BITS >= 8
    double p = 9.61812910762847716e-03;                                  // ln2^4/24
    p = fma(p, f, 5.55041086648215800e-02);                              // ln2^3/6
    p = fma(p, f, 2.40226506959100712e-01);                              // ln2^2/2
    p = fma(p, f, 6.93147180559945309e-01);                              // ln2
#else
    double p = 1.54035303933816061e-04;
    p = fma(p, f, 1.33335581464284411e-03);
    p = fma(p, f, 9.61812910762847688e-03);
    p = fma(p, f, 5.55041086648215762e-02);
    p = fma(p, f, 2.40226506959100694e-01);
    p = fma(p, f, 6.93147180559945286e-01);
#endif
    const double t = e2t[k & (kE2TabSize - 1)];
    p = fma(t * f, p, t);                                                // 2^(j/2^bits) (1 + f q(f))
    return __hiloint2double(__double2hiint(p) + ((k >> kE2TabBits) << 20), __double2loint(p)) - 1.0;
}


template <bool TAB>
__device__ __forceinline__ void planck_quad_f64(const double4 *__restrict__ b4, int K2, double iA, double iB,
                                                const double2 *__restrict__ tab, int ts, const double *__restrict__ e2t, double &SA, double &SB) {
    double a0 = 0., a1 = 0., b0 = 0., b1 = 0.;
#pragma unroll 2                                         // eight independent exp2 chains per lane: the FP64 pipe's latency needs them at 16 warps/SM
    for (const double4 *pb = b4, *const pe = b4 + K2; pb < pe; ++pb) {
        const double4 s = *pb;
        double w0 = s.z, w1 = s.w;
        if (TAB) { const double2 t = *tab; tab += ts; w0 = t.x; w1 = t.y; }
        const double dA0 = ex2m1_f64(s.x * iA, e2t), dA1 = ex2m1_f64(s.y * iA, e2t);
        const double dB0 = ex2m1_f64(s.x * iB, e2t), dB1 = ex2m1_f64(s.y * iB, e2t);
        const double p0 = dA0 * dB0, p1 = dA1 * dB1;
        const double r = rcp_f64(p0 * p1);
        const double t0 = w0 * (r * p1), t1 = w1 * (r * p0);       // w/(dA dB) of each sample
        a0 = fma(t0, dB0, a0); b0 = fma(t0, dA0, b0);
        a1 = fma(t1, dB1, a1); b1 = fma(t1, dA1, b1);
    }
    SA = a0 + a1;
    SB = b0 + b1;
}

// ---------------------------------------------------------------------------------------
// per-lane walker state in registers
// ---------------------------------------------------------------------------------------
// FP32 mode never converts a double on the hot path (F2F.F32.F64 shares the XU pipe with MUFU): epochs are kept
// relative to P.tref as unevaluated hi + lo float sums, and t - t_0 = (t_hi - t0_hi) + (t_lo - t0_lo) is good to
// one ulp of the DIFFERENCE (an MJD minus an MJD in one float would only be good to one ulp of 59000).
template <typename R> __device__ __forceinline__ R kconst(const ProblemDev &P, int i);
template <> __device__ __forceinline__ double kconst<double>(const ProblemDev &P, int i) { return P.dk[i]; }
template <> __device__ __forceinline__ float kconst<float>(const ProblemDev &P, int i) { return P.fk[i]; }

template <typename R> struct LaneWalker;
template <> struct LaneWalker<double> {
    double wc[kNumWC];
    double t0, t1;
    __device__ __forceinline__ void set_epochs(double a, double b, double) { t0 = a; t1 = b; }
};
template <> struct LaneWalker<float> {
    float wc[kNumWC];
    float t0h, t0l, t1h, t1l;
    __device__ __forceinline__ void set_epochs(double a, double b, double tref) {
        const double ra = a - tref, rb = b - tref;
        t0h = (float)ra; t0l = (float)(ra - (double)t0h);
        t1h = (float)rb; t1l = (float)(rb - (double)t1h);
    }
};
// time of photometry point p since the walker's explosion epoch (WHICH = 0) or SiFTO peak (WHICH = 1)
template <int WHICH> __device__ __forceinline__ double since(const ProblemDev &P, const LaneWalker<double> &w, int p) {
    return P.t[p] - (WHICH ? w.t1 : w.t0);
}
template <int WHICH> __device__ __forceinline__ float since(const ProblemDev &P, const LaneWalker<float> &w, int p) {
    const float2 t = P.t32[p];
    return __fadd_rn(__fsub_rn(t.x, WHICH ? w.t1h : w.t0h), __fsub_rn(t.y, WHICH ? w.t1l : w.t0l));
}

// SiFTO cubic spline (models.py:717, 817-826): NaN outside the knots -> 0
// `spl` is the shared-memory copy of the coefficient table (staged by TMA with the filter bank); knot origin, spacing and
// its reciprocal come as `real` constants (kconst 0..2 for these models).
template <typename R>
__device__ __forceinline__ R sifto_eval(const ProblemDev &P, const typename Vec4<R>::type *__restrict__ spl, int f, R tau) {
    const R x0 = kconst<R>(P, 0), dx = kconst<R>(P, 1), inv_dx = kconst<R>(P, 2);
    const R xn = x0 + dx * (R)P.spl_nint;
    if (!(tau >= x0 && tau <= xn)) return (R)0;
    int i = (int)floor((tau - x0) * inv_dx);
    i = min(max(i, 0), P.spl_nint - 1);
    R h = tau - (x0 + dx * (R)i);
    const typename Vec4<R>::type c = spl[f * P.spl_nint + i];
    R v = ((c.x * h + c.y) * h + c.z) * h + c.w;
    return (v != v) ? (R)0 : v;
}

// exponential of the front end: FP32 builds with LCF_FE_POLY keep it off the XU pipe (arguments clamped to the polynomial's range:
// 2^-126 stands for an underflow to zero, 2^126 for an overflow whose only use is exp(-huge)); a NaN argument must be
// handled by the caller (the polynomial does not propagate it)
// FP64 builds: libm's exp2 / log2 are ~40- and ~60-instruction dependent chains with fix-up branches, and four to six of them per
// (walker, point) were a third of the FP64 kernel's warp time (ncu source page, profiles/round2_ncu_fp64.txt: 32 % of the samples in the
// 537 instructions of the per-tile code).  The front end therefore uses the inner loop's own 2^x (the shared-memory table 2^(j/4096) and a
// degree-2 polynomial, truncation 1.0e-13: 7 FP64 operations) and a lean log2 (atanh series on the mantissa, truncation 1e-15, its
// division by MUFU.RCP64H + two Newton steps); anything unusual (|x| >= 1000 or NaN; a non-positive, subnormal or non-finite argument
// of the logarithm) still goes to libm, so the special values are libm's.  LCF_FE_LIBM = 1 restores the libm calls.
#ifndef LCF_FE_LIBM
#define LCF_FE_LIBM 0
#endif
__device__ __forceinline__ double ex2_tab_f64(double x, const double *__restrict__ e2t) {   // |x| < 1000
    // (explicit roundings: x is usually a product, and whether the compiler contracts it into these two additions depended on the
    //  kernel the front end was inlined into -- k_ring and k_pass then differed in the last bit of a log-posterior)
    const double magic = 6755399441055744.0 / (double)kE2TabSize;
    const double r = __dadd_rn(x, magic);
    const int k = __double2loint(r);                                     // 2^bits n + j (two's complement: j >= 0 for either sign of x)
    const double f = __dsub_rn(x, __dsub_rn(r, magic));
#if LCF_E2T_BITS >= 12
    double p = 2.40226506959100712e-01;
    p = fma(p, f, 6.93147180559945309e-01);
#else
    double p = 9.61812910762847716e-03;
    p = fma(p, f, 5.55041086648215800e-02);
    p = fma(p, f, 2.40226506959100712e-01);
    p = fma(p, f, 6.93147180559945309e-01);
#endif
    const double t = e2t[k & (kE2TabSize - 1)];
    p = fma(t * f, p, t);
    return __hiloint2double(__double2hiint(p) + ((k >> kE2TabBits) << 20), __double2loint(p));
}
__device__ __forceinline__ double log2_lean(double x) {                  // x positive, finite, normal
    const int hi = __double2hiint(x);
    const bool up = (hi & 0x000fffff) > 0x0006a09e;                      // mantissa above sqrt 2 (to 2^-20: the series only needs |s| small)
    const int e = (hi >> 20) - 1023 + (up ? 1 : 0);
    const double m = __hiloint2double((hi & 0x000fffff) | (up ? 0x3fe00000 : 0x3ff00000), __double2loint(x));   // [0.7071, 1.4143)
    const double s = (m - 1.0) * rcp_f64(m + 1.0), s2 = s * s;            // |s| <= 0.1716
    double p = 1. / 19.;
    p = fma(p, s2, 1. / 17.);
    p = fma(p, s2, 1. / 15.);
    p = fma(p, s2, 1. / 13.);
    p = fma(p, s2, 1. / 11.);
    p = fma(p, s2, 1. / 9.);
    p = fma(p, s2, 1. / 7.);
    p = fma(p, s2, 1. / 5.);
    p = fma(p, s2, 1. / 3.);
    p = fma(p, s2, 1.);
    return fma(2.8853900817779268 * s, p, (double)e);                   // e + (2 / ln 2) atanh(s)
}
template <typename R> __device__ __forceinline__ R fe_ex2(R x, const double *) { return Mth<R>::ex2(x); }
template <typename R> __device__ __forceinline__ R fe_lg2(R x) { return Mth<R>::lg2(x); }
#if !LCF_FE_LIBM
template <> __device__ __forceinline__ double fe_ex2<double>(double x, const double *e2t) {
    return (fabs(x) < 1000.) ? ex2_tab_f64(x, e2t) : exp2(x);
}
template <> __device__ __forceinline__ double fe_lg2<double>(double x) {
    const int hi = __double2hiint(x);
    return (hi >= 0x00100000 && hi < 0x7ff00000) ? log2_lean(x) : log2(x);
}
#endif
#if LCF_FE_POLY
template <> __device__ __forceinline__ float fe_ex2<float>(float x, const double *) { return ex2_fma(fminf(fmaxf(x, -126.f), 126.f)); }
#endif

// Front end of one (walker, point): the model value is
//      y = amp * S(invT) + add                      (ShockCooling4: min(amp S(invT), amp 0.74^-4 S(invT/0.74)))
// with S the Planck x transmission sum of the point's filter.  invT = 0 encodes "no blackbody": y = amp + add, where amp
// is 0 (before the explosion) or NaN (to propagate what the reference's numpy expression would produce).
// Written branch-free (exceptional cases are selects applied in reverse order of precedence) so that the compiler
// interleaves the MUFU chains of the unrolled items.
//   SW family (1,2,3): wc0 = K_T, wc1 = K_L, wc2 = log2(a/t_tr) | -inf, wc4 = 1/K_T, wc5 = K_L/K_T^4 (3: wc3 = E(B-V))
//        T = K_T t^eps_T ; L = K_L t^eps_L exp(-(a t/t_tr)^alpha) ; amp = L/T^4
//   SC4: wc0 = T_col_br/k_B, wc1 = c3^2 L_br, wc2 = -log2(t_br), wc3 = log2(a/t_tr), wc4/wc5 = 1/T coefficients
//   CS*: wc0, wc1 Kasen T/R^2 coefficients, wc2 Kasen factor, wc3 stretch, wc4.. r_r,r_i,r_U | dt_U,dt_i, wc8 = 1/wc0
//   SED: wc0 = 1/T, wc1 = R^2
template <int MODEL, typename R>
__device__ __forceinline__ void front_end(const ProblemDev &P, const LaneWalker<R> &w, int p, R &inv, R &amp, R &add,
                                          const typename Vec4<R>::type *__restrict__ s_spl, const double *__restrict__ e2t) {
    typedef Mth<R> M;
    add = (R)0;
    if (MODEL == 8) {
        const bool ok = w.wc[0] > (R)0;
        inv = ok ? w.wc[0] : (R)0;
        amp = ok ? w.wc[1] : w.wc[1] * (R)0;
        return;
    }
    const R dt = since<0>(P, w, p);
    const bool pos = dt > (R)0;
#ifdef LCF_X_FE_STUB   // experiment: what the kernel costs without the front-end transcendentals (results are wrong)
    inv = w.wc[4] * (R)0.5 * (dt > (R)3 ? (R)1.1 : (R)1); amp = w.wc[5]; return;
#endif
    const R lt = fe_lg2<R>(pos ? dt : (R)1);
    R i, a;
    if (MODEL >= 1 && MODEL <= 3) {
        // 4 transcendentals: lg2(t), (a t/t_tr)^alpha, L/T^4, 1/T
        const R epsT = kconst<R>(P, 0), epsA = kconst<R>(P, 1), alpha = kconst<R>(P, 2);
        const R pw_ = (w.wc[2] > -M::inf()) ? fe_ex2<R>(alpha * (lt + w.wc[2]), e2t) : (R)0;
        a = w.wc[5] * fe_ex2<R>(epsA * lt - (R)kLog2e * pw_, e2t);
        i = w.wc[4] * fe_ex2<R>(-epsT * lt, e2t);
        if (!(w.wc[4] > (R)0)) { i = (R)0; a = (w.wc[0] != w.wc[0]) ? M::nan() : w.wc[1] * (R)0; }   // T <= 0
        if (w.wc[1] < (R)0) { i = (R)0; a = M::nan(); }                       // L < 0: L ** 0.5 is NaN (models.py:268)
        if (!pos) { i = (R)0; a = (w.wc[0] * w.wc[1]) * (R)0; }                // t <= t_exp: zero (NaN constants propagate)
    } else if (MODEL == 4) {
        // 6 transcendentals: lg2(t), suppression (2), two powers of ttilde for L, one for 1/T (branch selected)
        const R A = kconst<R>(P, 0), alpha = kconst<R>(P, 1);
        const R ltt = lt + w.wc[2];                              // log2(ttilde); NaN when t_br is invalid
        const R sup = (w.wc[3] > -M::inf()) ? fe_ex2<R>((R)(-kLog2e) * fe_ex2<R>(alpha * (lt + w.wc[3]), e2t), e2t) : (R)1;
        const R L = w.wc[1] * (fe_ex2<R>((R)(-4. / 3.) * ltt, e2t) + A * sup * fe_ex2<R>((R)(-0.17) * ltt, e2t));
        // T = T_br min(0.97 u^-1/3, u^-0.45): the first branch is the smaller one for log2(u) < -log2(0.97)/(0.45-1/3)
        const bool early = ltt < (R)0.37665701296944757;
        i = (early ? w.wc[4] : w.wc[5]) * fe_ex2<R>((early ? (R)(1. / 3.) : (R)0.45) * ltt, e2t);
        if (LCF_FE_POLY && sizeof(R) == 4) i = (w.wc[2] != w.wc[2]) ? M::nan() : i;   // invalid t_br: the polynomial does not propagate the NaN
        const R i2 = i * i;
        a = L * (i2 * i2);
        if (!(i > (R)0)) { a = (i != i) ? M::nan() : L * (R)0; i = (R)0; }
        if (L < (R)0) { a = M::nan(); i = (R)0; }
        if (!pos) { a = (w.wc[0] * w.wc[1]) * (R)0; i = (R)0; }
    } else {
        // Kasen, models.py:752-754: 3 transcendentals, then the SiFTO template and the per-filter factors
        i = w.wc[8] * fe_ex2<R>((R)(74. / 144.) * lt, e2t);
        a = w.wc[1] * fe_ex2<R>((R)(14. / 9.) * lt, e2t);               // = R^2 here
        if (!(w.wc[8] > (R)0)) { i = (R)0; a = (w.wc[0] != w.wc[0]) ? M::nan() : (R)0; }
        if (!pos) { i = (R)0; a = (R)0; }
        const int f = P.pfilt[p];
        const int role = P.frole[f];
        const R tw = since<1>(P, w, p);                          // t_wrt_peak, models.py:816
        if (MODEL == 5) {
            const R ys = sifto_eval<R>(P, s_spl, f, tw * w.wc[9]);       // wc9 = 1 / stretch
            a *= (role & 1) ? w.wc[6] : (R)1;
            add = ys * ((role & 2) ? w.wc[4] : ((role & 4) ? w.wc[5] : (R)1));   // models.py:915
        } else {
            const R dtf = (role & 8) ? w.wc[4] : ((role & 16) ? w.wc[5] : (R)0);
            add = sifto_eval<R>(P, s_spl, f, (tw - dtf) * w.wc[9]);
            a *= w.wc[2];                                        // models.py:979, 1044
        }
    }
    inv = i;
    amp = a;
}

template <typename R> struct PointFE {
    R invT, amp;
    int state;
};

// Planck x transmission sums of up to two points of one filter for one walker, over the pair records [k_lo, k_lo + k_cnt) of
// the filter (the whole filter, or one chunk of it when a tile is split over several lanes).  S0 / S1: sums at the points'
// temperatures; S0s / S1s (ShockCooling4 only): at 0.74 T.  Lanes that share a (walker, point pair) take the same path, because
// the guards only look at the filter's exponent range and the two temperatures.
template <int MODEL, typename R, int TS>
__device__ __forceinline__ void blackbody_sums(const typename Vec4<R>::type *bank, const int4 fi, int k_lo, int k_cnt,
                                               const PointFE<R> &f0, const PointFE<R> &f1, bool n0, bool n1,
                                               const typename Vec2<R>::type *tab, int ts_rt, const double *e2t, R &S0, R &S0s, R &S1, R &S1s) {
    typedef typename Vec2<R>::type R2;
    typedef typename Vec4<R>::type R4;
    const int ts = TS > 0 ? TS : ts_rt;
    const int k0 = fi.x + k_lo, K2 = k_cnt;
    const R4 *b = bank + k0;
    const R2 *tb = tab + (size_t)k0 * ts;
    const R c74 = (R)(1. / 0.74);
    S0 = S0s = S1 = S1s = (R)0;
    if ((!n0 && !n1) || K2 <= 0) return;
    if (sizeof(R) == 4) {
        const float2 rng = make_float2(__int_as_float(fi.z), __int_as_float(fi.w));   // (a_min, a_max) of the filter
        const float i0 = n0 ? (float)f0.invT : (float)f1.invT, i1 = n1 ? (float)f1.invT : i0;
        const float imin = fminf(i0, i1);
        const float xsum = rng.y * (i0 + i1) * (MODEL == 4 ? (float)(1. + 1. / 0.74) : 2.f);   // sum of the 4 exponents
        if (rng.x * imin >= 0.0625f) {                      // no cancellation in 2^x - 1 anywhere in the filter
            const bool clamp = xsum > 124.f;                 // the product of four denominators could overflow: Wien form
#ifdef LCF_X_PATHCOUNT
            atomicAdd(&g_phase_clk[clamp ? 7 : 6], 1ull);        // lane-tiles on the plain / clamped fast path
#endif
            const float4 *bf = reinterpret_cast<const float4 *>(b);
            const float2 *tf = reinterpret_cast<const float2 *>(tb);
            float s0, s1, s0s = 0.f, s1s = 0.f;
            if (MODEL == 4) {
                if (clamp) planck_quad_sc4_f32<true>(bf, K2, i0, i1, s0, s0s, s1, s1s);
                else planck_quad_sc4_f32<false>(bf, K2, i0, i1, s0, s0s, s1, s1s);
            } else if (MODEL == 3) {
                if (clamp) planck_quad_f32<true, true, TS>(bf, K2, i0, i1, tf, ts, s0, s1);
                else planck_quad_f32<true, false, TS>(bf, K2, i0, i1, tf, ts, s0, s1);
            } else {
                if (clamp) planck_quad_f32<false, true, TS>(bf, K2, i0, i1, nullptr, ts, s0, s1);
                else planck_quad_f32<false, false, TS>(bf, K2, i0, i1, nullptr, ts, s0, s1);
            }
            S0 = (R)s0; S1 = (R)s1; S0s = (R)s0s; S1s = (R)s1s;
            return;
        }
    }
    if (sizeof(R) == 8) {
        const double amin = (double)__int_as_float(fi.z) * (1. - 1e-6);      // float copy of the filter's smallest a, rounded down
        double i0 = n0 ? (double)f0.invT : (double)f1.invT, i1 = n1 ? (double)f1.invT : i0;
        if (amin * fmin(i0, i1) >= 0.0009765625) {
            // exponents capped at 250 per POINT (h nu / k T = 173 at the filter's bluest sample: every sample of such a point is
            // below 2^-190 of its weight), so four denominators multiply to < 2^1000
            const double icap = 250. / ((double)__int_as_float(fi.w) * (MODEL == 4 ? (double)c74 : 1.) * (1. + 1e-6));
            i0 = fmin(i0, icap);
            i1 = fmin(i1, icap);
            const double4 *bd = reinterpret_cast<const double4 *>(b);
            const double2 *td = reinterpret_cast<const double2 *>(tb);
            double s0, s1, s0s = 0., s1s = 0.;
            if (MODEL == 3) planck_quad_f64<true>(bd, K2, i0, i1, td, ts, e2t, s0, s1);
            else planck_quad_f64<false>(bd, K2, i0, i1, nullptr, ts, e2t, s0, s1);
            if (MODEL == 4) planck_quad_f64<false>(bd, K2, i0 * (double)c74, i1 * (double)c74, nullptr, ts, e2t, s0s, s1s);
            S0 = (R)s0; S1 = (R)s1; S0s = (R)s0s; S1s = (R)s1s;
            return;
        }
    }
    // careful path (cancellation in 2^x - 1: Rayleigh-Jeans regime, either precision)
#ifdef LCF_X_PATHCOUNT
    atomicAdd(&g_phase_clk[5], 1ull << 40);                   // lane-tiles on the careful path (upper bits of slot 5)
#endif
    if (n0) {
        S0 = (MODEL == 3) ? planck_sum_safe<R, true>(b, K2, f0.invT, tb, ts) : planck_sum_safe<R, false>(b, K2, f0.invT, nullptr, 0);
        if (MODEL == 4) S0s = planck_sum_safe<R, false>(b, K2, f0.invT * c74, nullptr, 0);
    }
    if (n1) {
        S1 = (MODEL == 3) ? planck_sum_safe<R, true>(b, K2, f1.invT, tb, ts) : planck_sum_safe<R, false>(b, K2, f1.invT, nullptr, 0);
        if (MODEL == 4) S1s = planck_sum_safe<R, false>(b, K2, f1.invT * c74, nullptr, 0);
    }
}
// model value of a point from its sums (ShockCooling4: min over the two blackbodies, models.py:629-631)
template <int MODEL, typename R>
__device__ __forceinline__ R blackbody_value(const PointFE<R> &f, bool n, R S, R Ss) {
    if (!n) return f.amp;
    const R c74_4 = (R)(1. / (0.74 * 0.74 * 0.74 * 0.74));
    if (MODEL == 4) return Mth<R>::mn(f.amp * S, f.amp * c74_4 * Ss);
    return f.amp * S;
}

// Planck x transmission sums of up to FOUR points of one filter for one walker (FP32, one-blackbody models, whole filter): the
// same guards as blackbody_sums on the four temperatures (inactive points borrow an active one's and their sums are dropped).
template <int MODEL, int TS>
__device__ __forceinline__ void blackbody_sums4(const float4 *bank, const int4 fi, const PointFE<float> (&f)[4], const bool (&n)[4],
                                                const float2 *tab, int ts_rt, float (&S)[4]) {
    const int ts = TS > 0 ? TS : ts_rt;
    const float4 *b = bank + fi.x;
    const float2 *tb = tab + (size_t)fi.x * ts;
    const int K2 = fi.y;
    S[0] = S[1] = S[2] = S[3] = 0.f;
    if (!(n[0] || n[1] || n[2] || n[3]) || K2 <= 0) return;
    const float iref = n[0] ? f[0].invT : (n[1] ? f[1].invT : (n[2] ? f[2].invT : f[3].invT));
    const float i0 = n[0] ? f[0].invT : iref, i1 = n[1] ? f[1].invT : iref, i2 = n[2] ? f[2].invT : iref, i3 = n[3] ? f[3].invT : iref;
    const float2 rng = make_float2(__int_as_float(fi.z), __int_as_float(fi.w));   // (a_min, a_max) of the filter
    const float imin = fminf(fminf(i0, i1), fminf(i2, i3));
    if (rng.x * imin >= 0.0625f) {                       // no cancellation in 2^x - 1 anywhere in the filter
        const bool clamp = rng.y * ((i0 + i1) + (i2 + i3)) > 124.f;   // the product of four denominators could overflow: Wien form
        if (MODEL == 3) {
            if (clamp) planck_quad4_f32<true, true, TS>(b, K2, i0, i1, i2, i3, tb, ts, S[0], S[1], S[2], S[3]);
            else planck_quad4_f32<true, false, TS>(b, K2, i0, i1, i2, i3, tb, ts, S[0], S[1], S[2], S[3]);
        } else {
            if (clamp) planck_quad4_f32<false, true, TS>(b, K2, i0, i1, i2, i3, nullptr, ts, S[0], S[1], S[2], S[3]);
            else planck_quad4_f32<false, false, TS>(b, K2, i0, i1, i2, i3, nullptr, ts, S[0], S[1], S[2], S[3]);
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)                          // careful path (Rayleigh-Jeans regime)
        if (n[k]) S[k] = (MODEL == 3) ? planck_sum_safe<float, true>(b, K2, f[k].invT, tb, ts) : planck_sum_safe<float, false>(b, K2, f[k].invT, nullptr, 0);
}

// ---------------------------------------------------------------------------------------
// system-scope flag helpers for the fused multi-GPU exchange
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// wait until every peer has published `epoch` completed half-steps (bounded: ~4 s, then flag the error and go on)
__device__ __forceinline__ void peers_wait(const MoveDev &Mv) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int p = 0; p < Mv.npeers; ++p) {
            while ((int)(ld_acquire_sys(Mv.flags + Mv.peer_rank[p]) - Mv.epoch) < 0) {
                if (clock64() - t0 > 8000000000LL) { atomicExch(Mv.xstatus, 1); break; }
                __nanosleep(100);
            }
        }
    }
    __syncthreads();
}
// all stores of this CTA are out: count it; the last CTA of the launch tells the peers
__device__ __forceinline__ void peers_publish(const MoveDev &Mv) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(Mv.done_count, 1u);
        if (prev == gridDim.x - 1) {
            *Mv.done_count = 0u;
            __threadfence_system();
            for (int p = 0; p < Mv.npeers; ++p) st_release_sys(Mv.peer_flags[p] + Mv.myrank, Mv.epoch + 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------
// thread-block cluster helpers (a plain launch is a 1-CTA cluster: rank 0 of 1)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store a double into the same shared-memory location of CTA `rank` of this cluster (distributed shared memory)
__device__ __forceinline__ void dsmem_store_f64(double *local, uint32_t rank, double v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote), "d"(v) : "memory");
}

// ---------------------------------------------------------------------------------------
// shared-memory carve-up (dynamic), identical for the half-step and the chain kernels
// ---------------------------------------------------------------------------------------
constexpr int kMaxCluster = 8;      // portable cluster size limit
constexpr int kTermStride = kMaxTerms + kMaxDim + 2;   // per walker: model terms, prior terms, ln z, ln u

template <typename R> struct SmemLayout : SmemOffsets {
    __host__ __device__ explicit SmemLayout(const SmemOffsets &o) : SmemOffsets(o) {}
    // ncluster: largest cluster that may share a walker group (1 in the chain kernel: no cluster partials to hold)
    __host__ __device__ SmemLayout(int nsamples, int nfilters, int wpb, int nwarps, int ndim, bool tab, int nspl, int ncluster = kMaxCluster, int nq = 1) {
        size_t o = 0;
        auto at = [&](size_t bytes, bool align) { const unsigned int here = (unsigned int)o; o += bytes; if (align) o = (o + 15) & ~(size_t)15; return here; };
        off_e2t = at(sizeof(R) == 8 ? (size_t)kE2TabSize * sizeof(double) : 0, false);   // 2^(j/4096), FP64 loop
        off_bank = at((size_t)nsamples * 2 * sizeof(R), true);
        off_spl = at((size_t)nspl * 4 * sizeof(R), true);                                 // SiFTO cubic coefficients
        off_tab = at(tab ? (size_t)nsamples * (wpb < 32 ? wpb : 32) * sizeof(R) : 0, true);   // R2[nsamples/2][min(wpb, 32)]
        off_foff = at((size_t)nfilters * sizeof(int4), true);
        off_wc = at((size_t)wpb * kNumWC * sizeof(R), true);
        off_t = at((size_t)wpb * 2 * sizeof(double), false);
        off_q = at((size_t)wpb * ndim * sizeof(double), false);
        off_lp = at((size_t)wpb * sizeof(double), false);
        off_z = at((size_t)wpb * sizeof(double), false);
        off_old = at((size_t)wpb * sizeof(double), false);                               // log-probability of the walker being moved (prefetched)
        off_term = at((size_t)wpb * kTermStride * sizeof(double), false);
        off_part = at(nq > 1 ? (size_t)nq * nwarps * wpb * sizeof(R) : (size_t)nwarps * wpb * sizeof(double), true);
        off_cpart = at(ncluster > 1 ? (size_t)ncluster * wpb * sizeof(double) : 0, false);
        off_flag = at((size_t)wpb * sizeof(int), true);
        off_bar = at(16, false);
        total = (unsigned int)o;
    }
};

// ---------------------------------------------------------------------------------------
// One half-step (or one evaluation pass) for the walker group `g` of the active set.
// Called by every thread of the CTA.  A cluster of `csize` CTAs shares the group: CTA `crank` takes every csize-th
// row of tiles of the light curve and rank 0 reduces the chi-square partials through DSMEM.
// (Measured alternatives, see DESIGN.md: a chunked front-end pass through shared memory with one CTA barrier per chunk,
// and dedicated front-end producer warps feeding the loop warps through an mbarrier ring, were both slower: the XU
// pipe is fed best by 32 resident warps that ALL spend their time in the dense inner loop.)
// ---------------------------------------------------------------------------------------
// WL >= 0: walkers-per-CTA exponent known at compile time (k_pass instantiates WL = 5, the shape of every large ensemble:
// the lane arithmetic slot / ppt / wl and the tile guards fold to constants); WL = -1: read from Mv.wpb_log2.
// PLAIN: the launch is a move / log-posterior / log-likelihood pass without the intrinsic-scatter term (the reference's default):
// the per-tile mode and use_sigma branches are compiled out.
// PPL: photometry points per lane and tile (2; 4 in the 32-walker plain FP32 kernels of the one-blackbody models, see planck_quad4_f32).
// SEG: the filter bank does not fit in shared memory (k_pass_seg): `segs` lists runs of consecutive filters, (first filter, end filter,
// first pair record, pair records), whose slices of the bank do; the group sweeps the light curve once per segment with that slice
// (and, ShockCooling3, its weight table) staged, taking the tiles of the segment's filters.  Each warp still meets its tiles in table
// order, so the chi-square sums are those of an unsegmented launch of the same shape, bit for bit.
template <int MODEL, typename R, int WL = -1, bool PLAIN = false, int PPL = 2, bool SEG = false>
__device__ __forceinline__ void group_pass(const ProblemDev &P, const TileDev &TL, const MoveDev &Mv, long long g,
                                           unsigned char *smem, const SmemLayout<R> &L, bool need_stage, int crank, int csize,
                                           int q_lo = 0, int q_hi = 1, const int4 *segs = nullptr, int nseg = 1) {
    typedef typename Vec2<R>::type R2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int wl2 = WL >= 0 ? WL : Mv.wpb_log2;
    const int wpb = 1 << wl2;
    const int D = P.ndim;
    typedef typename Vec4<R>::type R4;
    R4 *s_bank = reinterpret_cast<R4 *>(smem + L.off_bank);
    R2 *s_tab = reinterpret_cast<R2 *>(smem + L.off_tab);
    const R4 *s_spl = reinterpret_cast<const R4 *>(smem + L.off_spl);
    const double *s_e2t = reinterpret_cast<const double *>(smem + L.off_e2t);
    int4 *s_finfo = reinterpret_cast<int4 *>(smem + L.off_foff);
    R *s_wc = reinterpret_cast<R *>(smem + L.off_wc);
    double *s_t = reinterpret_cast<double *>(smem + L.off_t);
    double *s_q = reinterpret_cast<double *>(smem + L.off_q);
    double *s_lp = reinterpret_cast<double *>(smem + L.off_lp);
    double *s_z = reinterpret_cast<double *>(smem + L.off_z);
    double *s_old = reinterpret_cast<double *>(smem + L.off_old);
    double *s_term = reinterpret_cast<double *>(smem + L.off_term);
    double *s_part = reinterpret_cast<double *>(smem + L.off_part);
    double *s_cpart = reinterpret_cast<double *>(smem + L.off_cpart);
    int *s_flag = reinterpret_cast<int *>(smem + L.off_flag);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.off_bar);

    LCF_TICK_INIT;
    // ---- phase 0: stage the packed filter bank with one TMA bulk copy -------------------
    if (need_stage) {
        if (tid == 0) {
            const uint32_t bytes = SEG ? 0u : (uint32_t)((size_t)(P.nsamples >> 1) * sizeof(R4));
            const uint32_t sbytes = (MODEL >= 5 && MODEL <= 7) ? (uint32_t)((size_t)P.nfilters * P.spl_nint * sizeof(R4)) : 0u;
            mbar_expect_tx(s_bar, bytes + sbytes);
            if (!SEG) tma_bulk_g2s(s_bank, P.bank, bytes, s_bar);
            if (sbytes) tma_bulk_g2s(smem + L.off_spl, P.spl, sbytes, s_bar);
        }
        for (int i = tid; i < P.nfilters; i += blockDim.x) s_finfo[i] = P.finfo[i];
    }

    // ---- phase 1: proposal + prior + per-walker model constants (FP64) ------------------------------
    // (every CTA of a cluster repeats it for the same walkers: bitwise identical, no communication)
    // 1a: one thread per walker: stretch-move draw and proposal
    constexpr int NT = ModelTerms<MODEL>::value;
    const bool with_prior = Mv.mode == MODE_MOVE || Mv.mode == MODE_LOGPOST;
    const bool q_ready = Mv.mode == MODE_LOGPOST && !Mv.qin;   // k_ring look-ahead rounds: spec_apply left the proposals in s_q (and synchronised)
    if (tid < wpb && !q_ready) {
        const long long i = g * wpb + tid;
        if (i < Mv.Ns) {
            double q[kMaxDim];
            if (Mv.mode == MODE_MOVE) {
                const long long row = Mv.act_rows ? (long long)Mv.act_rows[i] : Mv.act_base + i;
                const long long j = (row < Mv.n0) ? 2 * row : 2 * (row - Mv.n0) + 1;   // logical walker
                double z;
                long long pr;
                if (Mv.zin) {
                    z = Mv.zin[i];
                    pr = Mv.rin[i];
                } else {
                    double ua;
                    stretch_draw(Mv.seed, Mv.ctr, j, Mv.Nc, z, pr, ua);
                    s_term[tid * kTermStride + NT + kMaxDim + 1] = ua;   // u of the accept test
                }
                const long long crow = Mv.comp_rows ? (long long)Mv.comp_rows[pr] : Mv.comp_base + pr;
                const double *s = Mv.coords + row * D, *c = Mv.coords + crow * D;
                s_old[tid] = Mv.logp[row];          // needed only by the accept test: fetched now, off the critical path
                for (int d = 0; d < D; ++d)              // q = c - (c - s) z, numpy op order, no FMA
                    q[d] = __dsub_rn(c[d], __dmul_rn(__dsub_rn(c[d], s[d]), z));
                s_z[tid] = z;
            } else {
                const int nq = (Mv.mode == MODE_MODEL) ? P.nmodel : D;
                const long long qs = Mv.qstride ? Mv.qstride : nq;
                for (int d = 0; d < nq; ++d) q[d] = Mv.qin[i * qs + d];
                for (int d = nq; d < D; ++d) q[d] = 0.;
            }
            for (int d = 0; d < D; ++d) s_q[tid * D + d] = q[d];
        }
    }
    if (!q_ready) __syncthreads();
    LCF_TICK(6);
    // 1b: one thread per (walker, term): FP64 transcendentals of the model constants, log-priors, ln z and ln u
    {
        // one KIND of term per warp per pass (lanes = walkers): a warp never diverges into different double-precision routines, and
        // with few walkers per CTA the kinds run concurrently on different warps instead of serially inside warp 0
        const int per = NT + D + 2;
        for (int k = warp; k < per; k += nw)
        for (int w1 = lane; w1 < wpb; w1 += 32) {
            if (g * wpb + w1 >= Mv.Ns) continue;
            const double *q = s_q + w1 * D;
            double *t = s_term + w1 * kTermStride;
            if (k < NT) t[k] = walker_term<MODEL>(P, q, k);
            else if (k < NT + D) { if (with_prior) t[NT + (k - NT)] = prior_logp(P.prior, k - NT, q[k - NT]); }
            else if (Mv.mode == MODE_MOVE) {
                if (k == NT + D) t[NT + kMaxDim] = log(s_z[w1]);
                else if (!Mv.luin) t[NT + kMaxDim + 1] = log(t[NT + kMaxDim + 1]);
            }
        }
    }
    __syncthreads();
    LCF_TICK(7);
    // 1c: one thread per walker: sum of the log-priors (same order as the reference's loop), model constants
    if (tid < wpb) {
        const long long i = g * wpb + tid;
        int flag = 1;                                    // 1 = skip likelihood
        double lp = 0.;
        if (i < Mv.Ns) {
            const double *q = s_q + tid * D;
            const double *t = s_term + tid * kTermStride;
            if (with_prior)
                for (int d = 0; d < D; ++d) lp += t[NT + d];
            if (!isinf(lp)) {                            // fitting.py:125: prior -inf skips the likelihood
                flag = 0;
                WalkerSetup ws;
                setup_walker<MODEL>(P, q, t, ws);
                if (P.use_sigma) ws.wc[7] = q[D - 1] * q[D - 1];
                for (int k = 0; k < kNumWC; ++k) s_wc[tid * kNumWC + k] = (R)ws.wc[k];
                s_t[tid * 2] = ws.t0;
                s_t[tid * 2 + 1] = ws.t1;
            }
        }
        s_lp[tid] = lp;
        s_flag[tid] = flag;
    }
    __syncthreads();
    LCF_TICK(0);
    if (need_stage) mbar_wait(s_bar, 0);
    LCF_TICK(1);

    const int tstride = WL == 5 ? kTabStride : (wpb < 32 ? wpb : 32);     // weight-table columns
    // ShockCooling3: per-walker reddened weights  w_k 10^(-0.4 ebv kappa_k)  (filters.py:32-33), pair layout
    auto build_tab = [&](int npairs, int pair0) {          // pair0: first pair record of the staged slice (0: the whole bank)
        const R *kap = reinterpret_cast<const R *>(P.kappa) + 2 * (size_t)pair0;
        const int n = npairs * wpb;
        for (int idx = tid; idx < n; idx += blockDim.x) {
            const int kp = idx >> wl2, wl = idx & (wpb - 1);
            const R ebv = s_wc[wl * kNumWC + 3];
            const R4 rec = s_bank[kp];
            R2 v;
            v.x = rec.z * fe_ex2<R>(-ebv * kap[2 * kp], s_e2t);
            v.y = rec.w * fe_ex2<R>(-ebv * kap[2 * kp + 1], s_e2t);
            s_tab[kp * tstride + wl] = v;
        }
        __syncthreads();
    };
    if (MODEL == 3 && !SEG) build_tab(P.nsamples >> 1, 0);
    LCF_TICK(2);

    // ---- phase 2: tiles (each lane: up to two points of the tile's filter) ----------------
    // narrow groups (wpb <= 32): lane = (walker, point slot).  Wide groups (wpb = 64..256, chain kernel): the warps form
    // wpb/32 walker columns x nw/(wpb/32) tile stripes, so ONE proposal phase serves up to 256 walkers.
    // Split-K (ks > 0, small problems): the samples of a (walker, point pair) are split over 2^ks lanes -- lane = (walker, chunk,
    // point slot) -- and the partial sums are combined with warp shuffles, so that a 100-walker ensemble or a 5-point SED epoch
    // spends its half-step in many short loops instead of a few long ones.
    const int ncol = wpb > 32 ? wpb >> 5 : 1, col = warp % ncol, stripe = warp / ncol, nstripes = nw / ncol;
    const int ks = (WL == 5 || wpb > 32) ? 0 : Mv.ks;
    const int wl = wpb > 32 ? col * 32 + lane : lane & (wpb - 1);
    const int chunk = (lane >> wl2) & ((1 << ks) - 1);
    const int slot = wpb > 32 ? 0 : lane >> (wl2 + ks), ppt = wpb > 32 ? 1 : 32 >> (wl2 + ks);
    const long long iw = g * wpb + wl;
    const bool skip = s_flag[wl] != 0;
    LaneWalker<R> lw;
#pragma unroll
    for (int k = 0; k < kNumWC; ++k) lw.wc[k] = s_wc[wl * kNumWC + k];
    lw.set_epochs(s_t[wl * 2], s_t[wl * 2 + 1], P.tref);
    const R4 *pobs = reinterpret_cast<const R4 *>(P.obs);
    const R2 *s_tabw = s_tab + wl;
    const int4 *tiles = TL.tiles;
    // units [q_lo, q_hi) of the group's tile rows (nq = 1: the whole light curve, rows dealt to the CTAs of the cluster)
    const int nq = Mv.nq > 1 ? Mv.nq : 1;
    R *s_qpart = reinterpret_cast<R *>(smem + L.off_part);
    R chi = 0;
    int seg_f0 = 0, seg_f1 = 0, seg_pair0 = 0;
    for (int sg = 0; sg < (SEG ? nseg : 1); ++sg) {
    if constexpr (SEG) {                                   // stage the bank slice of segment sg (fallback path: a plain cooperative copy)
        const int4 s = __ldg(segs + sg);
        __syncthreads();                                   // every warp is done with the previous slice
        const R4 *gb = reinterpret_cast<const R4 *>(P.bank) + s.z;
        for (int i = tid; i < s.w; i += blockDim.x) s_bank[i] = gb[i];
        seg_f0 = s.x; seg_f1 = s.y; seg_pair0 = s.z;
        __syncthreads();
        if (MODEL == 3) build_tab(s.w, s.z);
    }
    for (int q = q_lo; q < q_hi; ++q) {
    if (nq > 1) chi = 0;
    for (int tile = (q * csize + crank) * nstripes + stripe; tile < TL.ntiles; tile += nstripes * csize * nq) {
        const int4 tl = __ldg(tiles + tile);                 // (first point, count, filter, -)
        if (SEG && (tl.z < seg_f0 || tl.z >= seg_f1)) continue;    // (warp-uniform)
        const bool active = !skip && slot < tl.y;
        const unsigned amask = ks ? __ballot_sync(0xffffffffu, active) : 0u;    // the chunk lanes of a point pair are active together
        if constexpr (PPL == 4 && sizeof(R) == 4) {
            // four points of the tile's filter per lane (32 walkers per CTA: lane = walker, no point slots, no split-K, plain chi-square)
            if (active) {
                typedef PointFE<float> FE;
                const float4 *pobs4 = reinterpret_cast<const float4 *>(P.obs);
                const LaneWalker<float> &lwf = reinterpret_cast<const LaneWalker<float> &>(lw);
                const int4 fi = s_finfo[tl.z];
                FE f[4];
                bool n[4];
                float2 ob[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int pk = tl.x + (k < tl.y ? k : 0);
                    const float4 o = pobs4[pk];
                    ob[k] = make_float2(o.x, o.y);
                    float addk;
                    front_end<MODEL, float>(P, lwf, pk, f[k].invT, f[k].amp, addk, reinterpret_cast<const float4 *>(s_spl), nullptr);
                    n[k] = k < tl.y && f[k].invT > 0.f;
                }
                float S[4];
                blackbody_sums4<MODEL, (WL == 5 ? kTabStride : 0)>(reinterpret_cast<const float4 *>(s_bank), fi, f, n,
                                                                    reinterpret_cast<const float2 *>(s_tabw), tstride, S);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (k < tl.y) {
                        const float yk = n[k] ? f[k].amp * S[k] : f[k].amp;
                        const float rk = (ob[k].x - yk) * ob[k].y;                  // models.py:135
                        chi = fma((R)rk, (R)rk, chi);
                    }
                }
            }
        } else
        if (active) {
            const int pa = tl.x + slot;
            const bool two = slot + ppt < tl.y;
            const int pb = two ? pa + ppt : pa;
            // observed values first: their (L1-resident) loads overlap the front end and the inner loop
            const R4 oa = pobs[pa], ob = pobs[pb];
            int4 fi = s_finfo[tl.z];
            if (SEG) fi.x -= seg_pair0;                          // pair records are addressed within the staged slice
            PointFE<R> fa, fb;
            R adda, addb;
            front_end<MODEL, R>(P, lw, pa, fa.invT, fa.amp, adda, s_spl, s_e2t);      // branch-free: the two MUFU chains interleave
            front_end<MODEL, R>(P, lw, pb, fb.invT, fb.amp, addb, s_spl, s_e2t);
            fa.state = fa.invT > (R)0 ? 1 : 0;
            fb.state = fb.invT > (R)0 ? 1 : 0;
            const bool na = fa.state == 1, nb = two && fb.state == 1;
            const int k_lo = ks ? (int)(((long long)fi.y * chunk) >> ks) : 0;
            const int k_cnt = ks ? (int)(((long long)fi.y * (chunk + 1)) >> ks) - k_lo : fi.y;
            R Sa, Sas, Sb, Sbs;
            blackbody_sums<MODEL, R, (WL == 5 ? kTabStride : 0)>(s_bank, fi, k_lo, k_cnt, fa, fb, na, nb, s_tabw, tstride, s_e2t, Sa, Sas, Sb, Sbs);
            if (ks) {
                for (int off = wpb; off < (wpb << ks); off <<= 1) {
                    Sa += __shfl_xor_sync(amask, Sa, off);
                    Sb += __shfl_xor_sync(amask, Sb, off);
                    if (MODEL == 4) { Sas += __shfl_xor_sync(amask, Sas, off); Sbs += __shfl_xor_sync(amask, Sbs, off); }
                }
            }
            R ya = blackbody_value<MODEL, R>(fa, na, Sa, Sas), yb = blackbody_value<MODEL, R>(fb, nb, Sb, Sbs);
            if (MODEL >= 5 && MODEL <= 7) { ya += adda; yb += addb; }
            if (chunk == 0) {
                if (!PLAIN && Mv.mode == MODE_MODEL) {
                    Mv.out[iw * P.npoints + pa] = (double)ya * P.scale;
                    if (two) Mv.out[iw * P.npoints + pb] = (double)yb * P.scale;
                } else if (!PLAIN && P.use_sigma) {
                    const R s2a = oa.y + lw.wc[7] * oa.z;                // models.py:130
                    const R ra = oa.x - ya;
                    chi += Mth<R>::lg2(s2a) * (R)kLn2 + ra * ra * Mth<R>::rcp(s2a);
                    if (two) {
                        const R s2b = ob.y + lw.wc[7] * ob.z;
                        const R rb = ob.x - yb;
                        chi += Mth<R>::lg2(s2b) * (R)kLn2 + rb * rb * Mth<R>::rcp(s2b);
                    }
                } else {
                    const R ra = (oa.x - ya) * oa.y;                     // models.py:135
                    chi = fma(ra, ra, chi);
                    if (two) {
                        const R rb = (ob.x - yb) * ob.y;
                        chi = fma(rb, rb, chi);
                    }
                }
            }
        }
    }
    if (nq > 1) {                                          // P(q, warp): this warp's share of unit q, per walker
        double cq = (double)chi;
        for (int off = 16; off >= wpb; off >>= 1) cq += __shfl_xor_sync(0xffffffffu, cq, off);
        if (lane < wpb) s_qpart[((q - q_lo) * nw + warp) * wpb + lane] = (R)cq;
    }
    }
    }
    LCF_TICK(3);
    if (!PLAIN && Mv.mode == MODE_MODEL) { __syncthreads(); return; }

    // ---- phase 3: reduce, accept, write back ----------------------------------------------
    const int pw = wpb < 32 ? wpb : 32;                    // walkers per warp
    if (nq == 1) {
        double chid = (double)chi;
        for (int off = 16; off >= wpb; off >>= 1) chid += __shfl_xor_sync(0xffffffffu, chid, off);
        if (lane < pw) s_part[warp * pw + lane] = chid;
    }
    __syncthreads();
    LCF_TICK(4);
    // structured sums: unit totals in warp order; a group that this CTA holds completely is finished here, a shared one by the CTA
    // that delivers the last units (arrival counter; the unit totals travel through global memory and are added in unit order)
    double tot_q = 0.;
    if (nq > 1) {
        const bool whole = q_hi - q_lo == nq;
        if (tid < wpb && g * wpb + tid < Mv.Ns) {
            for (int q = q_lo; q < q_hi; ++q) {
                double sq = 0.;
                for (int w2 = 0; w2 < nw; ++w2) sq += (double)s_qpart[((q - q_lo) * nw + w2) * wpb + tid];
                if (whole) tot_q += sq;
                else Mv.split_part[(g * nq + q) * wpb + tid] = sq;
            }
        }
        if (!whole) {
            int *s_fin = reinterpret_cast<int *>(smem + L.off_bar) + 2;
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned int n = (unsigned int)(q_hi - q_lo);
                const unsigned int prev = atomicAdd(Mv.split_tick + g, n);
                const int fin = prev + n == (unsigned int)nq;
                if (fin) { Mv.split_tick[g] = 0u; __threadfence(); }
                *s_fin = fin;
            }
            __syncthreads();
            if (!*s_fin) { __syncthreads(); return; }      // (barrier: s_fin is rewritten by the next group of this CTA)
            if (tid < wpb && g * wpb + tid < Mv.Ns)
                for (int q = 0; q < nq; ++q) tot_q += __ldcg(Mv.split_part + (g * nq + q) * wpb + tid);
        }
    }
    // partial sums of walker `tid`: over all warps (narrow) or over the stripes of its column (wide), in warp order
    auto cta_total = [&](int w) {
        double tot = 0.;
        if (wpb > 32) { for (int st = 0; st < nstripes; ++st) tot += s_part[(st * ncol + (w >> 5)) * 32 + (w & 31)]; }
        else { for (int w2 = 0; w2 < nw; ++w2) tot += s_part[w2 * wpb + w]; }
        return tot;
    };
    if (csize > 1) {                                       // partial sums of the cluster -> rank 0, in rank order
        if (tid < wpb) dsmem_store_f64(&s_cpart[crank * wpb + tid], 0u, cta_total(tid));
        cluster_sync_all();
    }
    if (crank == 0 && tid < wpb) {
        const long long i = g * wpb + tid;
        if (i < Mv.Ns) {
            double lp = s_lp[tid];
            double nlp = lp;
            if (!s_flag[tid]) {
                double tot = 0.;
                if (nq > 1) tot = tot_q;
                else if (csize > 1) { for (int r = 0; r < csize; ++r) tot += s_cpart[r * wpb + tid]; }
                else tot = cta_total(tid);
                nlp = lp + (-0.5 * (P.const_term + tot));
            }
            if (Mv.mode != MODE_MOVE) {
                Mv.out[i] = nlp;
                if (nlp != nlp) atomicAdd(Mv.nanflag, 1);
            } else {
                const long long row = Mv.act_rows ? (long long)Mv.act_rows[i] : Mv.act_base + i;
                const long long j = (row < Mv.n0) ? 2 * row : 2 * (row - Mv.n0) + 1;
                const double logu = Mv.luin ? Mv.luin[i] : s_term[tid * kTermStride + NT + kMaxDim + 1];
                if (nlp != nlp) atomicAdd(Mv.nanflag, 1);           // emcee: "Probability function returned NaN"
                const double old = s_old[tid];
                const bool acc = accept_test(D, s_term[tid * kTermStride + NT + kMaxDim], nlp, old, logu);
                double *crd = Mv.coords + row * D;
                if (acc) {
                    for (int d = 0; d < D; ++d) crd[d] = s_q[tid * D + d];
                    Mv.logp[row] = nlp;
                    if (Mv.accepted) Mv.accepted[j] += 1ull;
#pragma unroll
                    for (int p = 0; p < kMaxPeers; ++p) {            // the same update in every peer's replica (NVLink stores)
                        // (compile-time indices into Mv's peer arrays: a run-time index kept the whole MoveDev of k_chain / k_ring in local memory)
                        if (p >= Mv.npeers) break;
                        double *pc = Mv.peer_coords[p] + row * D;
                        for (int d = 0; d < D; ++d) pc[d] = s_q[tid * D + d];
                        Mv.peer_logp[p][row] = nlp;
                    }
                }
                if (Mv.chain_step) {
                    for (int d = 0; d < D; ++d) Mv.chain_step[j * D + d] = acc ? s_q[tid * D + d] : crd[d];
                    Mv.lnp_step[j] = acc ? nlp : old;
                }
            }
        }
    }
    if (csize > 1) cluster_sync_all();                     // rank 0 has read s_cpart; also orders the in-place update
    else __syncthreads();
    LCF_TICK(5);
}

// ---------------------------------------------------------------------------------------
// Kernel A: one launch = one half-step (or one evaluation pass) of ONE ensemble.
// Grid = walker groups x cluster size (cluster dimension set by the launch attribute; 1 for large ensembles).
// ---------------------------------------------------------------------------------------
template <int MODEL, typename R, int WL, bool PLAIN, int PPL = 2>
__global__ void __launch_bounds__(512, (sizeof(R) == 4 ? 2 : 1)) k_pass(const ProblemDev P, const TileDev TL, const MoveDev Mv) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wpb = 1 << (WL >= 0 ? WL : Mv.wpb_log2);
    // carve-up offsets: from the parameter block (constant bank), except in the ShockCooling4 32-walker kernels, where recomputing
    // them measured 1.6 % faster (ShockCooling3: reading them is 4.5 % faster; profiles/round2_kernel_variants.jsonl x8)
#ifndef LCF_X_LAYOUT_INKERNEL
    constexpr bool lay_in_kernel = MODEL == 4 && WL == 5 && sizeof(R) == 4;
#else
    constexpr bool lay_in_kernel = true;
#endif
    const SmemLayout<R> L = lay_in_kernel
        ? SmemLayout<R>(P.nsamples, P.nfilters, wpb, blockDim.x >> 5, P.ndim, MODEL == 3, (MODEL >= 5 && MODEL <= 7) ? P.nfilters * P.spl_nint : 0, kMaxCluster, Mv.nq)
        : SmemLayout<R>(Mv.lay);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    if (sizeof(R) == 8) stage_e2tab(reinterpret_cast<double *>(smem + L.off_e2t));
    __syncthreads();
    // Programmatic dependent launch: this grid may have been scheduled while the previous half-step was still running
    // (its launch latency and the prologue above are hidden); nothing the previous kernel wrote is read before this point.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
    if (csize > 1) cluster_sync_all();                     // every CTA of the cluster is running before its shared memory is written remotely
    const long long ngroups = (Mv.Ns + wpb - 1) / wpb;
    const long long nclusters = cluster_count_x();
    if (Mv.npeers) peers_wait(Mv);
#ifdef LCF_X_TIMING
    const unsigned long long t_cta0 = gtimer();
#endif
    bool first = true;
    // Cluster b takes the units [b U / B, (b + 1) U / B) of the (group, unit) space, U = groups x nq, B = clusters of the grid.  With
    // B = groups every cluster holds one whole group (plain sums, nq = 1: always); a flat split launches B = the co-resident CTA slots
    // of the device, so that all CTAs finish together (no partial last wave) -- a group shared by several CTAs is finished by the
    // one that delivers its last units.
    const int nq = Mv.nq > 1 ? Mv.nq : 1;
    const long long U = ngroups * nq, b = cluster_id_x();       // U < 2^31 (checked by the host)
    int u = (int)((b * U) / nclusters);
    const int u1 = (int)(((b + 1) * U) / nclusters);
    while (u < u1) {
        const int g = u / nq, qa = u - g * nq;
        const int qb = min(nq, qa + (u1 - u));
        group_pass<MODEL, R, WL, PLAIN, PPL>(P, TL, Mv, (long long)g, smem, L, first, crank, csize, qa, qb);
        first = false;
        u += qb - qa;
    }
#ifdef LCF_X_TIMING
    if (threadIdx.x == 0 && blockIdx.x < 4096) { g_cta_log[3 * blockIdx.x] = t_cta0; g_cta_log[3 * blockIdx.x + 1] = gtimer(); g_cta_log[3 * blockIdx.x + 2] = smid(); }
#endif
    if (Mv.npeers) peers_publish(Mv);
}

// ---------------------------------------------------------------------------------------
// Kernel A', the fallback of kernel A for a filter bank larger than shared memory (many densely sampled JWST / GALEX curves in one
// light curve, FP64): the bank is streamed through shared memory in segments of consecutive filters (group_pass<SEG>).  Same
// parameter block, same proposal / accept code and peer exchange as k_pass; one CTA per walker group, no cluster, plain sums.
// ---------------------------------------------------------------------------------------
struct SegDev {
    const int4 *segs;                // [nseg] (first filter, end filter, first pair record, pair records)
    int nseg;
};

template <int MODEL, typename R>
__global__ void __launch_bounds__(512, (sizeof(R) == 4 ? 2 : 1)) k_pass_seg(const ProblemDev P, const TileDev TL, const MoveDev Mv, const SegDev S) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wpb = 1 << Mv.wpb_log2;
    const SmemLayout<R> L(Mv.lay);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    if (sizeof(R) == 8) stage_e2tab(reinterpret_cast<double *>(smem + L.off_e2t));
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    const long long ngroups = (Mv.Ns + wpb - 1) / wpb;
    if (Mv.npeers) peers_wait(Mv);
    bool first = true;
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
        group_pass<MODEL, R, -1, false, 2, true>(P, TL, Mv, g, smem, L, first, 0, 1, 0, 1, S.segs, S.nseg);
        first = false;
    }
    if (Mv.npeers) peers_publish(Mv);
}

// ---------------------------------------------------------------------------------------
// Kernel B: batched independent ensembles.  One CTA per (problem, ensemble); the whole chain
// (burn-in + sampling) runs inside one launch with CTA-level barriers between half-steps.
// ---------------------------------------------------------------------------------------
struct BatchDev {
    const ProblemDev *probs;         // [nproblems]
    const TileDev *tiles;            // [nproblems]
    const int *order;                // [nproblems] CTA -> problem, by decreasing work (NULL: identity)
    double *coords;                  // [nproblems][W][D] colour-major
    double *logp;                    // [nproblems][W]
    unsigned long long *accepted;    // [nproblems][W]
    int *status;                     // [nproblems] NaN counters
    double *chain;                   // [nproblems][nsteps][W][D]
    double *lnp;                     // [nproblems][nsteps][W]
    long long W, n0, nproblems;
    long long nburn, nsteps, iter0;
    unsigned long long seed;
    int wpb_log2, init_logp, ks;
    int spec;                        // look-ahead rounds (W small: all n0 + 2 n1 virtual walkers of a step in ONE pass, see k_chain)
    double *spec_nlp;                // [nproblems][32] log-posteriors of the round's virtual walkers
    int *nan_scratch;                // NaN counter of the evaluation passes (only SELECTED evaluations count)
};

// at most 8 warps per CTA; FP32: 64 registers so that four CTAs (32 warps) share an SM, as in k_pass
#ifndef LCF_CHAIN_MINBLOCKS
#define LCF_CHAIN_MINBLOCKS 4
#endif
template <int MODEL, typename R>
__global__ void __launch_bounds__(256, (sizeof(R) == 4 ? LCF_CHAIN_MINBLOCKS : 2)) k_chain(const BatchDev B) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ ProblemDev sP;
    __shared__ TileDev sT;
    const long long prob = B.order ? B.order[blockIdx.x] : blockIdx.x;   // longest problems first: the tail of the grid is short ones
    {
        const int *src = reinterpret_cast<const int *>(B.probs + prob);
        int *dst = reinterpret_cast<int *>(&sP);
        for (int i = threadIdx.x; i < (int)(sizeof(ProblemDev) / sizeof(int)); i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) sT = B.tiles[prob];
    }
    __syncthreads();
    const int wpb = 1 << B.wpb_log2;
    SmemLayout<R> L(sP.nsamples, sP.nfilters, wpb, blockDim.x >> 5, sP.ndim, MODEL == 3, (MODEL >= 5 && MODEL <= 7) ? sP.nfilters * sP.spl_nint : 0, 1);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    if (sizeof(R) == 8) stage_e2tab(reinterpret_cast<double *>(smem + L.off_e2t));
    __syncthreads();
    const int D = sP.ndim;
    MoveDev Mv;
    Mv.coords = B.coords + prob * B.W * D;
    Mv.logp = B.logp + prob * B.W;
    unsigned long long *acc_ptr0 = B.accepted + prob * B.W;
    Mv.accepted = nullptr;
    Mv.nanflag = B.status + prob;
    Mv.W = B.W; Mv.n0 = B.n0;
    Mv.wpb_log2 = B.wpb_log2; Mv.ks = B.ks;
    Mv.nq = 1; Mv.split_part = nullptr; Mv.split_tick = nullptr;
    Mv.act_rows = nullptr; Mv.comp_rows = nullptr;
    Mv.zin = nullptr; Mv.rin = nullptr; Mv.luin = nullptr;
    Mv.npeers = 0;
    Mv.seed = B.seed ^ ((unsigned long long)(prob + 1) * 0x9E3779B97F4A7C15ull);
    Mv.qin = nullptr; Mv.out = nullptr; Mv.qstride = 0;
    bool first = true;
    const long long n1 = B.W - B.n0;
    unsigned long long *const acc_ptr = acc_ptr0;
    // Passes: the initial log-probabilities, then per step either two half-steps (MODE_MOVE) or -- look-ahead rounds, small
    // ensembles (B.spec) -- ONE log-posterior pass over the n0 + 2 n1 virtual walkers of the step: the proposals of the first colour
    // and, for every second-colour walker, its proposal against its partner's old position AND against its partner's proposal
    // (k_ring's look-ahead, with CTA barriers in place of the grid barrier and nothing deferred: the accept phases follow the
    // pass directly).  An SED epoch of 10 walkers is latency-bound like cfg1: half as many dependent passes per step.
    const bool spec = B.spec != 0;
    const int NV = (int)(B.n0 + 2 * n1), n0 = (int)B.n0;
    __shared__ double s_lz[32], s_lu[32];                  // ln z, ln u of the step's moves (by physical row)
    __shared__ int s_acc[32];                              // accept flags of the first colour
    double *s_q = reinterpret_cast<double *>(smem + L.off_q);
    double *nlp_out = B.spec_nlp + prob * 32;
    long long it = B.init_logp ? -1 : 0;
    int half = 0;
    while (it < B.nburn + B.nsteps) {
        const bool init = it < 0, round = spec && !init;
        const bool store = !init && it >= B.nburn;
        const unsigned int ctr0 = (unsigned int)(2 * (B.iter0 + (init ? 0 : it)));
        Mv.accepted = nullptr; Mv.chain_step = nullptr; Mv.lnp_step = nullptr;
        Mv.qin = nullptr; Mv.out = nullptr; Mv.nanflag = B.status + prob;
        if (init) {                                        // log-prob of the initial positions
            Mv.mode = MODE_LOGPOST;
            Mv.Ns = B.W; Mv.act_base = 0; Mv.Nc = 0; Mv.comp_base = 0;
            Mv.qin = Mv.coords; Mv.out = Mv.logp; Mv.ctr = 0;
        } else if (round) {
            // proposals: one thread per virtual walker; a second warp takes the logarithms of the W moves
            const int tid = threadIdx.x;
            if (tid < NV) {
                const int v = tid;
                int row, sel = 0;
                long long j, Nc, comp_base;
                unsigned int ctr;
                if (v < n0) { row = v; j = 2 * v; ctr = ctr0; Nc = n1; comp_base = n0; }
                else {
                    const int k = (v - n0) % (int)n1;
                    sel = (v - n0) / (int)n1;
                    row = n0 + k; j = 2 * k + 1; ctr = ctr0 + 1u; Nc = n0; comp_base = 0;
                }
                double z, u, za = 0., ua;
                long long pr, pra = 0;
                stretch_draw(Mv.seed, ctr, j, Nc, z, pr, u);
                const int prow = (int)(comp_base + pr);
                if (sel == 1) stretch_draw(Mv.seed, ctr0, 2 * prow, n1, za, pra, ua);   // the partner's own move of this step
                const double *own = Mv.coords + row * D, *cp = Mv.coords + prow * D, *cpp = Mv.coords + (n0 + pra) * D;
                for (int d = 0; d < D; ++d) {
                    double c = cp[d];
                    if (sel == 1) c = __dsub_rn(cpp[d], __dmul_rn(__dsub_rn(cpp[d], c), za));   // ... its proposal
                    s_q[v * D + d] = __dsub_rn(c, __dmul_rn(__dsub_rn(c, own[d]), z));         // q = c - (c - s) z, numpy op order, no FMA
                }
            }
            const bool two_warps = blockDim.x >= 64;
            if (two_warps ? (tid >= 32 && tid < 32 + (int)B.W) : tid < (int)B.W) {
                const int row = two_warps ? tid - 32 : tid;
                double z, u;
                long long pr;
                if (row < n0) stretch_draw(Mv.seed, ctr0, 2 * row, n1, z, pr, u);
                else stretch_draw(Mv.seed, ctr0 + 1u, 2 * (row - n0) + 1, B.n0, z, pr, u);
                s_lz[row] = log(z);
                s_lu[row] = log(u);
            }
            __syncthreads();
            Mv.mode = MODE_LOGPOST;
            Mv.Ns = NV; Mv.act_base = 0; Mv.Nc = 0; Mv.comp_base = 0;
            Mv.out = nlp_out; Mv.nanflag = B.nan_scratch; Mv.ctr = ctr0;
        } else {
            Mv.mode = MODE_MOVE;
            Mv.accepted = store ? acc_ptr : nullptr;          // acceptance counts cover the stored phase only
            Mv.chain_step = store ? B.chain + ((prob * B.nsteps + (it - B.nburn)) * B.W) * D : nullptr;
            Mv.lnp_step = store ? B.lnp + (prob * B.nsteps + (it - B.nburn)) * B.W : nullptr;
            Mv.ctr = ctr0 + (unsigned int)half;
            Mv.Ns = half ? n1 : B.n0;
            Mv.act_base = half ? B.n0 : 0;
            Mv.Nc = half ? B.n0 : n1;
            Mv.comp_base = half ? 0 : B.n0;
        }
        // two call sites, each with the mode known at compile time (one shared site with a run-time mode cost the half-step path 10-14 %)
        const long long ng = (Mv.Ns + wpb - 1) / wpb;
        if (init || round) {
            Mv.mode = MODE_LOGPOST;
            for (long long g = 0; g < ng; ++g) { group_pass<MODEL, R>(sP, sT, Mv, g, smem, L, first, 0, 1); first = false; }
        } else {
            Mv.mode = MODE_MOVE;
            for (long long g = 0; g < ng; ++g) { group_pass<MODEL, R>(sP, sT, Mv, g, smem, L, first, 0, 1); first = false; }
        }
        if (round) {
            // accept phases: first colour, then second colour with the evaluation that matches its partner's outcome
            double *cs = store ? B.chain + ((prob * B.nsteps + (it - B.nburn)) * B.W) * D : nullptr;
            double *ls = store ? B.lnp + (prob * B.nsteps + (it - B.nburn)) * B.W : nullptr;
            const int tid = threadIdx.x;
            for (int colour = 0; colour < 2; ++colour) {
                if (tid < (colour ? (int)n1 : n0)) {
                    const int row = colour ? n0 + tid : tid;
                    const long long j = colour ? 2 * tid + 1 : 2 * tid;
                    int v = row;
                    if (colour) {
                        double z, u;
                        long long pr;
                        stretch_draw(Mv.seed, ctr0 + 1u, j, B.n0, z, pr, u);
                        v = row + (s_acc[pr] ? (int)n1 : 0);
                    }
                    const double nlp = nlp_out[v], old = Mv.logp[row];
                    const bool acc = accept_test(D, s_lz[row], nlp, old, s_lu[row]);
                    if (nlp != nlp) atomicAdd(B.status + prob, 1);   // emcee: "Probability function returned NaN"
                    if (!colour) s_acc[row] = acc ? 1 : 0;
                    double *crd = Mv.coords + row * D;
                    if (acc) {
                        for (int d = 0; d < D; ++d) crd[d] = s_q[v * D + d];
                        Mv.logp[row] = nlp;
                        if (store) acc_ptr[j] += 1ull;
                    }
                    if (store) {
                        for (int d = 0; d < D; ++d) cs[j * D + d] = crd[d];
                        ls[j] = acc ? nlp : old;
                    }
                }
                __syncthreads();
            }
        }
        if (init) it = 0;
        else if (round || half == 1) { ++it; half = 0; }
        else half = 1;
    }
}

// ---------------------------------------------------------------------------------------
// Kernel C: ONE small ensemble, whole chain in one launch on a persistent co-resident grid (cooperative launch).
// A half-step of a 100-walker ensemble is ~1 us of arithmetic; launched as one k_pass per half-step it costs ~20 us of
// launch latency even with programmatic dependent launch.  Here the grid stays resident for the whole run: cluster c
// owns walker group c of every half-step (its CTAs split the light curve as in k_pass), and half-steps are separated by
// a device-side generation barrier (one atomic per CTA, release/acquire on a generation word) instead of a kernel
// boundary.  Same group_pass, same RNG keys, same order of partial sums as the k_pass launches of the same shape:
// chains are bit-identical to the per-half-step path.
// ---------------------------------------------------------------------------------------
struct RingDev {
    double *coords, *logp;
    unsigned long long *accepted;
    int *nanflag;
    double *chain, *lnp;               // [nsteps][W][D], [nsteps][W] for this run (NULL: not stored)
    long long W, n0;
    long long nsteps, iter0;
    unsigned long long seed;
    unsigned int *bar;                 // [2]: arrival counter, generation (zeroed before the launch)
    int wpb_log2, ks;
    int nq;                            // units of the structured chi-square sums (same value as the k_pass launches of this shape)
    // look-ahead rounds (spec != 0, see spec_prephase): double-buffered by round parity
    int spec;
    double *xbuf;                      // [2][W][D + 1]  positions and log-probabilities at the start of the round
    double *sq;                        // [2][NV][D]     proposals of the NV = n0 + 2 n1 virtual walkers of the round
    double *snlp;                      // [2][NV]        their log-posteriors
    double *smeta;                     // [2][W][2]      ln z and ln u of the walker's move of the round
    int *nan_scratch;                  // NaN counter of the evaluation passes (only SELECTED evaluations count, see spec_state)
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Grid barrier of the persistent kernel, split in two so that data-independent work of the next round can run in its shadow:
// every CTA adds one to a monotonic counter (red.release.gpu: its global writes are visible to whoever acquires the count) and
// round `gen` is complete when the counter reaches gen * gridDim.x -- no reset, no "last arriver" hop (an atomic's round trip,
// then a flag store, then the pollers' next load), one L2 word polled by one thread per CTA.
__device__ __forceinline__ void ring_arrive(unsigned int *bar, unsigned int &gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        gen += 1u;
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
    }
}
__device__ __forceinline__ void ring_wait(const unsigned int *bar, unsigned int gen) {
    if (threadIdx.x == 0) {
        const unsigned int target = gen * gridDim.x;
        while ((int)(ld_acquire_gpu(bar) - target) < 0) { }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Look-ahead rounds: one grid barrier per STEP instead of one per half-step.
// A half-step of a 100-walker ensemble is a ~10 us dependent chain (partner's position -> proposal -> FP64 constants -> tiles ->
// accept -> barrier) on a GPU that is 99 % idle, and the second half-step of a step can only start when the first has finished,
// because its proposals use the UPDATED positions of the first colour.  But an updated position is one of exactly two values known
// from the start of the step: the walker's old position, or its proposal.  So a round evaluates, concurrently,
//     n0 proposals of the first colour                       (virtual walkers [0, n0)), and
//     2 n1 proposals of the second colour: for BOTH outcomes of the partner's move   ([n0, n0 + n1): rejected, [n0 + n1, NV): accepted),
// stores proposals and log-posteriors, and passes ONE barrier.  Nothing else is written: at the start of the next round every CTA
// derives the positions it needs from the previous round's records (spec_state: a first-colour walker moved iff its accept test
// passed; a second-colour walker selects the evaluation that matches its partner's outcome, then takes its own test).  The owner of a
// walker (its first virtual walker) also materialises the state into xbuf, the chain row of the finished step, the acceptance count and
// the NaN flag.  1.5x the arithmetic of the plain scheme for half its serial chain; same draws, same proposals (bit for bit), same
// evaluation code (group_pass in MODE_LOGPOST on the same launch shape), same accept expression: the chain is bit-identical.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ long long spec_prev_partner(const RingDev &G, long long rd, long long row) {   // first-colour row that second-colour `row` moved against in round rd - 1
    if (rd <= 0 || row < G.n0) return -1;
    double z, u;
    long long pr;
    stretch_draw(G.seed, (unsigned int)(2 * (G.iter0 + rd - 1) + 1), 2 * (row - G.n0) + 1, G.n0, z, pr, u);
    return pr;
}
// Coordinate d and log-probability of physical row `row` at the start of round rd; acc / isnan: its move of round rd - 1.
// One thread per (walker, coordinate): every load is an independent scalar, so a state costs ONE round trip to L2 (a thread that
// walked the D coordinates of three candidate positions in loops paid one per element: 6 us of a 16 us round).
struct SpecElem { double x, lp; bool acc, isnan; };
struct SpecBase {                      // the previous round's records (32-bit offsets from here: the apply phase is one warp's
    const double *meta, *x, *nlp, *sq; //  dependent instruction stream, and 64-bit index arithmetic was most of it)
    int n0, n1, D;
    bool initial;                      // round 0: the state is the ensemble's (coords, logp)
    __device__ __forceinline__ SpecBase(const RingDev &G, long long rd, int D_) {
        const long long pb = (rd - 1) & 1, NV = G.n0 + 2 * (G.W - G.n0);
        initial = rd == 0;
        meta = G.smeta + pb * G.W * 2; x = G.xbuf + pb * G.W * (D_ + 1); nlp = G.snlp + pb * NV; sq = G.sq + pb * NV * D_;
        if (initial) { x = G.coords; nlp = G.logp; }
        n0 = (int)G.n0; n1 = (int)(G.W - G.n0); D = D_;
    }
};
// plain form (round 0 and the final state): every lane fetches everything it needs
__device__ __forceinline__ SpecElem spec_state(const SpecBase &B, int row, int prev_partner, int d) {
    SpecElem e;
    e.acc = false; e.isnan = false;
    const int D = B.D;
    if (B.initial) {                                             // (uniform)
        e.x = __ldcg(B.x + row * D + d);
        e.lp = __ldcg(B.nlp + row);
        return e;
    }
    // a first-colour row has one evaluation (v = row), a second-colour row two (v = row: partner rejected, v = row + n1: accepted)
    const bool second = row >= B.n0;
    const int vb = second ? row + B.n1 : row, a = second ? prev_partner : row;
    const double *m = B.meta + row * 2, *ma = B.meta + a * 2, *xo = B.x + row * (D + 1);
    const double lnz = __ldcg(m), logu = __ldcg(m + 1), old = __ldcg(xo + D), xold = __ldcg(xo + d);
    const double nlp0 = __ldcg(B.nlp + row), nlp1 = __ldcg(B.nlp + vb);
    const double x0 = __ldcg(B.sq + row * D + d), x1 = __ldcg(B.sq + vb * D + d);
    const double a_nlp = __ldcg(B.nlp + a), a_lnz = __ldcg(ma), a_logu = __ldcg(ma + 1), a_old = __ldcg(B.x + a * (D + 1) + D);
    const bool sel = second && accept_test(D, a_lnz, a_nlp, a_old, a_logu);   // did the first-colour partner move in the previous round?
    const double nlp = sel ? nlp1 : nlp0, xq = sel ? x1 : x0;
    e.acc = accept_test(D, lnz, nlp, old, logu);
    e.isnan = nlp != nlp;
    e.x = e.acc ? xq : xold;
    e.lp = e.acc ? nlp : old;
    return e;
}
// what the owner of `row` records when it derives the row's state at the start of round rd (rd = nsteps: the final state)
__device__ __forceinline__ void spec_commit(const RingDev &G, long long rd, int row, int d, int D, const SpecElem &e) {
    const int W = (int)G.W, n0 = (int)G.n0;
    const int j = (row < n0) ? 2 * row : 2 * (row - n0) + 1;
    if (rd < G.nsteps) {
        double *xo = G.xbuf + (rd & 1) * (long long)(W * (D + 1)) + row * (D + 1);
        xo[d] = e.x;
        if (d == 0) xo[D] = e.lp;
    } else {
        G.coords[row * D + d] = e.x;
        if (d == 0) G.logp[row] = e.lp;
    }
    if (rd > 0) {
        if (d == 0 && e.isnan) atomicAdd(G.nanflag, 1);        // emcee: "Probability function returned NaN" (selected evaluations only)
        if (G.chain) {
            G.chain[((rd - 1) * W + j) * D + d] = e.x;
            if (d == 0) {
                G.lnp[(rd - 1) * W + j] = e.lp;
                if (e.acc) atomicAdd(G.accepted + j, 1ull);      // (a reduction: nothing waits for the old count)
            }
        }
    }
}
// A virtual walker's round in two parts.  spec_plan (one thread per virtual walker, a second warp for the logarithms): everything
// that does not depend on the other walkers' outcomes -- the Philox draws of its own move, of its partner's move (second variant) and
// of the previous-round moves it will have to resolve, ln z, ln u -- computed in the shadow of the grid barrier and parked in shared
// memory (the walker's row of the term scratch, free between two passes).  spec_apply (after the barrier, one thread per (virtual
// walker, coordinate)): one round trip to L2 for the previous round's records, the selects, the proposal.
constexpr int kSpecDims = 16;            // threads per virtual walker in spec_apply (>= kMaxDim)
static_assert(kSpecDims >= kMaxDim && kMaxDim <= 13 && kTermStride >= 12, "spec_plan parks 12 words per walker in the term scratch; lanes 13..15 of a walker are the scalar lanes of spec_apply");
__device__ __forceinline__ void spec_plan(const RingDev &G, long long rd, long long g, int wpb, double *s_plan) {
    // lane = virtual walker of the group, warp = role: the up to five Philox draws of a walker's plan form chains (own draw -> partner
    // -> the partner's draw -> its partner -> that one's previous move); five warps walk them side by side (depth 3 instead of 5)
    const int tid = threadIdx.x, w = tid & 31, role = tid >> 5;
    const int nroles = blockDim.x >= 160 ? 5 : (blockDim.x >= 64 ? 2 : 1);
    if (w >= wpb || role >= nroles) return;
    const bool r_main = role == 0, r_meta = role == (nroles >= 2 ? 1 : 0);
    const bool r_part = role == (nroles == 5 ? 2 : 0), r_own = role == (nroles == 5 ? 3 : 0), r_pp = role == (nroles == 5 ? 4 : 0);
    const long long n0 = G.n0, n1 = G.W - n0, NV = n0 + 2 * n1;
    const unsigned int ctr0 = (unsigned int)(2 * (G.iter0 + rd));
    double *pd = s_plan + w * kTermStride;                       // words 0..7: eight ints (pl[0..7]); 8..11: z, za, ln z, ln u
    int *pl = reinterpret_cast<int *>(pd);
    const long long v = g * wpb + w;
    if (v >= NV) { if (r_main) pl[0] = -1; return; }
    long long row, j, Nc, comp_base;
    unsigned int ctr;
    int sel = 0;
    if (v < n0) { row = v; j = 2 * v; ctr = ctr0; Nc = n1; comp_base = n0; }
    else {
        const long long k = (v - n0) % n1;
        sel = (int)((v - n0) / n1);
        row = n0 + k; j = 2 * k + 1; ctr = ctr0 + 1u; Nc = n0; comp_base = 0;
    }
    if (r_own) pl[4] = (int)spec_prev_partner(G, rd, row);
    if (!(r_main || r_meta || r_part || r_pp)) return;
    double z, u;
    long long pr;
    stretch_draw(G.seed, ctr, j, Nc, z, pr, u);
    const long long prow = comp_base + pr;
    if (r_meta && sel == 0) { pd[10] = log(z); pd[11] = log(u); }
    if (r_main) { pl[0] = (int)v; pl[1] = (int)row; pl[2] = (int)prow; pl[7] = sel; pd[8] = z; }
    if (r_part) pl[5] = (int)spec_prev_partner(G, rd, prow);
    if (r_main || r_pp) {
        long long pprow = -1;
        double za = 0.;
        if (sel == 1) {                                          // the partner's own move of this round (it moves in the first half)
            double ua;
            long long pra;
            stretch_draw(G.seed, ctr0, 2 * prow, n1, za, pra, ua);
            pprow = n0 + pra;
        }
        if (r_main) { pl[3] = (int)pprow; pd[9] = za; }
        if (r_pp) pl[6] = sel == 1 ? (int)spec_prev_partner(G, rd, pprow) : -1;
    }
}
// s_q: the CTA's proposal rows in shared memory (group_pass reads them there when Mv.qin is NULL); ends with a CTA barrier
__device__ __forceinline__ void spec_apply(const RingDev &G, long long rd, int wpb, int D, int crank, const double *s_plan, double *s_q) {
    const int NV = (int)(G.n0 + 2 * (G.W - G.n0)), W = (int)G.W, n0 = (int)G.n0, n1 = W - n0;
#ifdef LCF_X_TIMING
    const long long tx0 = clock64();
#endif
    double *sq_out = G.sq + (rd & 1) * (long long)(NV * D), *meta_out = G.smeta + (rd & 1) * (long long)(2 * W);
    // The previous round's records live in ONE allocation (xbuf, sq, snlp, smeta): 32-bit offsets from its start.
    const double *base = G.xbuf;
    const int pb = (int)((rd - 1) & 1);
    const int xb = pb * W * (D + 1), sqb = (int)(G.sq - base) + pb * NV * D, nb = (int)(G.snlp - base) + pb * NV, mb = (int)(G.smeta - base) + pb * 2 * W;
    const unsigned int gmask = 0xFFFFu << (threadIdx.x & 16);    // the 16 lanes of this thread's walker
    const int g0 = threadIdx.x & 16;
    for (int t = threadIdx.x; t < wpb * kSpecDims; t += blockDim.x) {
        const int w = t / kSpecDims, d = t % kSpecDims;
        const double *pd = s_plan + w * kTermStride;
        const int *pl = reinterpret_cast<const int *>(pd);
        const int v = pl[0];
        if (v < 0) continue;                                     // (the whole group)
        const int sel = pl[7];
        SpecElem own;
        double c, cc;
        if (rd == 0) {                                           // (uniform) the ensemble's own arrays
            const SpecBase B(G, rd, D);
            const int dd = d < D ? d : 0;
            own = spec_state(B, pl[1], -1, dd);
            c = spec_state(B, pl[2], -1, dd).x;
            cc = spec_state(B, sel == 1 ? pl[3] : pl[2], -1, dd).x;
        } else {
            // Three states per proposal: the walker's own, its partner's, and (variant 1; variant 0 repeats the partner) the partner's
            // partner's.  A state is one of up to three stored positions, chosen by two accept tests on nine scalars.  Lanes 13..15 of
            // the walker's 16 fetch the nine scalars of one state each and take its tests; lanes d < D fetch the three candidate
            // coordinates of each state: NINE independent loads per lane whatever its role -- the same straight-line code with
            // per-lane offsets, one batch, one round trip to L2 -- then three shuffles hand the outcomes to the coordinate lanes.
            // (Each lane fetching everything was 36 loads against 64 registers: three to four round trips, 2 800 clocks.)
            const int r[3] = {pl[1], pl[2], sel == 1 ? pl[3] : pl[2]}, pp[3] = {pl[4], pl[5], sel == 1 ? pl[6] : pl[5]};
            const bool srole = d >= 13;
            const int ss = d - 13, dd = d < D ? d : 0;
            int o[9];
            {
                const int rs = ss == 0 ? r[0] : (ss == 1 ? r[1] : r[2]), ps = ss == 0 ? pp[0] : (ss == 1 ? pp[1] : pp[2]);
                const bool sec = rs >= n0;
                const int vbs = sec ? rs + n1 : rs, as = sec ? ps : rs;
                const int os[9] = {mb + 2 * rs, mb + 2 * rs + 1, xb + rs * (D + 1) + D, nb + rs, nb + vbs, nb + as, mb + 2 * as, mb + 2 * as + 1,
                                   xb + as * (D + 1) + D};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int vbk = r[k] >= n0 ? r[k] + n1 : r[k];
                    o[3 * k] = srole ? os[3 * k] : xb + r[k] * (D + 1) + dd;
                    o[3 * k + 1] = srole ? os[3 * k + 1] : sqb + r[k] * D + dd;
                    o[3 * k + 2] = srole ? os[3 * k + 2] : sqb + vbk * D + dd;
                }
            }
            double x[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) x[k] = __ldcg(base + o[k]);
            // scalar lanes: x = (ln z, ln u, old, nlp0, nlp1, partner's nlp, ln z, ln u, old)
            const int rs = ss == 0 ? r[0] : (ss == 1 ? r[1] : r[2]);
            const bool s_sel = rs >= n0 && accept_test(D, x[6], x[5], x[8], x[7]);
            const double s_nlp = s_sel ? x[4] : x[3];
            const bool s_acc = accept_test(D, x[0], s_nlp, x[2], x[1]);
            const int bits = (s_sel ? 1 : 0) | (s_acc ? 2 : 0) | (s_nlp != s_nlp ? 4 : 0);
            const double s_lp = s_acc ? s_nlp : x[2];
            const int b0 = __shfl_sync(gmask, bits, g0 + 13), b1 = __shfl_sync(gmask, bits, g0 + 14), b2 = __shfl_sync(gmask, bits, g0 + 15);
            own.lp = __shfl_sync(gmask, s_lp, g0 + 13);
            own.acc = (b0 & 2) != 0; own.isnan = (b0 & 4) != 0;
            // coordinate lanes: x = (old, evaluation 0, evaluation 1) x three states
            own.x = (b0 & 2) ? ((b0 & 1) ? x[2] : x[1]) : x[0];
            c = (b1 & 2) ? ((b1 & 1) ? x[5] : x[4]) : x[3];
            cc = (b2 & 2) ? ((b2 & 1) ? x[8] : x[7]) : x[6];
        }
        if (sel == 1) c = __dsub_rn(cc, __dmul_rn(__dsub_rn(cc, c), pd[9]));   // the partner after its own move of this round, had it been accepted
        const double q = __dsub_rn(c, __dmul_rn(__dsub_rn(c, own.x), pd[8]));   // q = c - (c - s) z, numpy op order, no FMA
#ifdef LCF_X_TIMING
        if (threadIdx.x == 0 && q == q) atomicAdd(&g_phase_clk[1], (unsigned long long)(clock64() - tx0));
#endif
        if (d >= D) continue;
        const int row = pl[1];
        s_q[w * D + d] = q;
        sq_out[v * D + d] = q;
        if (d == 0 && sel == 0) {                                // ln z, ln u of this round's move of `row`
            meta_out[row * 2] = pd[10];
            meta_out[row * 2 + 1] = pd[11];
        }
        if (sel == 0 && crank == 0) spec_commit(G, rd, row, d, D, own);
    }
#ifdef LCF_X_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_phase_clk[2], (unsigned long long)(clock64() - tx0));
#endif
    __syncthreads();
}

template <int MODEL, typename R>
__global__ void __launch_bounds__(512, (sizeof(R) == 4 ? 2 : 1)) k_ring(const ProblemDev P, const TileDev TL, const RingDev G) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int wpb = 1 << G.wpb_log2;
    SmemLayout<R> L(P.nsamples, P.nfilters, wpb, blockDim.x >> 5, P.ndim, MODEL == 3, (MODEL >= 5 && MODEL <= 7) ? P.nfilters * P.spl_nint : 0, kMaxCluster, G.nq);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + L.off_bar);
    if (threadIdx.x == 0) mbar_init(s_bar, 1);
    if (sizeof(R) == 8) stage_e2tab(reinterpret_cast<double *>(smem + L.off_e2t));
    __syncthreads();
    const int crank = (int)cluster_ctarank(), csize = (int)cluster_nctarank();
    if (csize > 1) cluster_sync_all();
    const long long cid = cluster_id_x(), nclusters = cluster_count_x();
    const int D = P.ndim;
    MoveDev Mv;
    Mv.coords = G.coords; Mv.logp = G.logp;
    Mv.nanflag = G.nanflag;
    Mv.W = G.W; Mv.n0 = G.n0;
    Mv.mode = MODE_MOVE; Mv.wpb_log2 = G.wpb_log2; Mv.ks = G.ks;
    Mv.nq = G.nq > 1 ? G.nq : 1; Mv.split_part = nullptr; Mv.split_tick = nullptr;   // every CTA holds whole groups here
    Mv.act_rows = nullptr; Mv.comp_rows = nullptr;
    Mv.zin = nullptr; Mv.rin = nullptr; Mv.luin = nullptr;
    Mv.npeers = 0;
    Mv.seed = G.seed;
    Mv.qin = nullptr; Mv.out = nullptr; Mv.qstride = 0;
    const long long n1 = G.W - G.n0;
    unsigned int gen = 0u;
    bool first = true;
    // rounds: a half-step each, or (look-ahead) a whole step each; ONE group_pass call site serves both
    const bool spec = G.spec != 0;
    const long long NV = G.n0 + 2 * n1, nrounds = spec ? G.nsteps : 2 * G.nsteps;
    Mv.accepted = nullptr; Mv.chain_step = nullptr; Mv.lnp_step = nullptr;
    Mv.Ns = 0; Mv.act_base = 0; Mv.Nc = 0; Mv.comp_base = 0; Mv.ctr = 0u;
    double *s_plan = reinterpret_cast<double *>(smem + L.off_term);       // (the term scratch is free between two passes)
    if (spec && cid < (NV + wpb - 1) / wpb) spec_plan(G, 0, cid, wpb, s_plan);
    __syncthreads();
    for (long long rd = 0; rd < nrounds; ++rd) {
        if (spec) {
            const int cb = (int)(rd & 1);
            Mv.mode = MODE_LOGPOST;
            Mv.Ns = NV;
            Mv.qin = nullptr;                                  // the proposals are already in shared memory (spec_apply)
            Mv.out = G.snlp + (long long)cb * NV;
            Mv.nanflag = G.nan_scratch;
        } else {
            const long long it = rd >> 1;
            const int half = (int)(rd & 1);
            Mv.accepted = G.chain ? G.accepted : nullptr;      // acceptance counts cover stored steps only (emcee's Backend)
            Mv.chain_step = G.chain ? G.chain + it * G.W * D : nullptr;
            Mv.lnp_step = G.chain ? G.lnp + it * G.W : nullptr;
            Mv.ctr = (unsigned int)(2 * (G.iter0 + it) + half);
            Mv.Ns = half ? n1 : G.n0;
            Mv.act_base = half ? G.n0 : 0;
            Mv.Nc = half ? G.n0 : n1;
            Mv.comp_base = half ? 0 : G.n0;
        }
        const long long ng = (Mv.Ns + wpb - 1) / wpb;
        for (long long g = cid; g < ng; g += nclusters) {      // (look-ahead: the host launches one cluster per group)
#ifdef LCF_X_TIMING
            const long long ta0 = clock64();
#endif
            if (spec) spec_apply(G, rd, wpb, D, crank, s_plan, reinterpret_cast<double *>(smem + L.off_q));
#ifdef LCF_X_TIMING
            if (threadIdx.x == 0) atomicAdd(&g_phase_clk[9], (unsigned long long)(clock64() - ta0));
#endif
            group_pass<MODEL, R>(P, TL, Mv, g, smem, L, first, crank, csize, 0, Mv.nq);
            first = false;
        }
#ifdef LCF_X_TIMING
        const long long tb0 = clock64();
#endif
        ring_arrive(G.bar, gen);
        if (spec && rd + 1 < nrounds && cid < ng) spec_plan(G, rd + 1, cid, wpb, s_plan);   // in the shadow of the barrier
        ring_wait(G.bar, gen);
#ifdef LCF_X_TIMING
        if (threadIdx.x == 0) atomicAdd(&g_phase_clk[8], (unsigned long long)(clock64() - tb0));
#endif
    }
    if (spec && crank == 0) {                               // final state, last chain row
        const long long ng = (NV + wpb - 1) / wpb;
        for (long long g = cid; g < ng; g += nclusters)
            for (int t = threadIdx.x; t < wpb * kSpecDims; t += blockDim.x) {
                const long long v = g * wpb + t / kSpecDims;
                const int d = t % kSpecDims;
                if (v < G.n0 + n1) {                                // the owners: first virtual walker of every row (v = row)
                    if (d < D) spec_commit(G, G.nsteps, (int)v, d, D, spec_state(SpecBase(G, G.nsteps, D), (int)v, (int)spec_prev_partner(G, G.nsteps, v), d));
                }
            }
    }
}

}  // namespace lcf
