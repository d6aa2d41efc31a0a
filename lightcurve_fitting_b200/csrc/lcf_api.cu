// lcf_api.cu -- host side of liblcf_b200.so: the C ABI declared in include/lcf.h.
//
// Everything here is plumbing around the kernels in lcf_device.cuh: argument checks,
// folding of physical constants / unit scales into the packed filter bank (FP64 on the
// host), device buffers, launch-shape heuristics, CUDA-event timing.  There is no CPU
// implementation of the hot path in this library.
#include "../../include/lcf.h"
#include "lcf_device.cuh"
#include "lcf_diag.cuh"

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: ranges show up in Nsight Systems / ncu --nvtx, free otherwise

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace lcf;

namespace {

struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_err;
int g_tune_wpb = 0, g_tune_nw = 0, g_tune_cluster = 0, g_tune_ks = -1;   // launch-shape overrides (0 / -1: heuristic)
int g_tune_flat = -1;                                                    // flat split of the (group, unit) space: -1 cost model, 0 off, 1 on (when the shape allows it)

int fail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(LCF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// physical constants, computed with the same libm calls as the reference's astropy expressions
// (models.py:10-12, 1101-1102): CODATA 2018 / IAU 2015.
struct Consts {
    double kB, c3, c4;
    Consts() {
        const double h = 6.62607015e-34, k = 1.380649e-23, c = 299792458.0, e = 1.602176634e-19;
        const double sigma = 2. * std::pow(M_PI, 5) * std::pow(k, 4) / (15. * std::pow(h, 3) * std::pow(c, 2));
        const double Rsun = 6.957e8, au = 1.495978707e11;
        const double Mpc = 1e6 * (au * 648000. / M_PI);
        kB = k / e * 1e3;
        c3 = std::pow(4. * M_PI * (sigma * 1e7 * std::pow(Rsun, 2) * 1e12), -0.5) / 1000.;
        c4 = 1. / (4. * M_PI * std::pow(Mpc, 2.));
    }
};
const Consts &consts() {
    static Consts c;
    return c;
}

template <typename T> int upload(const std::vector<T> &h, void **d) {
    *d = nullptr;
    size_t bytes = std::max<size_t>(h.size() * sizeof(T), 16);
    CUDA_TRY(cudaMalloc(d, bytes));
    if (!h.empty()) CUDA_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// All device arrays of one problem live in ONE allocation filled by ONE copy: survey batches create 10^4 problems, and ten
// cudaMalloc + cudaMemcpy pairs each were most of the host time per light curve.
struct Arena {
    std::vector<unsigned char> host;
    std::vector<std::pair<const void **, size_t>> slots;        // pointer field to patch, offset
    template <typename T> void add(const std::vector<T> &v, const void **slot) {
        size_t off = (host.size() + 255) & ~(size_t)255;
        host.resize(off + std::max<size_t>(v.size() * sizeof(T), 16));
        if (!v.empty()) memcpy(host.data() + off, v.data(), v.size() * sizeof(T));
        slots.push_back({slot, off});
    }
    int commit(std::vector<void *> &allocs) {
        void *d = nullptr;
        CUDA_TRY(cudaMalloc(&d, std::max<size_t>(host.size(), 16)));
        allocs.push_back(d);
        CUDA_TRY(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
        patch(d);
        return 0;
    }
    void patch(const void *base) {
        for (auto &s : slots) *s.first = reinterpret_cast<const unsigned char *>(base) + s.second;
    }
};

int model_nparams(int model) {
    switch (model) {
        case 1: return 5;
        case 2: return 4;
        case 3: return 7;
        case 4: return 5;
        case 5: return 8;
        case 6: return 7;
        case 7: return 7;
        case 8: return 2;
        default: return -1;
    }
}

}  // namespace

// -----------------------------------------------------------------------------------------
// handles
// -----------------------------------------------------------------------------------------
struct lcf_problem {
    ProblemDev dev;
    int precision = 0;
    std::vector<void *> allocs;                 // every device allocation, freed on destroy
    std::vector<int> h_point_filter;            // grouped by filter
    std::map<int, TileDev> tile_tabs;           // key = lanes-per-point exponent * 1024 + dealing period (0: natural order), see get_tiles
    std::vector<int> h_filter_records;          // pair records of every filter (tile cost)
    struct { long long Ns = -1; int l = 0, nw = 0, cluster = 1, tune = -1, ks = 0, nq = 1, seg = 0; long long flat = 0; size_t smem = 0; double cost = 0.; } shape_cache;
    int4 *d_segs = nullptr;                     // bank segments of a bank larger than shared memory (build_segments), else NULL
    int nseg = 0, seg_samples = 0, seg_wpb = 0;  // segments, samples of the largest one, walkers per CTA they were sized for
    double mean_samples = 0.;                   // mean transmission samples per photometry point
    struct { int wpb = 0, nw = 0, cluster = 0, variant = -1, ks = 0, nq = 1, ppl = 2; long long grid = 0, groups = 0; } last_launch;   // lcf_problem_last_launch
    double *d_eval_q = nullptr, *d_eval_out = nullptr;   // evaluation scratch, grow-only (no cudaMalloc / cudaFree per call)
    int *d_eval_nan = nullptr;
    size_t eval_q_cap = 0, eval_out_cap = 0;
    double *d_split_part = nullptr;             // flat-split scratch of the evaluation passes (ensembles own theirs)
    unsigned int *d_split_tick = nullptr;
    long long split_cap = 0;
    int device = 0;
    Arena *pending = nullptr;                   // device arrays not uploaded yet (problems of a batch share ONE allocation and copy)
    ~lcf_problem() {
        delete pending;
        for (void *p : allocs) cudaFree(p);
        cudaFree(d_eval_q); cudaFree(d_eval_out); cudaFree(d_eval_nan); cudaFree(d_split_part); cudaFree(d_split_tick);
    }
};

struct lcf_ensemble {
    lcf_problem *p = nullptr;
    long long W = 0, n0 = 0, n1 = 0;
    int D = 0, rank = 0, world = 1;
    long long own_begin[2] = {0, 0}, own_count[2] = {0, 0};
    unsigned long long seed = 0;
    long long iteration = 0;                    // RNG counter, never reset
    double *d_coords = nullptr, *d_logp = nullptr;
    unsigned long long *d_acc = nullptr;
    int *d_nan = nullptr;
    double *d_chain = nullptr, *d_lnp = nullptr;      // [cap][cw][D], [cap][cw]: only the walkers this rank owns
    long long cfirst = 0, cw = 0;                      // stored logical walkers [cfirst, cfirst + cw)
    double *d_stage = nullptr;                         // [W][D] + [W]: set_state staging, kept (with peer access on, every
    int *d_stage_flag = nullptr;                       // cudaMalloc / cudaFree maps into the peers too: slow and erratic)
    long long cap = 0, nstored = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_step = nullptr;
    bool has_state = false;
    double last_ms = 0.;
    long long last_launches = 0;
    // fused multi-GPU exchange (lcf_ensemble_peers_*)
    unsigned int *d_flags = nullptr;            // [kMaxPeers + 1] published half-step counters + [kMaxPeers + 1] done counter
    int npeers = 0;
    double *peer_coords[kMaxPeers] = {nullptr}, *peer_logp[kMaxPeers] = {nullptr};
    unsigned int *peer_flags[kMaxPeers] = {nullptr};
    int peer_rank[kMaxPeers] = {0};
    std::vector<void *> ipc_opened;             // mappings to close
    unsigned int epoch = 0;                     // fused half-steps launched so far (identical on every rank)
    unsigned int *d_ring_bar = nullptr;         // k_ring: arrival counter + generation word
    double *d_split_part = nullptr;             // flat split: unit sums of walker groups shared by several CTAs, [kSplitUnits][W + 32]
    unsigned int *d_split_tick = nullptr;       // [W + 1] arrival counters (zero between launches)
    int ring_ok = -1, ring_key = -1;            // -1 unknown, 0 the grid does not fit, 1 usable (cached per launch shape)
    int ring_spec = 0;                          // the cached decision covers the look-ahead grid (n0 + 2 n1 virtual walkers)
    double *d_spec = nullptr;                   // look-ahead rounds of k_ring: xbuf, sq, snlp, smeta (double-buffered) + a NaN counter
    ~lcf_ensemble() {
        for (void *m : ipc_opened) cudaIpcCloseMemHandle(m);
        cudaFree(d_flags); cudaFree(d_ring_bar); cudaFree(d_split_part); cudaFree(d_split_tick); cudaFree(d_spec);
        cudaFree(d_stage); cudaFree(d_stage_flag);
        cudaFree(d_coords); cudaFree(d_logp); cudaFree(d_acc); cudaFree(d_nan); cudaFree(d_chain); cudaFree(d_lnp);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_step) cudaEventDestroy(ev_step);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream && own_stream) cudaStreamDestroy(stream);
    }
};

struct lcf_batch {
    std::vector<lcf_problem *> probs;
    long long W = 0, n0 = 0, nprob = 0, nsteps_stored = 0, iteration = 0;
    int D = 0, model = 0, precision = 0, wpb_log2 = 0, nw = 0, ks = 0;
    size_t smem = 0;
    unsigned long long seed = 0;
    ProblemDev *d_probs = nullptr;
    TileDev *d_tiles = nullptr;
    int *d_order = nullptr;
    double *d_coords = nullptr, *d_logp = nullptr, *d_chain = nullptr, *d_lnp = nullptr;
    unsigned long long *d_acc = nullptr;
    int *d_status = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool has_state = false, need_init_logp = true;
    double last_ms = 0.;
    long long last_launches = 0;
    std::vector<lcf_problem *> owned;             // problems created by lcf_sed_batch_create (destroyed with the batch)
    void *d_shared = nullptr, *d_shared_tiles = nullptr;   // their device arrays / tile tables: one allocation each
    size_t chain_cap = 0, lnp_cap = 0;            // grow-only chain buffers
    double *d_lpseudo = nullptr, *d_summary = nullptr;
    size_t lpseudo_cap = 0;
    int spec = 0;                                 // look-ahead rounds in k_chain (small ensembles)
    double *d_spec = nullptr;                     // [nprob][32] log-posteriors of a round's virtual walkers + a NaN counter
    ~lcf_batch() {
        for (lcf_problem *p : owned) delete p;
        cudaFree(d_shared); cudaFree(d_shared_tiles); cudaFree(d_lpseudo); cudaFree(d_summary); cudaFree(d_spec);
        cudaFree(d_probs); cudaFree(d_tiles); cudaFree(d_order); cudaFree(d_coords); cudaFree(d_logp); cudaFree(d_chain); cudaFree(d_lnp);
        cudaFree(d_acc); cudaFree(d_status);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

// -----------------------------------------------------------------------------------------
// kernel dispatch
// -----------------------------------------------------------------------------------------
typedef void (*PassKernel)(const ProblemDev, const TileDev, const MoveDev);
typedef void (*SegKernel)(const ProblemDev, const TileDev, const MoveDev, const SegDev);
typedef void (*ChainKernel)(const BatchDev);
typedef void (*RingKernel)(const ProblemDev, const TileDev, const RingDev);

// LCF_DEV_ONLY_MODEL=<id> (tools/microbench builds): instantiate the FP32 half-step kernel of one model only, so that a
// kernel experiment compiles in seconds.  Never defined for the shipped library.
#ifdef LCF_DEV_ONLY_MODEL
template <typename R> PassKernel pass_kernel_for(int model, int l = -1, bool plain = false, int ppl = 2) {
    if (model != LCF_DEV_ONLY_MODEL) return nullptr;
    if constexpr (sizeof(R) == 4 && LCF_DEV_ONLY_MODEL <= 3) if (l == 5 && plain && ppl == 4) return k_pass<LCF_DEV_ONLY_MODEL, R, 5, true, 4>;
    if (l == 5) return plain ? k_pass<LCF_DEV_ONLY_MODEL, R, 5, true> : k_pass<LCF_DEV_ONLY_MODEL, R, 5, false>;
    return k_pass<LCF_DEV_ONLY_MODEL, R, -1, false>;
}
template <typename R> SegKernel seg_kernel_for(int model) { return model == LCF_DEV_ONLY_MODEL ? k_pass_seg<LCF_DEV_ONLY_MODEL, R> : nullptr; }
template <typename R> ChainKernel chain_kernel_for(int model) { return model == LCF_DEV_ONLY_MODEL ? k_chain<LCF_DEV_ONLY_MODEL, R> : nullptr; }
template <typename R> RingKernel ring_kernel_for(int model) { return model == LCF_DEV_ONLY_MODEL ? k_ring<LCF_DEV_ONLY_MODEL, R> : nullptr; }
#else
// l = walkers-per-CTA exponent of the launch: 5 (32 walkers, every large ensemble) has its own instantiation
// and, for it, one more without the per-tile mode / use_sigma branches (plain = no intrinsic scatter, not a model evaluation)
template <typename R> PassKernel pass_kernel_for(int model, int l = -1, bool plain = false, int ppl = 2) {
#define LCF_PASS_CASE(M) case M: return l == 5 ? (plain ? k_pass<M, R, 5, true> : k_pass<M, R, 5, false>) : k_pass<M, R, -1, false>;
    if constexpr (sizeof(R) == 4) if (l == 5 && plain && ppl == 4) {   // four points per lane (points_per_lane below)
        if (model == 1) return k_pass<1, R, 5, true, 4>;
        if (model == 2) return k_pass<2, R, 5, true, 4>;
        if (model == 3) return k_pass<3, R, 5, true, 4>;
    }
    switch (model) {
        LCF_PASS_CASE(1) LCF_PASS_CASE(2) LCF_PASS_CASE(3) LCF_PASS_CASE(4)
        LCF_PASS_CASE(5) LCF_PASS_CASE(6) LCF_PASS_CASE(7) LCF_PASS_CASE(8)
    }
#undef LCF_PASS_CASE
    return nullptr;
}
template <typename R> SegKernel seg_kernel_for(int model) {           // bank streamed in segments (larger than shared memory)
    switch (model) {
        case 1: return k_pass_seg<1, R>;
        case 2: return k_pass_seg<2, R>;
        case 3: return k_pass_seg<3, R>;
        case 4: return k_pass_seg<4, R>;
        case 5: return k_pass_seg<5, R>;
        case 6: return k_pass_seg<6, R>;
        case 7: return k_pass_seg<7, R>;
        case 8: return k_pass_seg<8, R>;
    }
    return nullptr;
}
template <typename R> ChainKernel chain_kernel_for(int model) {
    switch (model) {
        case 1: return k_chain<1, R>;
        case 2: return k_chain<2, R>;
        case 3: return k_chain<3, R>;
        case 4: return k_chain<4, R>;
        case 5: return k_chain<5, R>;
        case 6: return k_chain<6, R>;
        case 7: return k_chain<7, R>;
        case 8: return k_chain<8, R>;
    }
    return nullptr;
}
template <typename R> RingKernel ring_kernel_for(int model) {
    switch (model) {
        case 1: return k_ring<1, R>;
        case 2: return k_ring<2, R>;
        case 3: return k_ring<3, R>;
        case 4: return k_ring<4, R>;
        case 5: return k_ring<5, R>;
        case 6: return k_ring<6, R>;
        case 7: return k_ring<7, R>;
        case 8: return k_ring<8, R>;
    }
    return nullptr;
}
#endif

// nsamples >= 0: bank slice of that many samples staged at a time (segmented launches) instead of the whole bank
size_t smem_bytes(const lcf_problem *p, int wpb, int nw, int ncluster = kMaxCluster, int nq = 1, int nsamples = -1) {
    const int nspl = (p->dev.model >= 5 && p->dev.model <= 7) ? p->dev.nfilters * p->dev.spl_nint : 0;
    const int ns = nsamples >= 0 ? nsamples : p->dev.nsamples;
    if (p->precision == LCF_PRECISION_FP32)
        return SmemLayout<float>(ns, p->dev.nfilters, wpb, nw, p->dev.ndim, p->dev.model == 3, nspl, ncluster, nq).total;
    return SmemLayout<double>(ns, p->dev.nfilters, wpb, nw, p->dev.ndim, p->dev.model == 3, nspl, ncluster, nq).total;
}

constexpr int kSplitUnits = 8;        // units of the structured chi-square sums (MoveDev::nq) of the shapes that use them

constexpr size_t kSmemMax = 227 * 1024;

// cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel and process-wide, and a later, smaller value LOWERS it: two live
// problems that share an instantiation but need different amounts (a large UV / JWST bank next to a small one) would break each
// other's cached launches.  Keep a high-water mark per kernel and only ever raise the limit.
int ensure_dynamic_smem(const void *kernel, size_t bytes) {
    static std::mutex mu;
    static std::map<const void *, size_t> high;
    if (bytes <= 40 * 1024) return 0;                     // below the 48 KB every kernel gets (static shared memory included)
    std::lock_guard<std::mutex> lock(mu);
    size_t &h = high[kernel];
    if (bytes > h) {
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        h = bytes;
    }
    return 0;
}

// Tile table for lanes-per-point exponent lt (a tile = 2 * (32 >> lt) points of one filter), in the order of the points (filter by
// filter); warp slot j of a launch takes the tiles j, j + P, j + 2P, ...  (Sorting the tiles by cost and dealing them in serpentine
// order to balance the slots was measured this round: no gain on cfg1 / cfg4 / cfg5 -- profiles/round2_kernel_variants.jsonl -- so
// the natural order stays.)  Problems of a flat-array batch share one allocation for all their tables (negative ntiles marks those).
// Photometry points per lane and tile: 4 in the 32-walker plain FP32 launches of the one-blackbody models (ShockCooling 1-3: the
// per-tile work is paid once per four points), else 2.  LCF_PPL=2 (experiments) keeps two everywhere.
int points_per_lane(const lcf_problem *p, int l, bool plain) {
    static const bool two = [] { const char *e = getenv("LCF_PPL"); return e && e[0] == '2'; }();
    return (!two && l == 5 && plain && p->precision == LCF_PRECISION_FP32 && p->dev.model >= 1 && p->dev.model <= 3) ? 4 : 2;
}

int get_tiles(lcf_problem *p, int lt, int period, TileDev *out, int ppl = 2) {
    (void)period;
    const int key = lt * 1024 + (ppl == 2 ? 0 : ppl);
    auto it = p->tile_tabs.find(key);
    if (it != p->tile_tabs.end()) { *out = it->second; if (out->ntiles < 0) out->ntiles = -out->ntiles; return 0; }
    const int ppt = ppl * (32 >> lt);            // `ppl` points per lane
    std::vector<int4> t;
    const int N = p->dev.npoints;
    int i = 0;
    while (i < N) {
        int f = p->h_point_filter[i], j = i;
        while (j < N && p->h_point_filter[j] == f) ++j;
        for (int s = i; s < j; s += ppt) t.push_back(make_int4(s, std::min(ppt, j - s), f, 0));
        i = j;
    }
    void *d = nullptr;
    int rc = upload(t, &d);
    if (rc) return rc;
    p->allocs.push_back(d);
    TileDev td;
    td.tiles = reinterpret_cast<const int4 *>(d);
    td.ntiles = (int)t.size();
    p->tile_tabs[key] = td;
    *out = td;
    return 0;
}

// Tile tables of MANY problems (all six walkers-per-CTA exponents) in one allocation and one copy: the problems of a batch
// created from flat arrays are tiny (an SED epoch has 3-9 points), and a cudaMalloc + cudaMemcpy pair per problem and
// shape was most of the time spent creating them.  The caller owns `*block`.
int build_tiles_shared(const std::vector<lcf_problem *> &probs, void **block) {
    std::vector<int4> all;
    std::vector<size_t> first;
    std::vector<int> count;
    for (lcf_problem *p : probs) {
        const int N = p->dev.npoints;
        for (int l = 0; l < 6; ++l) {
            const int ppt = 2 * (32 >> l);
            first.push_back(all.size());
            int i = 0;
            while (i < N) {
                int f = p->h_point_filter[i], j = i;
                while (j < N && p->h_point_filter[j] == f) ++j;
                for (int s = i; s < j; s += ppt) all.push_back(make_int4(s, std::min(ppt, j - s), f, 0));
                i = j;
            }
            count.push_back((int)(all.size() - first.back()));
        }
    }
    int rc = upload(all, block);
    if (rc) return rc;
    size_t k = 0;
    for (lcf_problem *p : probs)
        for (int l = 0; l < 6; ++l, ++k) {
            TileDev td;
            td.tiles = reinterpret_cast<const int4 *>(*block) + first[k];
            td.ntiles = -count[k];                        // negative: a natural-order table that serves every dealing period (get_tiles)
            p->tile_tabs[l * 1024] = td;
        }
    return 0;
}

// Launch shape for an active set of Ns walkers: walkers per CTA (2^l), warps per CTA and cluster size S (CTAs that
// share one walker group and split its light curve tile-wise).  A cost model in SM clocks, calibrated on B200 shape
// sweeps of cfg1 / cfg2 / cfg4 (tools/bench_configs.py --tune, gpurun_out r2_sweep*), picks it:
//   * XU-bound time of the busiest SM: CTAs are dealt round-robin, so it executes n = ceil(CTAs / SMs) of them; the XU
//     pipe only saturates with ~32 resident warps (isolated loop: 71 % at 8 warps, 87 % at 16, 97 % at 24);
//   * latency-bound time: every batch of co-resident CTAs needs the serial FP64 proposal/setup phase plus one warp's
//     chain of tiles (a tile = one dependent pass over a transmission curve);
//   * a tile is always 64 (walker, point) pairs of ONE filter, so small wpb wastes lanes when a filter has few points:
//     that is in ntiles(l);
//   * clusters and very small CTAs carry measured penalties (redundant setup per CTA, co-scheduling constraints).
int count_tiles(const lcf_problem *p, int l, int ppl = 2) {
    const int slots = ppl * (32 >> l), N = p->dev.npoints;
    int tiles = 0, i = 0;
    while (i < N) {
        int f = p->h_point_filter[i], j = i;
        while (j < N && p->h_point_filter[j] == f) ++j;
        tiles += (j - i + slots - 1) / slots;
        i = j;
    }
    return tiles;
}

struct Shape {
    int l, nw, cluster; size_t smem; double cost; int ks;
    int nq;             // units of the structured chi-square sums: kSplitUnits for unsplit shapes whose warps have >= 2 nq tile rows, else 1
    long long flat;     // > 0: launch this many CTAs, each with the same share of the (group, unit) space (needs the caller's scratch)
    int seg;            // > 0: the bank is streamed through shared memory in segments of at most this many samples (k_pass_seg)
};

bool sum_units_enabled() {                 // LCF_SUM_UNITS=1 (experiments): plain per-warp sums everywhere, no flat split
    static const bool on = [] { const char *e = getenv("LCF_SUM_UNITS"); return !(e && e[0] == '1' && e[1] == 0); }();
    return on;
}

int flat_mode() {                          // lcf_set_tuning_flat, else LCF_FLAT=0/1 from the environment, else the cost model (-1)
    if (g_tune_flat >= 0) return g_tune_flat;
    static const int env = [] { const char *e = getenv("LCF_FLAT"); return e && (e[0] == '0' || e[0] == '1') ? e[0] - '0' : -1; }();
    return env;
}

// A filter bank larger than shared memory (FP64: ~14 000 transmission samples, e.g. a light curve through many JWST / GALEX
// filters) is streamed: runs of consecutive filters whose pair records (and, ShockCooling3, per-walker weight table) fit next to
// the per-walker state of a fixed small shape (8 walkers x 16 warps; ShockCooling3: 4 walkers).  `force` (LCF_SEG_SAMPLES, tests):
// at most that many samples per segment whatever fits.  Built once per problem.
int build_segments(lcf_problem *p, int force, Shape *bs) {
    const bool f32 = p->precision == LCF_PRECISION_FP32;
    static const int env_nw = [] { const char *e = getenv("LCF_SEG_WARPS"); const int v = e ? atoi(e) : 0; return v == 4 || v == 8 || v == 16 ? v : 0; }();   // (experiments)
    const int l = p->dev.model == 3 ? 2 : 3, nw = env_nw ? env_nw : 16, wpb = 1 << l;   // (16 warps: +18 % over 8 on the 30-filter FP64 case)
    if (!p->d_segs) {
        const size_t fixed = smem_bytes(p, wpb, nw, kMaxCluster, 1, 0);
        const size_t per_sample = (f32 ? 4 : 8) * (size_t)(2 + (p->dev.model == 3 ? wpb : 0));
        if (fixed + 256 > kSmemMax) return fail(LCF_ERR_ARG, "too many filters for shared memory");
        long long cap = (long long)((kSmemMax - fixed - 256) / per_sample) & ~1LL;     // samples per segment (256 B: alignment of the carve-up)
        int widest = 0;
        for (int r : p->h_filter_records) widest = std::max(widest, 2 * r);
        if (force > 0) cap = std::min<long long>(cap, std::max(force & ~1, widest));
        if (widest > cap)
            return fail(LCF_ERR_ARG, "one filter has %d transmission samples, more than the %lld that fit in shared memory", widest, cap);
        const int F = (int)p->h_filter_records.size();
        std::vector<int4> segs(F);
        const int ns = lcf_plan_bank_segments(p->h_filter_records.data(), F, cap, reinterpret_cast<int *>(segs.data()), F);
        if (ns < 1) return fail(LCF_ERR_STATE, "bank segmentation failed");
        segs.resize(ns);
        int most = 0;
        for (const int4 &sg : segs) most = std::max(most, 2 * sg.w);
        int4 *d = nullptr;
        CUDA_TRY(cudaMalloc(&d, segs.size() * sizeof(int4)));
        p->allocs.push_back(d);
        CUDA_TRY(cudaMemcpy(d, segs.data(), segs.size() * sizeof(int4), cudaMemcpyHostToDevice));
        p->d_segs = d; p->nseg = (int)segs.size(); p->seg_samples = most; p->seg_wpb = wpb;
    }
    bs->l = l; bs->nw = nw; bs->cluster = 1; bs->ks = 0; bs->nq = 1; bs->flat = 0; bs->seg = p->seg_samples;
    bs->smem = smem_bytes(p, wpb, nw, kMaxCluster, 1, p->seg_samples);
    bs->cost = 1e12;                                       // never a candidate for the persistent / look-ahead kernels
    if (bs->smem > kSmemMax) return fail(LCF_ERR_STATE, "segmented bank does not fit (%zu bytes)", bs->smem);
    return 0;
}

int choose_shape(lcf_problem *p, long long Ns, Shape *out) {
    const int g_flat = flat_mode();
    const int tune = (((g_tune_wpb * 64 + g_tune_nw) * 16 + g_tune_cluster) * 8 + (g_tune_ks + 1)) * 4 + (g_flat + 1);
    if (p->shape_cache.Ns == Ns && p->shape_cache.tune == tune) {
        out->l = p->shape_cache.l; out->nw = p->shape_cache.nw; out->cluster = p->shape_cache.cluster; out->smem = p->shape_cache.smem;
        out->cost = p->shape_cache.cost; out->ks = p->shape_cache.ks; out->nq = p->shape_cache.nq; out->flat = p->shape_cache.flat;
        out->seg = p->shape_cache.seg;
        return 0;
    }
    int force_seg = 0;
    if (const char *e = getenv("LCF_SEG_SAMPLES")) force_seg = atoi(e);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    const bool f32 = p->precision == LCF_PRECISION_FP32;
    const double nbb = p->dev.model == 4 ? 2. : 1.;
    const double K = std::max(1., p->mean_samples) * nbb;            // Planck samples per (walker, point)
    const double pipe_tile = 64. * (f32 ? (K + 5.) / 14.5 : 0.9 * K);    // SM clocks per tile at full pipe rate
    const double lat_tile = (f32 ? 70. : 600.) * K + 500.;           // clocks one warp needs for a tile on its own
    double best = 1e300, best_rank = 1e300;
    Shape bs = {5, 16, 1, 0, 0., 0, 1, 0, 0};
    for (int l = 5; l >= 0 && force_seg <= 0; --l) {
        if (g_tune_wpb > 0 && (1 << l) != g_tune_wpb) continue;
        const long long groups = (Ns + (1 << l) - 1) >> l;
        // split-K: 2^ks lanes share a (walker, point pair) and each sweeps 1/2^ks of the filter's samples (front end repeated)
        for (int ks = 0; l + ks <= 5; ++ks) {
            if (g_tune_ks >= 0 && ks != g_tune_ks) continue;
            const int ntiles = count_tiles(p, l + ks);
            const double Kc = K / (double)(1 << ks);
            const double pipe_tile = 64. * (f32 ? (Kc + 5.) / 14.5 : 0.9 * Kc + 4.);   // SM clocks per tile at full pipe rate
            // clocks one warp needs for a tile on its own.  FP64 from the split-K sweep of the look-ahead shape (cfg1, 1 walker x 8 warps:
            // 2 / 4 / 8 sample chunks at 17.4 / 16.3 / 18.2 us per step): ~55 clocks per sample and ~4 300 per tile (its FP64 front end)
            const double lat_tile = (f32 ? 70. * Kc + 500. : 60. * Kc + 4000.) + (ks ? 60. * ks : 0.);
            const int nw_cand[5] = {16, 8, 4, 2, g_tune_nw};   // candidates; the last entry is the override
            for (int ci = (g_tune_nw > 0 ? 4 : 0); ci < (g_tune_nw > 0 ? 5 : 4); ++ci) {
                const int nw = nw_cand[ci];
                // structured sums (and with them the flat split) for unsplit shapes whose warps see at least two tile rows per unit;
                // a function of the problem and the shape only, never of Ns: chains do not depend on how an ensemble is sharded
                const int rows = (count_tiles(p, l + ks, ks == 0 ? points_per_lane(p, l, true) : 2) + nw - 1) / nw;   // (of the coarsest tile table a launch of this shape uses)
                // (FP32: measured on cfg2, the unit loop costs ShockCooling3 1.2 % and a flat split only pays at 1-2 waves, so the
                //  cost model keeps plain sums there unless lcf_set_tuning_flat / LCF_FLAT asks for the structured ones)
                int nq_s1 = (ks == 0 && rows >= 2 * kSplitUnits && sum_units_enabled() && (!f32 || g_flat >= 0)) ? kSplitUnits : 1;
                if (nq_s1 > 1 && smem_bytes(p, 1 << l, nw, kMaxCluster, nq_s1) > kSmemMax) nq_s1 = 1;
                const size_t sm_plain = smem_bytes(p, 1 << l, nw);
                if (sm_plain > kSmemMax) continue;
                for (int S = 1; S <= kMaxCluster; S <<= 1) {
                    const int nq = S == 1 ? nq_s1 : 1;
                    const size_t sm = nq > 1 ? smem_bytes(p, 1 << l, nw, kMaxCluster, nq) : sm_plain;
                    int occ = (int)std::min<size_t>(kSmemMax / sm, (size_t)(f32 ? 1024 : 512) / (nw * 32));
                    occ = std::max(1, std::min(occ, 32));
                    if (g_tune_cluster > 0 && S != g_tune_cluster) continue;
                    if (S > 1 && (long long)nw * S > 2LL * ntiles && g_tune_cluster == 0) break;    // nothing left to split
                    const double tiles_warp = (double)((ntiles + nw * S - 1) / (nw * S));
                    const double tiles_cta = std::min<double>(ntiles, tiles_warp * nw);
                    const long long ctas = groups * S;
                    const double n = (double)((ctas + sms - 1) / sms);               // CTAs on the busiest SM
                    const double resident = std::min<double>(n, occ);
                    const double w = std::min(32., resident * nw);
                    const double xu_eff = w >= 8. ? 1. - 0.29 * (32. - w) * (32. - w) / 576. : 0.71 * w / 8.;
                    const double fixed = 12000. + (S > 1 ? 3000. : 0.);            // proposal + priors + FP64 model constants (+ DSMEM reduce)
                    const double t_xu = n * tiles_cta * pipe_tile / xu_eff + 0.3 * fixed;
                    const double t_lat = std::ceil(n / occ) * (fixed + tiles_warp * lat_tile);
                    double cost = std::max(t_xu, t_lat);
                    for (int s2 = S; s2 > 1; s2 >>= 1) cost *= 1.15;
                    if (nw < 8) cost *= 1.15;
                    if (ctas < sms) cost *= 1. + 0.25 * (1. - (double)ctas / sms);  // measured (cfg1 sweep): idle SMs cost more than the chain model says
                    // measured (cfg2 at 25 000 / 50 000 walkers, profiles/round2_shape_model_vs_sweep.jsonl): with several waves of CTAs per SM a CTA of
                    // fewer walkers repeats its prologue and the per-point front end more often -- 1 / 2 / 4-8 / 16 walkers per CTA cost
                    // +20 / +7 / +4 / +2.5 % per walker against 32; without it the model took one walker per CTA for 12 500 walkers (+22 %)
                    if (f32 && Ns >= 64LL * sms) { static const double few[6] = {1.20, 1.07, 1.04, 1.04, 1.025, 1.}; cost *= few[l]; }
                    // Flat split (one CTA per co-resident slot, each with groups / slots of the work) against the plain launch of the same
                    // shape, in units of the time two co-resident CTAs need for a group each.  Measured on B200 (profiles/round2_flat_split.txt):
                    // a CTA alone on its SM runs at 0.69 of the paired rate; the warp schedulers favour the older of two co-resident CTAs,
                    // which finishes a statically split kernel at 0.64 of its duration and leaves the SM half empty (x 1.12); each CTA of a
                    // flat grid repeats one more prologue than it has groups.  So the split pays with one CTA per SM (FP64: +3.4 % on cfg2),
                    // and when the plain launch would spend most of its time in a partial wave (1.3 waves: +19 %); not at 5.3 waves (-3.4 %).
                    long long flat = 0;
                    if (nq > 1 && g_flat != 0 && ((occ <= 2 && l == 5) || g_flat == 1)) {   // (calibrated on the 32-walker shapes, one or two CTAs per SM)
                        const long long slots = (long long)occ * sms;
                        const double wv = (double)groups / (double)slots, full = std::floor(wv), frac = wv - full;
                        double plain_rel;
                        if (groups <= slots) plain_rel = occ == 1 ? 1. : (groups <= sms ? 0.72 : 0.72 + 0.28 * (double)(groups - sms) / (double)(slots - sms));
                        else plain_rel = full + (frac <= 0. ? 0. : (occ == 1 || frac > 0.5 ? 1. : 0.72));
                        const double fixed_rel = fixed / std::max(1., occ * tiles_cta * pipe_tile);
                        const double flat_rel = wv * (occ > 1 ? 1.12 : 1.) + fixed_rel;
                        if (groups * nq >= 2 * slots && groups != slots && (g_flat == 1 || flat_rel < 0.97 * plain_rel)) {
                            cost *= std::min(1., flat_rel / plain_rel);       // (a forced split never changes the ranking of the shapes)
                            flat = slots;
                        }
                    }
                    // Look-ahead rounds of the persistent kernel (try_ring): a shape whose grid of ~3 Ns virtual walkers is co-resident with at
                    // most two CTAs per SM runs a whole step in about the time of one half-step (measured on cfg1: 0.5-0.55 of two half-steps),
                    // so among the shapes of a small ensemble it is worth more than its half-step cost says -- e.g. FP64, 100 walkers: one walker
                    // x 8 warps (150 CTAs, two per SM) 6.2 M walker-steps/s with look-ahead rounds against 4.0 M for the half-step optimum
                    // (clusters of two 8-warp CTAs: 300 CTAs do not fit).  A function of (problem, Ns) only, like the rest of the model.
                    double rank = cost;
                    {
                        const long long ctas3 = ((3 * Ns + (1LL << l) - 1) >> l) * S;
                        if (cost <= 120000. && ctas3 <= (long long)occ * sms && ctas3 <= 2LL * sms && !flat) rank *= 0.55;
                    }
                    if (rank < best_rank * 0.97) { best_rank = rank; best = cost; bs.l = l; bs.nw = nw; bs.cluster = S; bs.smem = sm; bs.ks = ks; bs.nq = nq; bs.flat = flat; }
                }
            }
        }
    }
    int rc = 0;
    if (best >= 1e300) {
        if (g_tune_nw > 16 || g_tune_wpb > 32) return fail(LCF_ERR_ARG, "bad tuning override");
        if ((rc = build_segments(p, force_seg, &bs))) return rc;      // the bank does not fit in shared memory: stream it
        best = bs.cost;
        const bool f32s = p->precision == LCF_PRECISION_FP32;
        SegKernel k = f32s ? seg_kernel_for<float>(p->dev.model) : seg_kernel_for<double>(p->dev.model);
        if (!k) return fail(LCF_ERR_ARG, "unknown model id %d", p->dev.model);
        if ((rc = ensure_dynamic_smem(reinterpret_cast<const void *>(k), bs.smem))) return rc;
    }
    for (int plain = 0; plain < 2 && !bs.seg; ++plain) {
        const int ppl = points_per_lane(p, bs.l, plain != 0);
        PassKernel k = f32 ? pass_kernel_for<float>(p->dev.model, bs.l, plain, ppl) : pass_kernel_for<double>(p->dev.model, bs.l, plain, ppl);
        if (!k) return fail(LCF_ERR_ARG, "unknown model id %d", p->dev.model);
        if ((rc = ensure_dynamic_smem(reinterpret_cast<const void *>(k), bs.smem))) return rc;
    }
    p->shape_cache.Ns = Ns; p->shape_cache.l = bs.l; p->shape_cache.nw = bs.nw; p->shape_cache.cluster = bs.cluster;
    p->shape_cache.smem = bs.smem; p->shape_cache.tune = tune; p->shape_cache.cost = best; p->shape_cache.ks = bs.ks;
    p->shape_cache.nq = bs.nq; p->shape_cache.flat = bs.flat; p->shape_cache.seg = bs.seg;
    bs.cost = best;
    if (getenv("LCF_DEBUG_SHAPE"))
        fprintf(stderr, "[lcf] launch shape for %lld walkers: %d walkers/CTA, %d warps, cluster %d, split-K %d, %d sum units, flat grid %lld, %zu B smem, %d bank segment(s) (model %d, modelled %.0f clk)\n",
                Ns, 1 << bs.l, bs.nw, bs.cluster, 1 << bs.ks, bs.nq, bs.flat, bs.smem, bs.seg ? p->nseg : 1, p->dev.model, best);
    *out = bs;
    return 0;
}

int launch_pass(lcf_problem *p, const MoveDev &mv_in, cudaStream_t stream, long long *launches) {
    MoveDev mv = mv_in;
    if (mv.Ns <= 0) return 0;
    Shape sh;
    int rc = choose_shape(p, mv.Ns, &sh);
    if (rc) return rc;
    mv.wpb_log2 = sh.l;
    mv.ks = sh.ks;
    mv.nq = sh.nq;
    {
        const int nspl = (p->dev.model >= 5 && p->dev.model <= 7) ? p->dev.nfilters * p->dev.spl_nint : 0;
        const int ns = sh.seg ? sh.seg : p->dev.nsamples;
        if (p->precision == LCF_PRECISION_FP32)
            mv.lay = SmemLayout<float>(ns, p->dev.nfilters, 1 << sh.l, sh.nw, p->dev.ndim, p->dev.model == 3, nspl, kMaxCluster, sh.nq);
        else
            mv.lay = SmemLayout<double>(ns, p->dev.nfilters, 1 << sh.l, sh.nw, p->dev.ndim, p->dev.model == 3, nspl, kMaxCluster, sh.nq);
        if (mv.lay.total != sh.smem) return fail(LCF_ERR_STATE, "shared-memory layout mismatch");
    }
    const bool plain = !sh.seg && mv.mode != MODE_MODEL && !p->dev.use_sigma;
    const int ppl = sh.seg ? 2 : points_per_lane(p, sh.l, plain);
    PassKernel k = (p->precision == LCF_PRECISION_FP32) ? pass_kernel_for<float>(p->dev.model, sh.l, plain, ppl)
                                                        : pass_kernel_for<double>(p->dev.model, sh.l, plain, ppl);
    const long long ngroups = (mv.Ns + (1 << sh.l) - 1) / (1 << sh.l);
    if (ngroups * sh.nq >= (1LL << 31)) return fail(LCF_ERR_ARG, "too many walker groups for one launch");
    long long clusters = std::min<long long>(ngroups, (1LL << 30) / sh.cluster);
    if (sh.flat > 0 && sh.nq > 1 && mv.split_part && mv.split_tick && mv.mode != MODE_MODEL) clusters = sh.flat;   // equal shares of the (group, unit) space
    else if (sh.nq > 1) { mv.split_part = nullptr; mv.split_tick = nullptr; }                                   // one whole group per CTA
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(clusters * sh.cluster), 1, 1);
    cfg.blockDim = dim3((unsigned)(sh.nw * 32), 1, 1);
    cfg.dynamicSmemBytes = sh.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (!getenv("LCF_NO_PDL")) {           // overlap this launch with the tail of the previous kernel (griddepcontrol.wait in k_pass)
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (sh.cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)sh.cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    TileDev tiles;
    if ((rc = get_tiles(p, sh.l + sh.ks, sh.nw * sh.cluster, &tiles, ppl))) return rc;
    if (sh.seg) {
        SegKernel ksg = (p->precision == LCF_PRECISION_FP32) ? seg_kernel_for<float>(p->dev.model) : seg_kernel_for<double>(p->dev.model);
        SegDev sd;
        sd.segs = p->d_segs; sd.nseg = p->nseg;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, ksg, p->dev, tiles, mv, sd));
    } else
    CUDA_TRY(cudaLaunchKernelEx(&cfg, k, p->dev, tiles, mv));
    p->last_launch.wpb = 1 << sh.l; p->last_launch.nw = sh.nw; p->last_launch.cluster = sh.cluster; p->last_launch.ks = sh.ks;
    p->last_launch.grid = clusters * sh.cluster; p->last_launch.variant = sh.seg ? 5 : (sh.l == 5 ? (plain ? 2 : 1) : 0);
    p->last_launch.nq = sh.nq; p->last_launch.groups = ngroups; p->last_launch.ppl = ppl;
    if (launches) ++*launches;
    return 0;
}

// 2^(j/256) table of the FP64 inner loop, written once per device (kernels copy it into shared memory)
int ensure_e2tab() {
    static std::mutex mu;
    static std::map<int, bool> done;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return 0;
    double h[kE2TabSize];
    for (int j = 0; j < kE2TabSize; ++j) h[j] = std::exp2((double)j / (double)kE2TabSize);
    CUDA_TRY(cudaMemcpyToSymbol(g_e2tab, h, sizeof(h)));
    done[dev] = true;
    return 0;
}

int check_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(LCF_ERR_CUDA, "no CUDA device available (%s); liblcf_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return 0;
}

template <typename R>
int build_problem_arrays(const lcf_problem_desc *d, lcf_problem *p, double scale, Arena &arena) {
    const int F = d->nfilters, N = d->npoints;
    // Samples whose weight is exactly zero (transmission curves start and end with T = 0) contribute exactly nothing to the
    // trapezoid sum, so they are left out of the device bank; every filter is then padded to an even number of samples
    // (pad: a = last a, w = 0): the kernels consume sample pairs and the TMA bulk copy moves multiples of 16 bytes.
    std::vector<std::vector<int>> keep(F);
    std::vector<int> foff(F + 1, 0);
    for (int f = 0; f < F; ++f) {
        const int b0 = d->bank_offsets[f], n = d->bank_offsets[f + 1] - b0;
        for (int k = 0; k < n; ++k)
            if (d->bank_w[b0 + k] != 0.) keep[f].push_back(b0 + k);
        if (keep[f].empty()) keep[f].push_back(b0);            // an all-zero curve: one (zero-weight) sample
        foff[f + 1] = foff[f] + (((int)keep[f].size() + 1) & ~1);
    }
    const int ns_pad = foff[F];
    typedef typename Vec2<R>::type R2;
    typedef typename Vec4<R>::type R4;
    const double wfac = (d->model_id == LCF_MODEL_SHOCKCOOLING3 ? consts().c4 : 1.) / scale;
    std::vector<R4> bank(ns_pad / 2);                     // pair records (a0, a1, w0, w1)
    std::vector<int4> finfo(F);
    std::vector<R> kap(ns_pad, (R)0);
    double kept = 0.;
    for (int f = 0; f < F; ++f) {
        const int n = (int)keep[f].size();
        double mn = INFINITY, mx = 0.;
        for (int k = 0; k < foff[f + 1] - foff[f]; ++k) {
            const int src = keep[f][std::min(k, n - 1)];
            const R a = (R)(d->bank_alpha[src] * kLog2e);
            const R w = (k < n) ? (R)(d->bank_w[src] * wfac) : (R)0;
            R4 &rec = bank[(foff[f] + k) >> 1];
            if (k & 1) { rec.y = a; rec.w = w; } else { rec.x = a; rec.z = w; }
            if (d->bank_kappa) kap[foff[f] + k] = (R)(d->bank_kappa[src] * 0.4 * 3.3219280948873623479);   // 0.4*log2(10)
            mn = std::min(mn, (double)a);
            mx = std::max(mx, (double)a);
        }
        const float fmn = (float)mn, fmx = (float)mx;       // FP32 fast-path guards (unused in FP64 mode)
        int bmn, bmx;
        memcpy(&bmn, &fmn, 4);
        memcpy(&bmx, &fmx, 4);
        finfo[f] = make_int4(foff[f] >> 1, (foff[f + 1] - foff[f]) >> 1, bmn, bmx);
        p->h_filter_records.push_back((foff[f + 1] - foff[f]) >> 1);
    }
    for (int i = 0; i < N; ++i) kept += (double)(foff[d->point_filter[i] + 1] - foff[d->point_filter[i]]);
    p->mean_samples = kept / std::max(1, N);                // device samples per photometry point (launch-shape model)
    std::vector<R4> obs(N);
    for (int i = 0; i < N; ++i) {
        obs[i].x = (R)(d->y[i] / scale);
        if (d->use_sigma) {
            double dys = d->dy[i] / scale;
            double su = (d->sigma_type == 0 ? d->dy[i] : d->sigma_unit_abs) / scale;
            obs[i].y = (R)(dys * dys);
            obs[i].z = (R)(su * su);
        } else {
            obs[i].y = (R)(scale / d->dy[i]);
            obs[i].z = (R)0;
        }
        obs[i].w = (R)0;
    }
    ProblemDev &P = p->dev;
    arena.add(bank, &P.bank);
    arena.add(kap, &P.kappa);
    arena.add(finfo, reinterpret_cast<const void **>(&P.finfo));
    arena.add(obs, &P.obs);
    P.nsamples = ns_pad;
    if (d->sifto_coef && d->sifto_nknots >= 2) {
        const int nint = d->sifto_nknots - 1;
        typedef typename Vec4<R>::type R4;
        std::vector<R4> spl((size_t)F * nint);
        for (size_t i = 0; i < spl.size(); ++i) {
            spl[i].x = (R)(d->sifto_coef[4 * i + 0] / scale);
            spl[i].y = (R)(d->sifto_coef[4 * i + 1] / scale);
            spl[i].z = (R)(d->sifto_coef[4 * i + 2] / scale);
            spl[i].w = (R)(d->sifto_coef[4 * i + 3] / scale);
        }
        arena.add(spl, &P.spl);
        P.spl_nint = nint;
        P.spl_x0 = d->sifto_x0;
        P.spl_dx = d->sifto_dx;
        if (d->model_id >= 5 && d->model_id <= 7) {        // spline origin, spacing, reciprocal spacing as `real` constants
            P.dk[0] = d->sifto_x0; P.dk[1] = d->sifto_dx; P.dk[2] = 1. / d->sifto_dx;
            for (int i = 0; i < 4; ++i) P.fk[i] = (float)P.dk[i];
        }
    }
    return 0;
}

}  // namespace

// -----------------------------------------------------------------------------------------
// C ABI
// -----------------------------------------------------------------------------------------
extern "C" {

int lcf_abi_version(void) { return LCF_ABI_VERSION; }
const char *lcf_last_error(void) { return g_err.c_str(); }

int lcf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lcf_set_device(int device) {
    int rc = check_device();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    return 0;
}

int lcf_set_tuning(int walkers_per_cta, int warps_per_cta) {
    if (walkers_per_cta < 0 || walkers_per_cta > 32 || (walkers_per_cta & (walkers_per_cta - 1)))
        return fail(LCF_ERR_ARG, "walkers_per_cta must be 0 or a power of two <= 32");
    if (warps_per_cta < 0 || warps_per_cta > 16) return fail(LCF_ERR_ARG, "warps_per_cta must be in [0, 16]");
    g_tune_wpb = walkers_per_cta;
    g_tune_nw = warps_per_cta;
    g_tune_cluster = 0;
    return 0;
}

#ifdef LCF_X_TIMING
int lcf_debug_phase_clocks(unsigned long long *out) {   // experiment builds only: read and reset the per-phase clock sums
    unsigned long long z[10] = {0};
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_phase_clk, sizeof(z));
    cudaMemcpyToSymbol(g_phase_clk, z, sizeof(z));
    return 0;
}
int lcf_debug_cta_log(unsigned long long *out /* [4096][3] */) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_cta_log, sizeof(unsigned long long) * 3 * 4096);
    return 0;
}
#endif

int lcf_set_tuning_split(int sample_chunks) {
    if (sample_chunks < 0 || sample_chunks > 32 || (sample_chunks & (sample_chunks - 1))) return fail(LCF_ERR_ARG, "sample_chunks must be 0 or a power of two <= 32");
    g_tune_ks = -1;
    for (int k = 0; sample_chunks && k <= 5; ++k) if ((1 << k) == sample_chunks) g_tune_ks = k;
    return 0;
}

// Host-only planner of the segmented launches (no device needed; build_segments uses it, tests call it on the CPU).
int lcf_plan_bank_segments(const int *pair_records, int nfilters, int64_t cap_samples, int *segs_out, int max_segs) {
    if (!pair_records || !segs_out || nfilters <= 0 || max_segs <= 0) return -1;
    int n = 0, f0 = 0;
    long long pair0 = 0, pairs = 0;
    auto emit = [&](int f1) {
        if (n >= max_segs) return false;
        segs_out[4 * n] = f0; segs_out[4 * n + 1] = f1; segs_out[4 * n + 2] = (int)pair0; segs_out[4 * n + 3] = (int)pairs;
        ++n;
        return true;
    };
    for (int f = 0; f < nfilters; ++f) {
        const long long r = pair_records[f];
        if (r < 0 || 2 * r > cap_samples) return -1;                  // one filter alone exceeds the capacity
        if (f > f0 && 2 * (pairs + r) > cap_samples) {
            if (!emit(f)) return -1;
            f0 = f; pair0 += pairs; pairs = 0;
        }
        pairs += r;
    }
    if (!emit(nfilters)) return -1;
    return n;
}

int lcf_set_tuning_flat(int mode) {
    if (mode < -1 || mode > 1) return fail(LCF_ERR_ARG, "flat-split mode must be -1 (cost model), 0 (off) or 1 (on)");
    g_tune_flat = mode;
    return 0;
}

int lcf_set_tuning_ex(int walkers_per_cta, int warps_per_cta, int cluster_size) {
    if (cluster_size < 0 || cluster_size > kMaxCluster || (cluster_size & (cluster_size - 1)))
        return fail(LCF_ERR_ARG, "cluster_size must be 0 or a power of two <= %d", kMaxCluster);
    int rc = lcf_set_tuning(walkers_per_cta, warps_per_cta);
    if (rc) return rc;
    g_tune_cluster = cluster_size;
    return 0;
}

static int problem_create_impl(const lcf_problem_desc *d, lcf_problem **out, bool defer) {
    if (!d || !out) return fail(LCF_ERR_ARG, "null argument");
    *out = nullptr;
    const int nm = model_nparams(d->model_id);
    if (nm < 0) return fail(LCF_ERR_ARG, "unknown model id %d", d->model_id);
    if (d->precision != LCF_PRECISION_FP64 && d->precision != LCF_PRECISION_FP32) return fail(LCF_ERR_ARG, "bad precision");
    if (d->ndim != nm + (d->use_sigma ? 1 : 0))
        return fail(LCF_ERR_ARG, "ndim = %d but model %d takes %d parameters%s", d->ndim, d->model_id, nm,
                    d->use_sigma ? " + sigma" : "");
    if (d->ndim > LCF_MAX_NDIM) return fail(LCF_ERR_ARG, "ndim too large");
    if (d->sigma_type != 0 && d->sigma_type != 1)
        return fail(LCF_ERR_ARG, "sigma_type must either be \"relative\" or \"absolute\"");   // models.py:126
    if (d->npoints <= 0 || d->nfilters <= 0) return fail(LCF_ERR_ARG, "empty light curve or filter bank");
    if (!d->bank_offsets || !d->bank_alpha || !d->bank_w || !d->t || !d->point_filter || !d->y || !d->dy || !d->prior_kind ||
        !d->prior_min || !d->prior_max)
        return fail(LCF_ERR_ARG, "null array in problem description");
    if (d->model_id == LCF_MODEL_SHOCKCOOLING3 && !d->bank_kappa) return fail(LCF_ERR_ARG, "ShockCooling3 needs bank_kappa");
    if (d->model_id >= 5 && d->model_id <= 7 && (!d->sifto_coef || !d->filter_role || d->sifto_nknots < 2))
        return fail(LCF_ERR_ARG, "CompanionShocking models need the SiFTO spline table and filter roles");
    if (d->bank_offsets[0] != 0) return fail(LCF_ERR_ARG, "bank_offsets[0] must be 0");
    for (int f = 0; f < d->nfilters; ++f)
        if (d->bank_offsets[f + 1] < d->bank_offsets[f] + 2) return fail(LCF_ERR_ARG, "filter %d has fewer than 2 samples", f);
    for (int i = 0; i < d->npoints; ++i) {
        if (d->point_filter[i] < 0 || d->point_filter[i] >= d->nfilters) return fail(LCF_ERR_ARG, "point_filter out of range");
        if (i && d->point_filter[i] < d->point_filter[i - 1]) return fail(LCF_ERR_ARG, "points must be grouped by filter");
    }
    int rc = check_device();
    if (rc) return rc;

    if (d->precision == LCF_PRECISION_FP64 && (rc = ensure_e2tab())) return rc;
    lcf_problem *p = new lcf_problem();
    cudaGetDevice(&p->device);
    p->precision = d->precision;
    ProblemDev &P = p->dev;
    memset(&P, 0, sizeof(P));
    P.model = d->model_id;
    P.ndim = d->ndim;
    P.nmodel = nm;
    P.use_sigma = d->use_sigma ? 1 : 0;
    P.npoints = d->npoints;
    P.nfilters = d->nfilters;
    P.kB = consts().kB;
    P.c3sq = consts().c3 * consts().c3;
    for (int i = 0; i < 16; ++i) P.mc[i] = d->model_consts[i];
    for (int i = 0; i < 4; ++i) P.dk[i] = 0.;
    if (d->model_id >= 1 && d->model_id <= 3) {          // T ~ t^eps_T, L/T^4 ~ t^(eps_L - 4 eps_T), suppression exponent alpha
        P.dk[0] = 2. * P.mc[3] - 0.5;
        P.dk[1] = -2. * P.mc[4] - 4. * (2. * P.mc[3] - 0.5);
        P.dk[2] = P.mc[2];
    } else if (d->model_id == 4) {
        P.dk[0] = P.mc[0];
        P.dk[1] = P.mc[2];
    }
    for (int i = 0; i < 4; ++i) P.fk[i] = (float)P.dk[i];
    for (int i = 0; i < d->ndim; ++i) {
        P.prior.kind[i] = d->prior_kind[i];
        P.prior.pmin[i] = d->prior_min[i];
        P.prior.pmax[i] = d->prior_max[i];
        P.prior.mean[i] = d->prior_mean ? d->prior_mean[i] : 0.;
        P.prior.std[i] = d->prior_std ? d->prior_std[i] : 1.;
    }
    // FP32 unit scale: the median uncertainty, so residuals, sigmas and model values are O(1..1e3)
    double scale = 1.;
    if (d->precision == LCF_PRECISION_FP32) {
        std::vector<double> v;
        for (int i = 0; i < d->npoints; ++i)
            if (std::isfinite(d->dy[i]) && d->dy[i] > 0.) v.push_back(d->dy[i]);
        if (!v.empty()) {
            std::nth_element(v.begin(), v.begin() + v.size() / 2, v.end());
            scale = v[v.size() / 2];
        }
    }
    P.scale = scale;
    double ct = 0.;
    if (d->use_sigma) {
        ct = d->npoints * (std::log(2. * M_PI) + 2. * std::log(scale));
    } else {
        for (int i = 0; i < d->npoints; ++i) ct += std::log(2. * M_PI * (d->dy[i] * d->dy[i]));   // models.py:135
    }
    P.const_term = ct;
    p->h_point_filter.assign(d->point_filter, d->point_filter + d->npoints);

    Arena *arena_p = new Arena();
    p->pending = arena_p;
    Arena &arena = *arena_p;
    std::vector<int> role(d->nfilters, 0);
    if (d->filter_role) role.assign(d->filter_role, d->filter_role + d->nfilters);
    std::vector<double> t(d->t, d->t + d->npoints);
    arena.add(role, reinterpret_cast<const void **>(&P.frole));
    arena.add(t, reinterpret_cast<const void **>(&P.t));
    {   // FP32 epochs: t - tref as hi + lo floats (see LaneWalker<float>), tref = first finite epoch
        P.tref = 0.;
        for (int i = 0; i < d->npoints; ++i) if (std::isfinite(d->t[i])) { P.tref = d->t[i]; break; }
        std::vector<float2> t32(d->npoints);
        for (int i = 0; i < d->npoints; ++i) {
            const double r = d->t[i] - P.tref;
            t32[i].x = (float)r;
            t32[i].y = (float)(r - (double)t32[i].x);
        }
        arena.add(t32, reinterpret_cast<const void **>(&P.t32));
        arena.add(p->h_point_filter, reinterpret_cast<const void **>(&P.pfilt));
        double sk = 0.;
        for (int i = 0; i < d->npoints; ++i) sk += d->bank_offsets[d->point_filter[i] + 1] - d->bank_offsets[d->point_filter[i]];
        p->mean_samples = sk / d->npoints;
    }
    rc = (d->precision == LCF_PRECISION_FP32) ? build_problem_arrays<float>(d, p, scale, arena) : build_problem_arrays<double>(d, p, scale, arena);
    if (!rc && !defer) {
        rc = arena.commit(p->allocs);
        delete p->pending;
        p->pending = nullptr;
    }
    if (rc) { delete p; return rc; }
    *out = p;
    return 0;
}

int lcf_problem_create(const lcf_problem_desc *d, lcf_problem **out) { return problem_create_impl(d, out, false); }

void lcf_problem_destroy(lcf_problem *p) { delete p; }

int lcf_problem_last_launch(lcf_problem *p, int *walkers_per_cta, int *warps_per_cta, int *cluster_size, int64_t *grid, int *kernel_variant) {
    if (!p) return fail(LCF_ERR_ARG, "null problem");
    if (p->last_launch.variant < 0) return fail(LCF_ERR_STATE, "no kernel has been launched for this problem yet");
    if (walkers_per_cta) *walkers_per_cta = p->last_launch.wpb;
    if (warps_per_cta) *warps_per_cta = p->last_launch.nw;
    if (cluster_size) *cluster_size = p->last_launch.cluster;
    if (grid) *grid = p->last_launch.grid;
    if (kernel_variant) *kernel_variant = p->last_launch.variant;
    return 0;
}

int lcf_problem_last_launch_ex(lcf_problem *p, int64_t *groups, int *sum_units, int *points_per_lane) {
    if (!p) return fail(LCF_ERR_ARG, "null problem");
    if (p->last_launch.variant < 0) return fail(LCF_ERR_STATE, "no kernel has been launched for this problem yet");
    if (groups) *groups = p->last_launch.groups;
    if (sum_units) *sum_units = p->last_launch.nq;
    if (points_per_lane) *points_per_lane = p->last_launch.ppl;
    return 0;
}

static int eval_common(lcf_problem *p, int mode, long long nsets, int ncols, const double *params, double *out, size_t out_per_set,
                       long long *nan_count) {
    NvtxRange r("lcf_eval");
    if (!p || !params || !out) return fail(LCF_ERR_ARG, "null argument");
    if (nsets <= 0) return 0;
    int rc = check_device();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(p->device));
    const size_t need_q = sizeof(double) * nsets * ncols, need_out = sizeof(double) * nsets * out_per_set;
    if (need_q > p->eval_q_cap) {
        cudaFree(p->d_eval_q); p->d_eval_q = nullptr; p->eval_q_cap = 0;
        CUDA_TRY(cudaMalloc(&p->d_eval_q, need_q));
        p->eval_q_cap = need_q;
    }
    if (need_out > p->eval_out_cap) {
        cudaFree(p->d_eval_out); p->d_eval_out = nullptr; p->eval_out_cap = 0;
        CUDA_TRY(cudaMalloc(&p->d_eval_out, need_out));
        p->eval_out_cap = need_out;
    }
    if (!p->d_eval_nan) CUDA_TRY(cudaMalloc(&p->d_eval_nan, sizeof(int)));
    double *d_q = p->d_eval_q, *d_out = p->d_eval_out;
    int *d_nan = p->d_eval_nan;
    CUDA_TRY(cudaMemcpy(d_q, params, need_q, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(d_nan, 0, sizeof(int)));
    if (mode == MODE_MODEL) CUDA_TRY(cudaMemset(d_out, 0, need_out));
    MoveDev mv;
    memset(&mv, 0, sizeof(mv));
    mv.mode = mode;
    mv.Ns = nsets;
    mv.qin = d_q;
    mv.out = d_out;
    mv.nanflag = d_nan;
    if (mode != MODE_MODEL && nsets >= 4096) {           // large batches may be launched as a flat split (scratch: grow-only, zeroed counters)
        if (nsets > p->split_cap) {
            cudaFree(p->d_split_part); cudaFree(p->d_split_tick); p->d_split_part = nullptr; p->d_split_tick = nullptr; p->split_cap = 0;
            CUDA_TRY(cudaMalloc(&p->d_split_part, sizeof(double) * kSplitUnits * (nsets + 32)));
            CUDA_TRY(cudaMalloc(&p->d_split_tick, sizeof(unsigned int) * (nsets + 1)));
            CUDA_TRY(cudaMemset(p->d_split_tick, 0, sizeof(unsigned int) * (nsets + 1)));
            p->split_cap = nsets;
        }
        mv.split_part = p->d_split_part;
        mv.split_tick = p->d_split_tick;
    }
    rc = launch_pass(p, mv, 0, nullptr);
    if (!rc) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = fail(LCF_ERR_CUDA, "kernel failed: %s", cudaGetErrorString(e));
    }
    if (!rc) {
        cudaMemcpy(out, d_out, sizeof(double) * nsets * out_per_set, cudaMemcpyDeviceToHost);
        int h_nan = 0;
        cudaMemcpy(&h_nan, d_nan, sizeof(int), cudaMemcpyDeviceToHost);
        if (nan_count) *nan_count = h_nan;
    }
    return rc;
}

int lcf_model_eval(lcf_problem *p, int64_t nsets, const double *params, double *out) {
    if (!p) return fail(LCF_ERR_ARG, "null problem");
    return eval_common(p, MODE_MODEL, nsets, p->dev.nmodel, params, out, (size_t)p->dev.npoints, nullptr);
}
int lcf_log_likelihood(lcf_problem *p, int64_t nsets, const double *params, double *out) {
    if (!p) return fail(LCF_ERR_ARG, "null problem");
    return eval_common(p, MODE_LOGLIKE, nsets, p->dev.ndim, params, out, 1, nullptr);
}
int lcf_log_posterior(lcf_problem *p, int64_t nsets, const double *params, double *out, int64_t *nan_count) {
    if (!p) return fail(LCF_ERR_ARG, "null problem");
    long long nc = 0;
    int rc = eval_common(p, MODE_LOGPOST, nsets, p->dev.ndim, params, out, 1, &nc);
    if (nan_count) *nan_count = nc;
    return rc;
}

// ---------------------------------------------------------------------------------------
// ensemble
// ---------------------------------------------------------------------------------------
int lcf_ensemble_create(lcf_problem *p, int64_t nwalkers, uint64_t seed, int rank, int world, lcf_ensemble **out) {
    if (!p || !out) return fail(LCF_ERR_ARG, "null argument");
    *out = nullptr;
    if (nwalkers < 2) return fail(LCF_ERR_ARG, "need at least 2 walkers");
    if (nwalkers < 2LL * p->dev.ndim)
        return fail(LCF_ERR_NWALKERS, "It is unadvisable to use a red-blue move with fewer walkers than twice the number of dimensions.");
    if (world < 1 || rank < 0 || rank >= world) return fail(LCF_ERR_ARG, "bad rank/world");
    if (nwalkers >= (1LL << 31)) return fail(LCF_ERR_ARG, "too many walkers");
    int rc = check_device();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(p->device));
    lcf_ensemble *e = new lcf_ensemble();
    e->p = p;
    e->W = nwalkers;
    e->D = p->dev.ndim;
    e->n0 = (nwalkers + 1) / 2;
    e->n1 = nwalkers - e->n0;
    e->rank = rank;
    e->world = world;
    e->seed = seed;
    for (int h = 0; h < 2; ++h) {           // contiguous slice of each colour block
        long long n = h ? e->n1 : e->n0;
        long long per = (n + world - 1) / world;
        long long b = std::min(n, per * rank), en = std::min(n, per * (rank + 1));
        e->own_begin[h] = b;
        e->own_count[h] = en - b;
    }
    // the stored chain holds the walkers this rank updates: all of them on one GPU; with equal colour slices the logical
    // walkers 2b .. 2(b + c) - 1 of a shared ensemble (1/world of the chain memory and of its allocation time)
    e->cfirst = 0; e->cw = e->W;
    if (world > 1 && e->own_begin[0] == e->own_begin[1] && e->own_count[0] == e->own_count[1]) {
        e->cfirst = 2 * e->own_begin[0];
        e->cw = 2 * e->own_count[0];
    }
    cudaError_t ce;
    if ((ce = cudaMalloc(&e->d_coords, sizeof(double) * e->W * e->D)) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_logp, sizeof(double) * e->W)) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_acc, sizeof(unsigned long long) * e->W)) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_nan, 2 * sizeof(int))) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_flags, 2 * (kMaxPeers + 1) * sizeof(unsigned int))) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_split_part, sizeof(double) * kSplitUnits * (e->W + 32))) != cudaSuccess ||
        (ce = cudaMalloc(&e->d_split_tick, sizeof(unsigned int) * (e->W + 1))) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaEventCreate(&e->ev0)) != cudaSuccess || (ce = cudaEventCreate(&e->ev1)) != cudaSuccess) {
        delete e;
        return fail(LCF_ERR_CUDA, "ensemble allocation failed: %s", cudaGetErrorString(ce));
    }
    cudaMemset(e->d_acc, 0, sizeof(unsigned long long) * e->W);
    cudaMemset(e->d_nan, 0, 2 * sizeof(int));
    cudaMemset(e->d_flags, 0, 2 * (kMaxPeers + 1) * sizeof(unsigned int));
    cudaMemset(e->d_split_tick, 0, sizeof(unsigned int) * (e->W + 1));
    *out = e;
    return 0;
}

void lcf_ensemble_destroy(lcf_ensemble *e) { delete e; }

static inline long long phys_row(const lcf_ensemble *e, long long j) { return (j & 1) ? e->n0 + (j >> 1) : (j >> 1); }

static int check_nan(lcf_ensemble *e) {
    int h[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(h, e->d_nan, 2 * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (h[0] || h[1]) cudaMemsetAsync(e->d_nan, 0, 2 * sizeof(int), e->stream);
    if (h[1]) return fail(LCF_ERR_CUDA, "multi-GPU exchange timed out waiting for a peer's half-step");
    if (h[0]) return fail(LCF_ERR_NAN, "Probability function returned NaN");
    return 0;
}

int lcf_ensemble_set_state(lcf_ensemble *e, const double *coords, const double *log_prob) {
    if (!e || !coords) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    const int D = e->D;
    const size_t nc = (size_t)e->W * D;
    // logical order in, colour-major on the device: the permutation and emcee's NaN / infinity checks run in a kernel
    if (!e->d_stage) {
        CUDA_TRY(cudaMalloc(&e->d_stage, sizeof(double) * (nc + (size_t)e->W)));
        CUDA_TRY(cudaMalloc(&e->d_stage_flag, sizeof(int)));
    }
    double *d_in = e->d_stage, *d_lp = log_prob ? e->d_stage + nc : nullptr;
    int *d_fl = e->d_stage_flag;
    cudaError_t ce = cudaMemsetAsync(d_fl, 0, sizeof(int), e->stream);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(d_in, coords, sizeof(double) * nc, cudaMemcpyHostToDevice, e->stream);
    if (ce == cudaSuccess && log_prob) ce = cudaMemcpyAsync(d_lp, log_prob, sizeof(double) * e->W, cudaMemcpyHostToDevice, e->stream);
    int h_fl = 0;
    if (ce == cudaSuccess) {
        k_set_state<<<(unsigned)((nc + 255) / 256), 256, 0, e->stream>>>(d_in, d_lp, e->W, D, e->n0, e->d_coords, e->d_logp, d_fl);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(&h_fl, d_fl, sizeof(int), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    if (ce != cudaSuccess) return fail(LCF_ERR_CUDA, "set_state failed: %s", cudaGetErrorString(ce));
    if (h_fl) e->has_state = false;                     // the previous state has been overwritten
    if (h_fl & 1) return fail(LCF_ERR_ARG, "At least one parameter value was infinite");   // emcee's initial-state check
    if (h_fl & 2) return fail(LCF_ERR_ARG, "At least one parameter value was NaN");
    if (!log_prob) {
        MoveDev mv;
        memset(&mv, 0, sizeof(mv));
        mv.mode = MODE_LOGPOST;
        mv.Ns = e->W;
        mv.qin = e->d_coords;
        mv.out = e->d_logp;
        mv.nanflag = e->d_nan;
        mv.split_part = e->d_split_part;
        mv.split_tick = e->d_split_tick;
        int rc = launch_pass(e->p, mv, e->stream, nullptr);
        if (rc) return rc;
        rc = check_nan(e);
        if (rc) return rc;
    }
    e->has_state = true;
    return 0;
}

// Start state of a SHARED ensemble: every rank passes only the walkers it owns (logical [first, first + count), both colours
// of its slice).  They are uploaded, permuted into this rank's replica, their log-posteriors evaluated here, and -- with the fused
// exchange attached -- rows and log-probabilities are stored into every peer's replica over NVLink.  The caller must make sure
// no peer is still sampling (its kernels read the replicas), and put ONE barrier between this call on all ranks and the first
// step.  Replaces "every rank uploads the full [W][D] array, evaluates its slice, all-gathers the log-probabilities".
int lcf_ensemble_set_state_slice(lcf_ensemble *e, int64_t first, int64_t count, const double *coords) {
    if (!e || !coords) return fail(LCF_ERR_ARG, "null argument");
    if (first != e->cfirst || count != e->cw) return fail(LCF_ERR_ARG, "the slice must be exactly the walkers this rank owns");
    CUDA_TRY(cudaSetDevice(e->p->device));
    const int D = e->D;
    const size_t nc = (size_t)count * D;
    if (!e->d_stage) {
        CUDA_TRY(cudaMalloc(&e->d_stage, sizeof(double) * ((size_t)e->W * D + (size_t)e->W)));
        CUDA_TRY(cudaMalloc(&e->d_stage_flag, sizeof(int)));
    }
    int h_fl = 0;
    CUDA_TRY(cudaMemsetAsync(e->d_stage_flag, 0, sizeof(int), e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->d_stage, coords, sizeof(double) * nc, cudaMemcpyHostToDevice, e->stream));
    k_set_state_slice<<<(unsigned)((nc + 255) / 256), 256, 0, e->stream>>>(e->d_stage, first, count, D, e->n0, e->d_coords, e->d_stage_flag);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(&h_fl, e->d_stage_flag, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    PublishDev pd;
    memset(&pd, 0, sizeof(pd));
    pd.npeers = e->npeers;
    for (int p = 0; p < e->npeers; ++p) { pd.peer_coords[p] = e->peer_coords[p]; pd.peer_logp[p] = e->peer_logp[p]; }
    for (int h = 0; h < 2; ++h) {                          // own rows of each colour block: evaluate, then publish
        const long long row0 = (h ? e->n0 : 0) + e->own_begin[h], nrows = e->own_count[h];
        if (nrows <= 0) continue;
        MoveDev mv;
        memset(&mv, 0, sizeof(mv));
        mv.mode = MODE_LOGPOST;
        mv.Ns = nrows;
        mv.qin = e->d_coords + row0 * D;
        mv.out = e->d_logp + row0;
        mv.nanflag = e->d_nan;
        mv.split_part = e->d_split_part;
        mv.split_tick = e->d_split_tick;
        int rc = launch_pass(e->p, mv, e->stream, nullptr);
        if (rc) return rc;
        if (e->npeers) {
            const long long n = nrows * (D + 1);
            k_publish_rows<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(e->d_coords, e->d_logp, row0, nrows, D, pd);
            CUDA_TRY(cudaGetLastError());
        }
    }
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (h_fl) e->has_state = false;
    if (h_fl & 1) return fail(LCF_ERR_ARG, "At least one parameter value was infinite");
    if (h_fl & 2) return fail(LCF_ERR_ARG, "At least one parameter value was NaN");
    int rc = check_nan(e);
    if (rc) return rc;
    e->has_state = true;
    return 0;
}

int lcf_ensemble_get_state(lcf_ensemble *e, double *coords, double *log_prob) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (!e->has_state) return fail(LCF_ERR_STATE, "ensemble has no state");
    CUDA_TRY(cudaSetDevice(e->p->device));
    const int D = e->D;
    std::vector<double> hc((size_t)e->W * D), hl(e->W);
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaMemcpy(hc.data(), e->d_coords, sizeof(double) * hc.size(), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(hl.data(), e->d_logp, sizeof(double) * hl.size(), cudaMemcpyDeviceToHost));
    for (long long j = 0; j < e->W; ++j) {
        long long r = phys_row(e, j);
        if (coords) for (int d = 0; d < D; ++d) coords[j * D + d] = hc[r * D + d];
        if (log_prob) log_prob[j] = hl[r];
    }
    return 0;
}

int lcf_ensemble_reset(lcf_ensemble *e) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    e->nstored = 0;
    CUDA_TRY(cudaMemsetAsync(e->d_acc, 0, sizeof(unsigned long long) * e->W, e->stream));
    return 0;
}

// The chain lives in stream-ordered pool memory (cudaMallocAsync): nobody but this rank touches it, and ordinary
// cudaMalloc / cudaFree of a few hundred MB become slow and erratic (tens of ms) once peer access is on, because every
// allocation is then mapped into the peers as well.
static int ensure_capacity(lcf_ensemble *e, long long need) {
    if (need <= e->cap) return 0;
    static bool pool_ready = false;
    if (!pool_ready) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, e->p->device) == cudaSuccess) {
            unsigned long long keep = ~0ull;                    // keep freed blocks in the pool for the next run
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_ready = true;
    }
    long long ncap = std::max(need, e->cap * 2);
    double *nc = nullptr, *nl = nullptr;
    CUDA_TRY(cudaMallocAsync(&nc, std::max<size_t>(16, sizeof(double) * ncap * e->cw * e->D), e->stream));
    CUDA_TRY(cudaMallocAsync(&nl, std::max<size_t>(16, sizeof(double) * ncap * e->cw), e->stream));
    if (e->nstored) {
        CUDA_TRY(cudaMemcpyAsync(nc, e->d_chain, sizeof(double) * e->nstored * e->cw * e->D, cudaMemcpyDeviceToDevice, e->stream));
        CUDA_TRY(cudaMemcpyAsync(nl, e->d_lnp, sizeof(double) * e->nstored * e->cw, cudaMemcpyDeviceToDevice, e->stream));
    }
    if (e->d_chain) CUDA_TRY(cudaFreeAsync(e->d_chain, e->stream));
    if (e->d_lnp) CUDA_TRY(cudaFreeAsync(e->d_lnp, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));                  // the new blocks are valid on every stream from here on
    e->d_chain = nc;
    e->d_lnp = nl;
    e->cap = ncap;
    return 0;
}

static void fill_move(lcf_ensemble *e, int half, int store, MoveDev &mv) {
    memset(&mv, 0, sizeof(mv));
    mv.coords = e->d_coords;
    mv.logp = e->d_logp;
    mv.accepted = store ? e->d_acc : nullptr;          // emcee counts accepted moves of stored steps only (Backend.save_step)
    mv.nanflag = e->d_nan;
    mv.W = e->W;
    mv.n0 = e->n0;
    mv.mode = MODE_MOVE;
    mv.Ns = e->own_count[half];
    mv.act_base = (half ? e->n0 : 0) + e->own_begin[half];
    mv.Nc = half ? e->n0 : e->n1;
    mv.comp_base = half ? 0 : e->n0;
    mv.seed = e->seed;
    mv.ctr = (unsigned int)(2 * e->iteration + half);
    mv.split_part = e->d_split_part;
    mv.split_tick = e->d_split_tick;
    if (store) {
        // the kernel indexes by logical walker j: bias the step pointers by the first stored walker
        mv.chain_step = e->d_chain + (e->nstored * e->cw - e->cfirst) * e->D;
        mv.lnp_step = e->d_lnp + (e->nstored * e->cw - e->cfirst);
    }
    if (e->npeers) {                                   // fused exchange: one epoch per half-step launch
        mv.npeers = e->npeers;
        mv.myrank = e->rank;
        for (int p = 0; p < e->npeers; ++p) {
            mv.peer_coords[p] = e->peer_coords[p];
            mv.peer_logp[p] = e->peer_logp[p];
            mv.peer_flags[p] = e->peer_flags[p];
            mv.peer_rank[p] = e->peer_rank[p];
        }
        mv.flags = e->d_flags;
        mv.done_count = e->d_flags + (kMaxPeers + 1);
        mv.epoch = e->epoch++;
        mv.xstatus = e->d_nan + 1;
    }
}

// ---- fused multi-GPU exchange: peer replicas mapped into this process (cudaIpc) or passed as raw pointers ----------
int lcf_ensemble_ipc_export(lcf_ensemble *e, unsigned char *out /* [3][64] */) {
    if (!e || !out) return fail(LCF_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CUDA_TRY(cudaSetDevice(e->p->device));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, e->d_coords)); memcpy(out, &h, 64);
    CUDA_TRY(cudaIpcGetMemHandle(&h, e->d_logp));   memcpy(out + 64, &h, 64);
    CUDA_TRY(cudaIpcGetMemHandle(&h, e->d_flags));  memcpy(out + 128, &h, 64);
    return 0;
}

int lcf_ensemble_peers_attach_ptrs(lcf_ensemble *e, void *const *coords, void *const *log_prob, void *const *flags) {
    if (!e || !coords || !log_prob || !flags) return fail(LCF_ERR_ARG, "null argument");
    if (e->world < 2) return fail(LCF_ERR_ARG, "ensemble was created with world = 1");
    if (e->world - 1 > kMaxPeers) return fail(LCF_ERR_ARG, "at most %d GPUs share one ensemble", kMaxPeers + 1);
    int n = 0;
    for (int r = 0; r < e->world; ++r) {
        if (r == e->rank) continue;
        if (!coords[r] || !log_prob[r] || !flags[r]) return fail(LCF_ERR_ARG, "null peer pointer for rank %d", r);
        e->peer_coords[n] = reinterpret_cast<double *>(coords[r]);
        e->peer_logp[n] = reinterpret_cast<double *>(log_prob[r]);
        e->peer_flags[n] = reinterpret_cast<unsigned int *>(flags[r]);
        e->peer_rank[n] = r;
        ++n;
    }
    e->npeers = n;
    return 0;
}

int lcf_ensemble_peers_attach_ipc(lcf_ensemble *e, const unsigned char *handles /* [world][3][64] */) {
    if (!e || !handles) return fail(LCF_ERR_ARG, "null argument");
    if (e->world < 2 || e->world - 1 > kMaxPeers) return fail(LCF_ERR_ARG, "bad world size for a shared ensemble");
    CUDA_TRY(cudaSetDevice(e->p->device));
    void *c[kMaxPeers + 1] = {nullptr}, *l[kMaxPeers + 1] = {nullptr}, *f[kMaxPeers + 1] = {nullptr};
    for (int r = 0; r < e->world; ++r) {
        if (r == e->rank) continue;
        void **dst[3] = {&c[r], &l[r], &f[r]};
        for (int k = 0; k < 3; ++k) {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + ((size_t)r * 3 + k) * 64, 64);
            CUDA_TRY(cudaIpcOpenMemHandle(dst[k], h, cudaIpcMemLazyEnablePeerAccess));
            e->ipc_opened.push_back(*dst[k]);
        }
    }
    return lcf_ensemble_peers_attach_ptrs(e, c, l, f);
}

int lcf_ensemble_peers_detach(lcf_ensemble *e) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    e->npeers = 0;
    for (void *m : e->ipc_opened) cudaIpcCloseMemHandle(m);
    e->ipc_opened.clear();
    return 0;
}

int lcf_ensemble_exchange_view(lcf_ensemble *e, void **d_flags, int *npeers) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (d_flags) *d_flags = e->d_flags;
    if (npeers) *npeers = e->npeers;
    return 0;
}

int lcf_ensemble_reserve(lcf_ensemble *e, int64_t nsteps) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    return ensure_capacity(e, e->nstored + nsteps);
}

int lcf_ensemble_half_step(lcf_ensemble *e, int half, int store) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (!e->has_state) return fail(LCF_ERR_STATE, "run_mcmc before an initial state was set");
    if (half != 0 && half != 1) return fail(LCF_ERR_ARG, "half must be 0 or 1");
    CUDA_TRY(cudaSetDevice(e->p->device));
    if (store) {
        int rc = ensure_capacity(e, e->nstored + 1);
        if (rc) return rc;
    }
    MoveDev mv;
    fill_move(e, half, store, mv);
    return launch_pass(e->p, mv, e->stream, &e->last_launches);
}

int lcf_ensemble_end_step(lcf_ensemble *e, int store) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    e->iteration += 1;
    if (store) e->nstored += 1;
    return 0;
}

// Small ensembles: the whole run as ONE cooperative launch of the persistent kernel k_ring (device-side barrier between half-steps)
// instead of two k_pass launches per step.  Used when the grid of the chosen launch shape is co-resident and a half-step is short
// enough to be launch-bound (cost model of choose_shape); LCF_RING=0 forces it off, 1 / 2 on (when it fits): 1 with one half-step per
// grid barrier, 2 with look-ahead rounds (the default whenever their larger grid is co-resident and leaves every SM at most two CTAs).
static int try_ring(lcf_ensemble *e, long long nsteps, int store, bool *used) {
    *used = false;
    lcf_problem *p = e->p;
    if (e->world != 1 || e->npeers || nsteps <= 0) return 0;
    const char *env = getenv("LCF_RING");
    if (env && env[0] == '0') return 0;
    Shape sh;
    int rc = choose_shape(p, std::max(e->n0, e->n1), &sh);
    if (rc) return rc;
    if (sh.seg) return 0;                                                    // segmented bank: half-step launches only
    if (sh.flat > 0) return 0;                                               // a flat split is a large ensemble by construction
    if (points_per_lane(p, sh.l, !p->dev.use_sigma) != 2) return 0;          // k_ring runs the two-points-per-lane code: same arithmetic as the launches it replaces
    if (!(env && (env[0] == '1' || env[0] == '2')) && sh.cost > 120000.) return 0;   // > ~60 us per half-step: launch latency is already hidden
    const bool f32 = p->precision == LCF_PRECISION_FP32;
    RingKernel k = f32 ? ring_kernel_for<float>(p->dev.model) : ring_kernel_for<double>(p->dev.model);
    if (!k) return 0;
    if ((rc = ensure_dynamic_smem(reinterpret_cast<const void *>(k), sh.smem))) return rc;
    // Look-ahead rounds (k_ring, spec_prephase): a whole step per grid barrier, for n0 + 2 n1 virtual walkers.  Worth 1.5x the
    // arithmetic only while the device is latency-bound, i.e. while the larger grid still leaves every SM at most two CTAs.
    // LCF_RING=1 keeps one half-step per barrier, LCF_RING=2 asks for the look-ahead whenever its grid is co-resident.
    const long long NV = e->n0 + 2 * e->n1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    const long long ngroups_spec = (NV + (1 << sh.l) - 1) >> sh.l;
    bool spec = e->n1 > 0 && !(env && env[0] == '1') && ((env && env[0] == '2') || ngroups_spec * sh.cluster <= 2LL * sms);
    long long ngroups = spec ? ngroups_spec : (std::max(e->n0, e->n1) + (1 << sh.l) - 1) >> sh.l;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(ngroups * sh.cluster), 1, 1);
    cfg.blockDim = dim3((unsigned)(sh.nw * 32), 1, 1);
    cfg.dynamicSmemBytes = sh.smem;
    cfg.stream = e->stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
    if (sh.cluster > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = (unsigned)sh.cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const int ring_key = (((sh.l * 64 + sh.nw) * 16 + sh.cluster) * 8 + sh.ks) * 4 + (env ? (env[0] & 3) : 0);
    if (e->ring_ok < 0 || e->ring_key != ring_key) {            // does the whole grid fit on the device at once?
        e->ring_key = ring_key;
        long long capacity = 0;
        if (sh.cluster > 1) {
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, k, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
            capacity = (long long)ncl * sh.cluster;
        } else {
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, sh.nw * 32, sh.smem) != cudaSuccess) { cudaGetLastError(); nb = 0; }
            capacity = (long long)nb * sms;
        }
        if (spec && ngroups * sh.cluster > capacity) {          // the look-ahead grid does not fit: one half-step per barrier
            spec = false;
            ngroups = (std::max(e->n0, e->n1) + (1 << sh.l) - 1) >> sh.l;
        }
        e->ring_spec = spec ? 1 : 0;
        e->ring_ok = (ngroups * sh.cluster <= capacity) ? 1 : 0;
        if (getenv("LCF_DEBUG_SHAPE"))
            fprintf(stderr, "[lcf] persistent chain kernel%s: grid %lld CTAs, co-resident capacity %lld -> %s\n", spec ? " (look-ahead rounds)" : "",
                    ngroups * sh.cluster, capacity, e->ring_ok ? "used" : "not used");
    } else if (spec != (e->ring_spec != 0)) {
        spec = e->ring_spec != 0;
        ngroups = spec ? ngroups_spec : (std::max(e->n0, e->n1) + (1 << sh.l) - 1) >> sh.l;
    }
    if (!e->ring_ok) return 0;
    cfg.gridDim = dim3((unsigned)(ngroups * sh.cluster), 1, 1);
    const size_t nx = 2 * (size_t)e->W * (e->D + 1), nq = 2 * (size_t)NV * e->D, nl = 2 * (size_t)NV, nm = 4 * (size_t)e->W;
    if (spec && !e->d_spec) CUDA_TRY(cudaMalloc(&e->d_spec, (nx + nq + nl + nm + 2) * sizeof(double)));
    if (!e->d_ring_bar) CUDA_TRY(cudaMalloc(&e->d_ring_bar, 2 * sizeof(unsigned int)));
    CUDA_TRY(cudaMemsetAsync(e->d_ring_bar, 0, 2 * sizeof(unsigned int), e->stream));
    TileDev tiles;
    if ((rc = get_tiles(p, sh.l + sh.ks, sh.nw * sh.cluster, &tiles))) return rc;
    RingDev G;
    memset(&G, 0, sizeof(G));
    G.coords = e->d_coords; G.logp = e->d_logp; G.accepted = e->d_acc; G.nanflag = e->d_nan;
    G.chain = store ? e->d_chain + e->nstored * e->W * e->D : nullptr;
    G.lnp = store ? e->d_lnp + e->nstored * e->W : nullptr;
    G.W = e->W; G.n0 = e->n0;
    G.nsteps = nsteps; G.iter0 = e->iteration;
    G.seed = e->seed;
    G.bar = e->d_ring_bar;
    G.wpb_log2 = sh.l;
    G.ks = sh.ks;
    G.nq = sh.nq;
    G.spec = spec ? 1 : 0;
    if (spec) {
        G.xbuf = e->d_spec; G.sq = G.xbuf + nx; G.snlp = G.sq + nq; G.smeta = G.snlp + nl;
        G.nan_scratch = reinterpret_cast<int *>(G.smeta + nm);
    }
    CUDA_TRY(cudaLaunchKernelEx(&cfg, k, p->dev, tiles, G));
    p->last_launch.wpb = 1 << sh.l; p->last_launch.nw = sh.nw; p->last_launch.cluster = sh.cluster; p->last_launch.ks = sh.ks;
    p->last_launch.grid = ngroups * sh.cluster; p->last_launch.variant = spec ? 4 : 3; p->last_launch.nq = sh.nq; p->last_launch.groups = ngroups; p->last_launch.ppl = 2;
    e->last_launches += 1;
    e->iteration += nsteps;
    if (store) e->nstored += nsteps;
    *used = true;
    return 0;
}

int lcf_ensemble_run(lcf_ensemble *e, int64_t nsteps, int store) {
    NvtxRange r("lcf_ensemble_run");
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (!e->has_state) return fail(LCF_ERR_STATE, "run_mcmc before an initial state was set");
    CUDA_TRY(cudaSetDevice(e->p->device));
    if (store) {
        int rc = ensure_capacity(e, e->nstored + nsteps);
        if (rc) return rc;
    }
    e->last_launches = 0;
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    bool ring = false;
    int rrc = try_ring(e, nsteps, store, &ring);
    if (rrc) return rrc;
    for (long long s = 0; s < nsteps && !ring; ++s) {
        for (int half = 0; half < 2; ++half) {
            MoveDev mv;
            fill_move(e, half, store, mv);
            int rc = launch_pass(e->p, mv, e->stream, &e->last_launches);
            if (rc) return rc;
        }
        e->iteration += 1;
        if (store) e->nstored += 1;
    }
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->last_ms = ms;
    return check_nan(e);
}

// Like lcf_ensemble_run(store = 1), and every finished step is copied to the caller's host buffers on a second stream
// while the next steps compute (event-ordered; page-locked host memory makes the copies truly asynchronous).  The chain
// also stays in HBM (lcf_ensemble_get_chain, diagnostics).  chain_host [nsteps][nwalkers][ndim], log_prob_host [nsteps][nwalkers].
int lcf_ensemble_run_to_host_slice(lcf_ensemble *e, int64_t nsteps, int64_t first, int64_t count, double *chain_host,
                                   double *log_prob_host) {
    NvtxRange r("lcf_ensemble_run_to_host");
    if (!e || !chain_host || !log_prob_host) return fail(LCF_ERR_ARG, "null argument");
    if (!e->has_state) return fail(LCF_ERR_STATE, "run_mcmc before an initial state was set");
    if (first < e->cfirst || count < 0 || first + count > e->cfirst + e->cw)
        return fail(LCF_ERR_ARG, "walker range outside the walkers this rank stores");
    CUDA_TRY(cudaSetDevice(e->p->device));
    int rc = ensure_capacity(e, e->nstored + nsteps);
    if (rc) return rc;
    if (!e->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&e->ev_step, cudaEventDisableTiming));
    }
    e->last_launches = 0;
    const size_t D = e->D, cw = (size_t)e->cw * D, lw = (size_t)e->cw;
    const size_t rel = (size_t)(first - e->cfirst);
    CUDA_TRY(cudaEventRecord(e->ev0, e->stream));
    bool ring = false;
    const long long stored0 = e->nstored;
    rc = try_ring(e, nsteps, 1, &ring);
    if (rc) return rc;
    if (ring) {            // a small ensemble: the whole chain in one launch, then one copy of the (small) chain
        CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
        CUDA_TRY(cudaMemcpy2DAsync(chain_host, count * D * sizeof(double), e->d_chain + stored0 * cw + rel * D, cw * sizeof(double),
                                   count * D * sizeof(double), nsteps, cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaMemcpy2DAsync(log_prob_host, count * sizeof(double), e->d_lnp + stored0 * lw + rel, lw * sizeof(double),
                                   count * sizeof(double), nsteps, cudaMemcpyDeviceToHost, e->stream));
        CUDA_TRY(cudaStreamSynchronize(e->stream));
        float rms = 0.f;
        cudaEventElapsedTime(&rms, e->ev0, e->ev1);
        e->last_ms = rms;
        return check_nan(e);
    }
    for (long long s = 0; s < nsteps; ++s) {
        for (int half = 0; half < 2; ++half) {
            MoveDev mv;
            fill_move(e, half, 1, mv);
            rc = launch_pass(e->p, mv, e->stream, &e->last_launches);
            if (rc) return rc;
        }
        CUDA_TRY(cudaEventRecord(e->ev_step, e->stream));
        CUDA_TRY(cudaStreamWaitEvent(e->copy_stream, e->ev_step, 0));
        CUDA_TRY(cudaMemcpyAsync(chain_host + s * count * D, e->d_chain + e->nstored * cw + rel * D, sizeof(double) * count * D,
                                 cudaMemcpyDeviceToHost, e->copy_stream));
        CUDA_TRY(cudaMemcpyAsync(log_prob_host + s * count, e->d_lnp + e->nstored * lw + rel, sizeof(double) * count,
                                 cudaMemcpyDeviceToHost, e->copy_stream));
        e->iteration += 1;
        e->nstored += 1;
    }
    CUDA_TRY(cudaEventRecord(e->ev1, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->copy_stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->last_ms = ms;
    return check_nan(e);
}

int lcf_ensemble_run_to_host(lcf_ensemble *e, int64_t nsteps, double *chain_host, double *log_prob_host) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (e->cw != e->W) return fail(LCF_ERR_ARG, "this rank stores only its own walkers: use lcf_ensemble_run_to_host_slice");
    return lcf_ensemble_run_to_host_slice(e, nsteps, 0, e->W, chain_host, log_prob_host);
}

int lcf_ensemble_run_replay(lcf_ensemble *e, int64_t nsteps, int store, const int32_t *split, const double *z,
                            const int32_t *partner, const double *logu) {
    if (!e || !split || !z || !partner || !logu) return fail(LCF_ERR_ARG, "null argument");
    if (!e->has_state) return fail(LCF_ERR_STATE, "run_mcmc before an initial state was set");
    if (e->world != 1) return fail(LCF_ERR_ARG, "replay mode is single-GPU");
    CUDA_TRY(cudaSetDevice(e->p->device));
    const long long W = e->W;
    if (store) {
        int rc = ensure_capacity(e, e->nstored + nsteps);
        if (rc) return rc;
    }
    // physical rows of the split-0 walkers (ascending walker index) then of the split-1 walkers, per step
    std::vector<int> rows((size_t)nsteps * W), ns0(nsteps);
    for (long long s = 0; s < nsteps; ++s) {
        long long k = 0;
        for (int sp = 0; sp < 2; ++sp) {
            for (long long j = 0; j < W; ++j) {
                int v = split[s * W + j];
                if (v != 0 && v != 1) return fail(LCF_ERR_ARG, "split labels must be 0 or 1");
                if (v == sp) rows[s * W + k++] = (int)phys_row(e, j);
            }
            if (sp == 0) ns0[s] = (int)k;
        }
        if (ns0[s] == 0 || ns0[s] == W) return fail(LCF_ERR_ARG, "degenerate split");
    }
    int *d_rows = nullptr, *d_part = nullptr;
    double *d_z = nullptr, *d_lu = nullptr;
    CUDA_TRY(cudaMalloc(&d_rows, sizeof(int) * nsteps * W));
    CUDA_TRY(cudaMalloc(&d_part, sizeof(int) * nsteps * W));
    CUDA_TRY(cudaMalloc(&d_z, sizeof(double) * nsteps * W));
    CUDA_TRY(cudaMalloc(&d_lu, sizeof(double) * nsteps * W));
    CUDA_TRY(cudaMemcpy(d_rows, rows.data(), sizeof(int) * nsteps * W, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_part, partner, sizeof(int) * nsteps * W, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_z, z, sizeof(double) * nsteps * W, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_lu, logu, sizeof(double) * nsteps * W, cudaMemcpyHostToDevice));
    e->last_launches = 0;
    int rc = 0;
    cudaEventRecord(e->ev0, e->stream);
    for (long long s = 0; s < nsteps && !rc; ++s) {
        for (int half = 0; half < 2 && !rc; ++half) {
            MoveDev mv;
            fill_move(e, half, store, mv);
            const long long off = s * W + (half ? ns0[s] : 0);
            mv.Ns = half ? W - ns0[s] : ns0[s];
            mv.Nc = W - mv.Ns;
            mv.act_rows = d_rows + off;
            mv.comp_rows = d_rows + s * W + (half ? 0 : ns0[s]);
            mv.zin = d_z + off;
            mv.rin = d_part + off;
            mv.luin = d_lu + off;
            rc = launch_pass(e->p, mv, e->stream, &e->last_launches);
        }
        e->iteration += 1;
        if (store) e->nstored += 1;
    }
    cudaEventRecord(e->ev1, e->stream);
    cudaError_t ce = cudaStreamSynchronize(e->stream);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->last_ms = ms;
    cudaFree(d_rows); cudaFree(d_part); cudaFree(d_z); cudaFree(d_lu);
    if (rc) return rc;
    if (ce != cudaSuccess) return fail(LCF_ERR_CUDA, "replay failed: %s", cudaGetErrorString(ce));
    return check_nan(e);
}

int64_t lcf_ensemble_nstored(lcf_ensemble *e) { return e ? e->nstored : 0; }

int lcf_ensemble_get_chain(lcf_ensemble *e, double *chain) {
    if (!e || !chain) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    // [nstored][W][D] on the host; a rank of a shared ensemble fills the columns of its own walkers only
    if (e->nstored)
        CUDA_TRY(cudaMemcpy2D(chain + e->cfirst * e->D, sizeof(double) * e->W * e->D, e->d_chain, sizeof(double) * e->cw * e->D,
                              sizeof(double) * e->cw * e->D, e->nstored, cudaMemcpyDeviceToHost));
    return 0;
}
int lcf_ensemble_get_log_prob(lcf_ensemble *e, double *log_prob) {
    if (!e || !log_prob) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (e->nstored)
        CUDA_TRY(cudaMemcpy2D(log_prob + e->cfirst, sizeof(double) * e->W, e->d_lnp, sizeof(double) * e->cw, sizeof(double) * e->cw,
                              e->nstored, cudaMemcpyDeviceToHost));
    return 0;
}
int lcf_ensemble_get_chain_slice(lcf_ensemble *e, int64_t first, int64_t count, double *chain, double *log_prob) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (first < e->cfirst || count < 0 || first + count > e->cfirst + e->cw)
        return fail(LCF_ERR_ARG, "walker range outside the walkers this rank stores");
    CUDA_TRY(cudaSetDevice(e->p->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (!e->nstored || !count) return 0;
    const size_t D = e->D, rel = (size_t)(first - e->cfirst);
    if (chain)
        CUDA_TRY(cudaMemcpy2D(chain, count * D * sizeof(double), e->d_chain + rel * D, e->cw * D * sizeof(double),
                              count * D * sizeof(double), e->nstored, cudaMemcpyDeviceToHost));
    if (log_prob)
        CUDA_TRY(cudaMemcpy2D(log_prob, count * sizeof(double), e->d_lnp + rel, e->cw * sizeof(double), count * sizeof(double),
                              e->nstored, cudaMemcpyDeviceToHost));
    return 0;
}
int lcf_ensemble_get_accepted(lcf_ensemble *e, int64_t *accepted) {
    if (!e || !accepted) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    CUDA_TRY(cudaMemcpy(accepted, e->d_acc, sizeof(unsigned long long) * e->W, cudaMemcpyDeviceToHost));
    return 0;
}
// ---- convergence diagnostics on the device-resident chain ----------------------------------
// tau[d]: integrated autocorrelation time with emcee's definition (emcee.autocorr.integrated_time: normalised ACF of every
// walker averaged over walkers, tau(M) = 2 sum_{k<=M} f(k) - 1, Sokal's automatic window = first M with M >= c tau(M)),
// window[d]: that M (or -1 when no window exists inside max_lag: the estimate is then taken at the last lag); walkers that
// never moved in the range (exactly constant series, for which emcee returns NaN) are left out of the average;
// rhat[d]: split Gelman-Rubin statistic over the 2 W half-chains.  Steps [discard, nstored) are used.
int lcf_ensemble_diagnostics(lcf_ensemble *e, int64_t discard, double c, int64_t max_lag, double *tau, int64_t *window,
                             double *rhat) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (e->world != 1) return fail(LCF_ERR_ARG, "diagnostics need the whole chain on one GPU");
    const long long n = e->nstored - discard, W = e->W;
    const int D = e->D;
    if (discard < 0 || n < 4) return fail(LCF_ERR_STATE, "need at least 4 stored steps after discard");
    if (!(c > 0.)) return fail(LCF_ERR_ARG, "c must be positive");
    CUDA_TRY(cudaSetDevice(e->p->device));
    long long nlag = n;
    if (max_lag > 0) nlag = std::min<long long>(n, max_lag + 1);
    double *d_mom = nullptr, *d_f = nullptr;
    CUDA_TRY(cudaMalloc(&d_mom, sizeof(double) * W * D * 6));
    CUDA_TRY(cudaMalloc(&d_f, sizeof(double) * D * nlag));
    const long long ns = W * D;
    k_series_moments<<<(unsigned)((ns + 255) / 256), 256, 0, e->stream>>>(e->d_chain, discard, n, W, D, d_mom);
    if (tau || window) {
        dim3 grid((unsigned)((nlag + kLagBlock - 1) / kLagBlock), (unsigned)D);
        k_mean_acf<<<grid, 256, 0, e->stream>>>(e->d_chain, discard, n, W, D, d_mom, nlag, d_f);
    }
    cudaError_t ce = cudaGetLastError();
    std::vector<double> mom((size_t)ns * 6), f((size_t)D * nlag);
    if (ce == cudaSuccess) ce = cudaMemcpyAsync(mom.data(), d_mom, sizeof(double) * mom.size(), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess && (tau || window))
        ce = cudaMemcpyAsync(f.data(), d_f, sizeof(double) * f.size(), cudaMemcpyDeviceToHost, e->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->stream);
    cudaFree(d_mom);
    cudaFree(d_f);
    if (ce != cudaSuccess) return fail(LCF_ERR_CUDA, "diagnostics failed: %s", cudaGetErrorString(ce));
    if (tau || window) {
        for (int d = 0; d < D; ++d) {
            long long moving = 0;                                       // walkers that contributed (stuck ones are left out)
            for (long long w = 0; w < W; ++w) moving += mom[(size_t)(w * D + d) * 6 + 1] > 0.;
            const double renorm = moving ? (double)W / (double)moving : NAN;
            double cum = 0., t_at = 0.;
            long long win = -1;
            for (long long k = 0; k < nlag; ++k) {
                cum += f[(size_t)d * nlag + k] * renorm;
                const double tk = 2. * cum - 1.;
                t_at = tk;
                if (!((double)k < c * tk)) { win = k; break; }        // emcee auto_window: first k with k >= c tau(k)
            }
            if (tau) tau[d] = t_at;
            if (window) window[d] = win;
        }
    }
    if (rhat) {
        const long long h = n / 2;
        for (int d = 0; d < D; ++d) {
            // 2 W chains of h steps: B/h = variance of the chain means, Wv = mean within-chain variance
            double mm = 0., wv = 0.;
            for (long long w = 0; w < W; ++w) {
                const double *o = &mom[(size_t)(w * D + d) * 6];
                mm += o[2] + o[4];
                wv += (o[3] + o[5]) / (double)(h - 1);
            }
            const double M = 2. * (double)W;
            mm /= M;
            wv /= M;
            double b = 0.;
            for (long long w = 0; w < W; ++w) {
                const double *o = &mom[(size_t)(w * D + d) * 6];
                b += (o[2] - mm) * (o[2] - mm) + (o[4] - mm) * (o[4] - mm);
            }
            b /= (M - 1.);                                             // variance of the means = B / h
            const double var_plus = (double)(h - 1) / (double)h * wv + b;
            rhat[d] = std::sqrt(var_plus / wv);
        }
    }
    return 0;
}

// ---- batched blackbody least squares (replaces the per-epoch scipy curve_fit of bolometric.py:483-531) ----
int lcf_blackbody_lstsq_batch(int64_t nepochs, const int32_t *offsets, const double *nu, const double *lum, double c1, double c2,
                              double cutoff_freq, const double *p0, const double *lower, const double *upper, double *popt,
                              double *pcov, int32_t *status) {
    if (!offsets || !nu || !lum || !p0 || !lower || !upper || !popt || !pcov || !status) return fail(LCF_ERR_ARG, "null argument");
    if (nepochs <= 0) return 0;
    int rc = check_device();
    if (rc) return rc;
    const long long npts = offsets[nepochs];
    for (long long e = 0; e < nepochs; ++e)
        if (offsets[e + 1] < offsets[e] + 1) return fail(LCF_ERR_ARG, "epoch %lld has no photometry point", e);
    int *d_off = nullptr, *d_st = nullptr;
    double *d_nu = nullptr, *d_lum = nullptr, *d_p = nullptr, *d_c = nullptr;
    cudaError_t ce = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; return ce == cudaSuccess; };
    ok(cudaMalloc(&d_off, sizeof(int) * (nepochs + 1))) && ok(cudaMalloc(&d_st, sizeof(int) * nepochs)) &&
        ok(cudaMalloc(&d_nu, sizeof(double) * npts)) && ok(cudaMalloc(&d_lum, sizeof(double) * npts)) &&
        ok(cudaMalloc(&d_p, sizeof(double) * 2 * nepochs)) && ok(cudaMalloc(&d_c, sizeof(double) * 4 * nepochs)) &&
        ok(cudaMemcpy(d_off, offsets, sizeof(int) * (nepochs + 1), cudaMemcpyHostToDevice)) &&
        ok(cudaMemcpy(d_nu, nu, sizeof(double) * npts, cudaMemcpyHostToDevice)) &&
        ok(cudaMemcpy(d_lum, lum, sizeof(double) * npts, cudaMemcpyHostToDevice));
    if (ce == cudaSuccess) {
        k_bb_lstsq<<<(unsigned)((nepochs + 63) / 64), 64>>>(nepochs, d_off, d_nu, d_lum, c1, c2, cutoff_freq, p0[0], p0[1], lower[0],
                                                            upper[0], lower[1], upper[1], d_p, d_c, d_st);
        ok(cudaGetLastError()) && ok(cudaMemcpy(popt, d_p, sizeof(double) * 2 * nepochs, cudaMemcpyDeviceToHost)) &&
            ok(cudaMemcpy(pcov, d_c, sizeof(double) * 4 * nepochs, cudaMemcpyDeviceToHost)) &&
            ok(cudaMemcpy(status, d_st, sizeof(int) * nepochs, cudaMemcpyDeviceToHost));
    }
    cudaFree(d_off); cudaFree(d_st); cudaFree(d_nu); cudaFree(d_lum); cudaFree(d_p); cudaFree(d_c);
    if (ce != cudaSuccess) return fail(LCF_ERR_CUDA, "blackbody least squares failed: %s", cudaGetErrorString(ce));
    return 0;
}

int lcf_ensemble_device_view(lcf_ensemble *e, void **d_coords, void **d_log_prob, void **stream, int64_t *n0, int64_t *own_begin,
                             int64_t *own_count) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (d_coords) *d_coords = e->d_coords;
    if (d_log_prob) *d_log_prob = e->d_logp;
    if (stream) *stream = e->stream;
    if (n0) *n0 = e->n0;
    if (own_begin) { own_begin[0] = e->own_begin[0]; own_begin[1] = e->own_begin[1]; }
    if (own_count) { own_count[0] = e->own_count[0]; own_count[1] = e->own_count[1]; }
    return 0;
}
int lcf_ensemble_set_stream(lcf_ensemble *e, void *stream) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
    e->stream = reinterpret_cast<cudaStream_t>(stream);
    e->own_stream = false;
    return 0;
}
int lcf_ensemble_sync(lcf_ensemble *e) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(e->p->device));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return check_nan(e);
}
int lcf_ensemble_last_timing(lcf_ensemble *e, double *ms, int64_t *launches) {
    if (!e) return fail(LCF_ERR_ARG, "null argument");
    if (ms) *ms = e->last_ms;
    if (launches) *launches = e->last_launches;
    return 0;
}

// ---------------------------------------------------------------------------------------
// batch
// ---------------------------------------------------------------------------------------
int lcf_batch_create(int64_t nproblems, lcf_problem *const *problems, int64_t nwalkers, uint64_t seed, lcf_batch **out) {
    if (!problems || !out || nproblems <= 0) return fail(LCF_ERR_ARG, "bad argument");
    *out = nullptr;
    lcf_problem *p0 = problems[0];
    if (!p0) return fail(LCF_ERR_ARG, "null problem");
    for (long long i = 0; i < nproblems; ++i) {
        lcf_problem *p = problems[i];
        if (!p) return fail(LCF_ERR_ARG, "null problem");
        if (p->dev.model != p0->dev.model || p->precision != p0->precision || p->dev.ndim != p0->dev.ndim ||
            p->dev.use_sigma != p0->dev.use_sigma || p->device != p0->device)
            return fail(LCF_ERR_ARG, "all problems of a batch must share model, precision, ndim, use_sigma and device");
    }
    if (nwalkers < 2LL * p0->dev.ndim)
        return fail(LCF_ERR_NWALKERS, "It is unadvisable to use a red-blue move with fewer walkers than twice the number of dimensions.");
    int rc = check_device();
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(p0->device));
    lcf_batch *b = new lcf_batch();
    b->probs.assign(problems, problems + nproblems);
    b->nprob = nproblems;
    b->W = nwalkers;
    b->n0 = (nwalkers + 1) / 2;
    b->D = p0->dev.ndim;
    b->model = p0->dev.model;
    b->precision = p0->precision;
    b->seed = seed;
    // shape: walkers per CTA pass = smallest power of two covering a half-ensemble: <= 32 (narrow groups), or 64..256
    // ("wide" groups: wpb/32 walker columns of warps, one proposal phase for the whole half-ensemble) when the model
    // has no per-walker weight table and the per-CTA tables still let four CTAs share an SM
    // Look-ahead rounds (k_chain): when the n0 + 2 n1 virtual walkers of a step fit ONE narrow group (W <= 21), a step is one
    // log-posterior pass plus two accept phases instead of two dependent half-steps; LCF_CHAIN_LA=0 keeps the half-steps.
    const long long n1 = nwalkers - b->n0, NV = b->n0 + 2 * n1;
    {
        const char *env = getenv("LCF_CHAIN_LA");
        b->spec = (n1 > 0 && NV <= 32 && g_tune_wpb <= 0 && !(env && env[0] == '0')) ? 1 : 0;
    }
    int l = 0;
    if (g_tune_wpb > 0) { while ((1 << l) < g_tune_wpb && l < 5) ++l; }
    else { while ((1 << l) < (b->spec ? NV : b->n0) && l < 8) ++l; }
    if (p0->dev.model == 3) l = std::min(l, 5);
    size_t smem = 0;
    int max_tiles = 1;
    for (;;) {
        smem = 0;
        for (lcf_problem *p : b->probs) smem = std::max(smem, smem_bytes(p, 1 << l, 8, 1));
        // four (wide groups) / two CTAs per SM, counting the 1 KB the driver reserves per CTA and the kernel's static 0.8 KB
        if (smem + 2048 <= (l > 5 ? kSmemMax / 4 : kSmemMax / 2) || l == 0) break;
        --l;
    }
    if (b->spec && (1 << l) < NV) {                       // the group no longer holds the whole step: back to half-steps
        b->spec = 0;
        l = 0;
        while ((1 << l) < b->n0 && l < 8) ++l;
    }
    // split-K for narrow groups of small problems (an SED epoch has 3-9 points, one per filter): with 2^ks lanes per (walker, point
    // pair) a warp is full of short loops instead of a quarter full of long ones.  Chosen from the mean number of points that
    // share a filter: a tile of 2 * 32 / (wpb 2^ks) points should still be about full.
    int ks = 0;
    if (l < 5) {
        double runs = 0., pts = 0.;
        for (lcf_problem *p : b->probs) {
            const int N = p->dev.npoints;
            for (int i = 0; i < N; ++i) runs += (i == 0 || p->h_point_filter[i] != p->h_point_filter[i - 1]);
            pts += N;
        }
        const double per_filter = pts / std::max(1., runs);
        while (l + ks < 5 && (double)(32 >> (l + ks)) >= per_filter) ++ks;     // 2 * (32 >> (l + ks + 1)) >= points per filter
        if (g_tune_ks >= 0) ks = std::min(g_tune_ks, 5 - l);
    }
    b->ks = ks;
    std::vector<ProblemDev> hp(nproblems);
    std::vector<TileDev> ht(nproblems);
    const int lt = std::min(l, 5) + ks;                  // wide groups use the 32-walker tiles (two points per lane)
    for (long long i = 0; i < nproblems; ++i) max_tiles = std::max(max_tiles, count_tiles(b->probs[i], lt));
    int nw = g_tune_nw > 0 ? g_tune_nw : std::min(8, max_tiles);
    nw = std::max(1, std::min(nw, 8));                    // k_chain is compiled for <= 256 threads
    if (l > 5) nw = 8;                                    // wide groups: 2^(l-5) walker columns of warps must divide nw
    const int nstripes = l > 5 ? nw >> (l - 5) : nw;      // warps that share the tiles of one walker column: the dealing period
    for (long long i = 0; i < nproblems; ++i) {
        if ((rc = get_tiles(b->probs[i], lt, nstripes, &ht[i]))) { delete b; return rc; }
        hp[i] = b->probs[i]->dev;
    }
    smem = 0;
    for (lcf_problem *p : b->probs) smem = std::max(smem, smem_bytes(p, 1 << l, nw, 1));
    if (smem > kSmemMax) { delete b; return fail(LCF_ERR_ARG, "filter bank does not fit in shared memory"); }
    b->wpb_log2 = l;
    b->nw = nw;
    b->smem = smem;
    if (getenv("LCF_DEBUG_SHAPE"))
        fprintf(stderr, "[lcf] batch shape: %lld problems, %lld walkers, %d walkers/pass, %d warps, split-K %d, %zu B smem\n", (long long)nproblems,
                (long long)nwalkers, 1 << l, nw, 1 << b->ks, smem);
    void *dp;
    if ((rc = upload(hp, &dp))) { delete b; return rc; }
    b->d_probs = reinterpret_cast<ProblemDev *>(dp);
    if ((rc = upload(ht, &dp))) { delete b; return rc; }
    b->d_tiles = reinterpret_cast<TileDev *>(dp);
    {   // every CTA runs a whole chain, so the grid's makespan is a scheduling problem: longest processing time first
        std::vector<int> order(nproblems);
        std::vector<double> work(nproblems);
        for (long long i = 0; i < nproblems; ++i) {
            order[i] = (int)i;
            work[i] = (double)b->probs[i]->dev.npoints * std::max(1., b->probs[i]->mean_samples);
        }
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return work[x] > work[y]; });
        if ((rc = upload(order, &dp))) { delete b; return rc; }
        b->d_order = reinterpret_cast<int *>(dp);
    }
    cudaError_t ce;
    if ((ce = cudaMalloc(&b->d_coords, sizeof(double) * nproblems * b->W * b->D)) != cudaSuccess ||
        (ce = cudaMalloc(&b->d_logp, sizeof(double) * nproblems * b->W)) != cudaSuccess ||
        (ce = cudaMalloc(&b->d_acc, sizeof(unsigned long long) * nproblems * b->W)) != cudaSuccess ||
        (ce = cudaMalloc(&b->d_status, sizeof(int) * nproblems)) != cudaSuccess ||
        (ce = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (ce = cudaEventCreate(&b->ev0)) != cudaSuccess || (ce = cudaEventCreate(&b->ev1)) != cudaSuccess) {
        delete b;
        return fail(LCF_ERR_CUDA, "batch allocation failed: %s", cudaGetErrorString(ce));
    }
    cudaMemset(b->d_acc, 0, sizeof(unsigned long long) * nproblems * b->W);
    cudaMemset(b->d_status, 0, sizeof(int) * nproblems);
    *out = b;
    return 0;
}

void lcf_batch_destroy(lcf_batch *b) { delete b; }

int lcf_batch_set_state(lcf_batch *b, const double *coords) {
    if (!b || !coords) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(b->probs[0]->device));
    const int D = b->D;
    const long long W = b->W;
    std::vector<double> hc((size_t)b->nprob * W * D);
    for (long long p = 0; p < b->nprob; ++p)
        for (long long j = 0; j < W; ++j) {
            long long r = (j & 1) ? b->n0 + (j >> 1) : (j >> 1);
            for (int d = 0; d < D; ++d) hc[(p * W + r) * D + d] = coords[(p * W + j) * D + d];
        }
    CUDA_TRY(cudaMemcpy(b->d_coords, hc.data(), sizeof(double) * hc.size(), cudaMemcpyHostToDevice));
    b->has_state = true;
    b->need_init_logp = true;
    return 0;
}

int lcf_batch_run(lcf_batch *b, int64_t nburn, int64_t nsteps) {
    NvtxRange r("lcf_batch_run");
    if (!b) return fail(LCF_ERR_ARG, "null argument");
    if (!b->has_state) return fail(LCF_ERR_STATE, "batch has no initial state");
    if (nburn < 0 || nsteps < 0) return fail(LCF_ERR_ARG, "negative step count");
    CUDA_TRY(cudaSetDevice(b->probs[0]->device));
    const size_t need_chain = std::max<size_t>(16, sizeof(double) * b->nprob * nsteps * b->W * b->D);
    const size_t need_lnp = std::max<size_t>(16, sizeof(double) * b->nprob * nsteps * b->W);
    if (need_chain > b->chain_cap) {
        cudaFree(b->d_chain); b->d_chain = nullptr; b->chain_cap = 0;
        CUDA_TRY(cudaMalloc(&b->d_chain, need_chain));
        b->chain_cap = need_chain;
    }
    if (need_lnp > b->lnp_cap) {
        cudaFree(b->d_lnp); b->d_lnp = nullptr; b->lnp_cap = 0;
        CUDA_TRY(cudaMalloc(&b->d_lnp, need_lnp));
        b->lnp_cap = need_lnp;
    }
    CUDA_TRY(cudaMemsetAsync(b->d_acc, 0, sizeof(unsigned long long) * b->nprob * b->W, b->stream));
    BatchDev B;
    memset(&B, 0, sizeof(B));
    B.probs = b->d_probs; B.tiles = b->d_tiles; B.order = b->d_order;
    B.coords = b->d_coords; B.logp = b->d_logp; B.accepted = b->d_acc; B.status = b->d_status;
    B.chain = b->d_chain; B.lnp = b->d_lnp;
    B.W = b->W; B.n0 = b->n0; B.nproblems = b->nprob;
    B.nburn = nburn; B.nsteps = nsteps; B.iter0 = b->iteration;
    B.seed = b->seed; B.wpb_log2 = b->wpb_log2; B.ks = b->ks; B.init_logp = b->need_init_logp ? 1 : 0;
    if (!b->d_spec) CUDA_TRY(cudaMalloc(&b->d_spec, sizeof(double) * (32 * (size_t)b->nprob + 1)));
    B.spec = b->spec; B.spec_nlp = b->d_spec; B.nan_scratch = reinterpret_cast<int *>(b->d_spec + 32 * (size_t)b->nprob);
    ChainKernel k = (b->precision == LCF_PRECISION_FP32) ? chain_kernel_for<float>(b->model) : chain_kernel_for<double>(b->model);
    if (!k) return fail(LCF_ERR_ARG, "unknown model");
    { int rc = ensure_dynamic_smem(reinterpret_cast<const void *>(k), b->smem); if (rc) return rc; }
    CUDA_TRY(cudaEventRecord(b->ev0, b->stream));
    k<<<(unsigned)b->nprob, b->nw * 32, b->smem, b->stream>>>(B);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(b->ev1, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, b->ev0, b->ev1);
    b->last_ms = ms;
    b->last_launches = 1;
    b->iteration += nburn + nsteps;
    b->nsteps_stored = nsteps;
    b->need_init_logp = false;
    return 0;
}

// ---- a whole table of SED epochs as ONE batch (calculate_bolometric at survey scale) ---------------------
// Flat arrays in, one handle out: per-epoch problems are built in C (points grouped by filter, sub-bank gathered from the shared
// filter bank, sigma units = median dy as bolometric.py:147-150), their device arrays share ONE allocation filled by ONE copy.
int lcf_sed_batch_create(int64_t nepochs, const int32_t *offsets, const int32_t *point_filter, const double *y, const double *dy,
                         int32_t nfilters, const int32_t *bank_offsets, const double *bank_alpha, const double *bank_w, int32_t ndim,
                         int32_t use_sigma, int32_t sigma_type, const int32_t *prior_kind, const double *prior_min,
                         const double *prior_max, const double *prior_mean, const double *prior_std, int32_t precision,
                         int64_t nwalkers, uint64_t seed, lcf_batch **out) {
    if (!offsets || !point_filter || !y || !dy || !bank_offsets || !bank_alpha || !bank_w || !prior_kind || !prior_min || !prior_max || !out)
        return fail(LCF_ERR_ARG, "null argument");
    *out = nullptr;
    if (nepochs <= 0) return fail(LCF_ERR_ARG, "no epochs");
    std::vector<lcf_problem *> probs;
    auto cleanup = [&]() { for (lcf_problem *p : probs) delete p; };
    std::vector<int> order, local, sub_off, pf;
    std::vector<double> sub_alpha, sub_w, sub_kappa, t0, yy, dd, med;
    for (int64_t e = 0; e < nepochs; ++e) {
        const int b0 = offsets[e], n = offsets[e + 1] - b0;
        if (n <= 0) { cleanup(); return fail(LCF_ERR_ARG, "epoch %lld has no photometry point", (long long)e); }
        order.resize(n);
        for (int i = 0; i < n; ++i) {
            order[i] = i;
            if (point_filter[b0 + i] < 0 || point_filter[b0 + i] >= nfilters) { cleanup(); return fail(LCF_ERR_ARG, "point_filter out of range"); }
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return point_filter[b0 + a] < point_filter[b0 + b]; });
        sub_off.assign(1, 0); sub_alpha.clear(); sub_w.clear(); pf.resize(n); yy.resize(n); dd.resize(n);
        int last = -1, nf = 0;
        for (int i = 0; i < n; ++i) {
            const int g = point_filter[b0 + order[i]];
            if (g != last) {
                sub_alpha.insert(sub_alpha.end(), bank_alpha + bank_offsets[g], bank_alpha + bank_offsets[g + 1]);
                sub_w.insert(sub_w.end(), bank_w + bank_offsets[g], bank_w + bank_offsets[g + 1]);
                sub_off.push_back((int)sub_alpha.size());
                last = g; ++nf;
            }
            pf[i] = nf - 1;
            yy[i] = y[b0 + order[i]];
            dd[i] = dy[b0 + order[i]];
        }
        med.assign(dd.begin(), dd.end());
        std::sort(med.begin(), med.end());
        t0.assign(n, 0.);
        sub_kappa.assign(sub_alpha.size(), 0.);
        lcf_problem_desc d;
        memset(&d, 0, sizeof(d));
        d.model_id = LCF_MODEL_BLACKBODY_SED;
        d.precision = precision; d.ndim = ndim; d.use_sigma = use_sigma; d.sigma_type = sigma_type;
        d.npoints = n; d.nfilters = nf;
        d.sigma_unit_abs = (n & 1) ? med[n / 2] : 0.5 * (med[n / 2 - 1] + med[n / 2]);      // np.median(dy)
        d.bank_offsets = sub_off.data(); d.bank_alpha = sub_alpha.data(); d.bank_w = sub_w.data(); d.bank_kappa = sub_kappa.data();
        d.t = t0.data(); d.point_filter = pf.data(); d.y = yy.data(); d.dy = dd.data();
        d.prior_kind = prior_kind; d.prior_min = prior_min; d.prior_max = prior_max; d.prior_mean = prior_mean; d.prior_std = prior_std;
        lcf_problem *p = nullptr;
        int rc = problem_create_impl(&d, &p, true);
        if (rc) { cleanup(); return rc; }
        probs.push_back(p);
    }
    // one allocation, one copy for the device arrays of every epoch
    size_t total = 0;
    std::vector<size_t> base(probs.size());
    for (size_t i = 0; i < probs.size(); ++i) { base[i] = total; total += (probs[i]->pending->host.size() + 255) & ~(size_t)255; }
    std::vector<unsigned char> host(std::max<size_t>(total, 16));
    for (size_t i = 0; i < probs.size(); ++i) memcpy(host.data() + base[i], probs[i]->pending->host.data(), probs[i]->pending->host.size());
    void *dblock = nullptr;
    cudaError_t ce = cudaMalloc(&dblock, host.size());
    if (ce == cudaSuccess) ce = cudaMemcpy(dblock, host.data(), host.size(), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { cudaFree(dblock); cleanup(); return fail(LCF_ERR_CUDA, "batch upload failed: %s", cudaGetErrorString(ce)); }
    for (size_t i = 0; i < probs.size(); ++i) {
        probs[i]->pending->patch(reinterpret_cast<unsigned char *>(dblock) + base[i]);
        delete probs[i]->pending;
        probs[i]->pending = nullptr;
    }
    void *tblock = nullptr;
    int rc = build_tiles_shared(probs, &tblock);
    if (rc) { cudaFree(dblock); cleanup(); return rc; }
    lcf_batch *b = nullptr;
    rc = lcf_batch_create(nepochs, probs.data(), nwalkers, seed, &b);
    if (rc) { cudaFree(dblock); cudaFree(tblock); cleanup(); return rc; }
    b->owned = probs;
    b->d_shared = dblock;
    b->d_shared_tiles = tblock;
    *out = b;
    return 0;
}

// Posterior summaries of every epoch of a SED batch on the device (bolometric.py:792-798): T, R, Stefan-Boltzmann L_bol and the
// pseudo-bolometric luminosity of every stored sample (`comb`: the 1-THz comb problem of pseudo(), evaluated IN PLACE on the
// HBM-resident chain), then median_and_unc of each: out[nproblems][4][3] = (median, median - lower, upper - median).
int lcf_batch_summary(lcf_batch *b, lcf_problem *comb, double sigma_sb, double perc_contained, double *out) {
    if (!b || !comb || !out) return fail(LCF_ERR_ARG, "null argument");
    if (comb->dev.model != LCF_MODEL_BLACKBODY_SED || comb->dev.npoints != 1) return fail(LCF_ERR_ARG, "comb must be a one-point blackbody problem");
    if (b->model != LCF_MODEL_BLACKBODY_SED) return fail(LCF_ERR_ARG, "summary is defined for blackbody SED batches");
    const long long n = b->nsteps_stored * b->W, total = n * b->nprob;
    if (n < 1) return fail(LCF_ERR_STATE, "no stored samples");
    CUDA_TRY(cudaSetDevice(b->probs[0]->device));
    int npad = 1;
    while (npad < n) npad <<= 1;
    const size_t smem = sizeof(double) * (size_t)npad;
    if (smem > kSmemMax) return fail(LCF_ERR_ARG, "more than %zu samples per epoch: take the percentiles of get_chain() on the host", kSmemMax / 8);
    if ((size_t)total > b->lpseudo_cap) {
        cudaFree(b->d_lpseudo); b->d_lpseudo = nullptr; b->lpseudo_cap = 0;
        CUDA_TRY(cudaMalloc(&b->d_lpseudo, sizeof(double) * total));
        b->lpseudo_cap = (size_t)total;
    }
    if (!b->d_summary) CUDA_TRY(cudaMalloc(&b->d_summary, sizeof(double) * 12 * b->nprob));
    if (!comb->d_eval_nan) CUDA_TRY(cudaMalloc(&comb->d_eval_nan, sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(comb->d_eval_nan, 0, sizeof(int), b->stream));
    CUDA_TRY(cudaMemsetAsync(b->d_lpseudo, 0, sizeof(double) * total, b->stream));
    MoveDev mv;
    memset(&mv, 0, sizeof(mv));
    mv.mode = MODE_MODEL;
    mv.Ns = total;
    mv.qin = b->d_chain;
    mv.qstride = b->D;
    mv.out = b->d_lpseudo;
    mv.nanflag = comb->d_eval_nan;
    int rc = launch_pass(comb, mv, b->stream, nullptr);
    if (rc) return rc;
    rc = ensure_dynamic_smem(reinterpret_cast<const void *>(k_batch_summary), smem);
    if (rc) return rc;
    k_batch_summary<<<(unsigned)b->nprob, 256, smem, b->stream>>>(b->d_chain, b->d_lpseudo, n, b->D, npad, sigma_sb, perc_contained, b->d_summary);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, b->d_summary, sizeof(double) * 12 * b->nprob, cudaMemcpyDeviceToHost, b->stream));
    CUDA_TRY(cudaStreamSynchronize(b->stream));
    return 0;
}

int lcf_batch_get_chain(lcf_batch *b, double *chain) {
    if (!b || !chain) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(b->probs[0]->device));
    if (b->nsteps_stored)
        CUDA_TRY(cudaMemcpy(chain, b->d_chain, sizeof(double) * b->nprob * b->nsteps_stored * b->W * b->D, cudaMemcpyDeviceToHost));
    return 0;
}
int lcf_batch_get_log_prob(lcf_batch *b, double *log_prob) {
    if (!b || !log_prob) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(b->probs[0]->device));
    if (b->nsteps_stored)
        CUDA_TRY(cudaMemcpy(log_prob, b->d_lnp, sizeof(double) * b->nprob * b->nsteps_stored * b->W, cudaMemcpyDeviceToHost));
    return 0;
}
int lcf_batch_get_accepted(lcf_batch *b, int64_t *accepted) {
    if (!b || !accepted) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(b->probs[0]->device));
    CUDA_TRY(cudaMemcpy(accepted, b->d_acc, sizeof(unsigned long long) * b->nprob * b->W, cudaMemcpyDeviceToHost));
    return 0;
}
int lcf_batch_get_status(lcf_batch *b, int32_t *status) {
    if (!b || !status) return fail(LCF_ERR_ARG, "null argument");
    CUDA_TRY(cudaSetDevice(b->probs[0]->device));
    std::vector<int> h(b->nprob);
    CUDA_TRY(cudaMemcpy(h.data(), b->d_status, sizeof(int) * b->nprob, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < b->nprob; ++i) status[i] = h[i] ? LCF_ERR_NAN : 0;
    return 0;
}
int lcf_batch_last_timing(lcf_batch *b, double *ms, int64_t *launches) {
    if (!b) return fail(LCF_ERR_ARG, "null argument");
    if (ms) *ms = b->last_ms;
    if (launches) *launches = b->last_launches;
    return 0;
}

}  // extern "C"
