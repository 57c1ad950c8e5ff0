// C-ABI of libsst_b200.so (declared in include/sst_b200.h): context, device tables, kernel launches.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <cmath>
#include <new>
#include <vector>

#include "../../include/sst_b200.h"
#include "sst_common.cuh"
#include "sst_direct.cuh"
#include "sst_enum.cuh"
#include "sst_explain.cuh"
#include "sst_ladder.cuh"
#include "sst_table.cuh"

using namespace sst;

namespace {

constexpr int kMaxPending = 32;
constexpr size_t kReplaySmemLimit = 200 * 1024;  // dynamic shared memory a replay CTA may ask for (k_memo_phase_a)

// Host-side stopwatch of the asynchronous entries (diagnostics, sst_host_profile): wall time between marks inside the
// submitting calls, summed per section.  Off unless switched on; a mark is two loads when off.
struct HostProf {
    bool on = false;
    uint64_t ns[32] = {0}, calls[32] = {0};
    uint64_t last = 0;
};
HostProf g_hp;
cudaEvent_t g_trace_base = nullptr;  // common origin of every context's trace events
inline uint64_t host_now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ULL + (uint64_t)ts.tv_nsec;
}
inline void hp_begin() {
    if (g_hp.on) g_hp.last = host_now_ns();
}
inline void hp_mark(int k) {
    if (!g_hp.on) return;
    const uint64_t t = host_now_ns();
    g_hp.ns[k] += t - g_hp.last;
    g_hp.calls[k]++;
    g_hp.last = t;
}

struct DevBuf {  // grow-only device scratch
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct sst_table {
    uint64_t* tbl = nullptr;
    uint4* H = nullptr;
    uint32_t* d_any = nullptr;  // last-row summary, one bit per table word (k_last_row_summary)
    uint8_t* d_wbucket = nullptr;  // first row at or above every multiple of 1024 (classification's singleton test)
    int32_t* d_weights = nullptr;
    int32_t* d_step = nullptr;
    int32_t* d_shift = nullptr;
    int* d_flags = nullptr;  // one completion flag per (tile, row group)
    int R = 0;
    int64_t C = 0;
    int n_tiles = 0;
    int64_t step_min = 0;
    int64_t w_min = 0;
    int64_t w_host[128] = {0};
    uint64_t last_mask = ~0ULL;
    uint32_t leaf_mul = 0;  // collision-free multiplier of the weight -> row hash (0 = none found)
    uint32_t* d_lamq = nullptr;  // scheduling cost model (CostModel): compositions per unit of mass, Q16, per coarse mass bucket
    std::vector<uint32_t> h_lamq;  // the same on the host (the asynchronous entry estimates a batch's cost from a sample)
    uint32_t lam_width = 1, lam_K = 0;
    float build_ms = 0.f, transpose_ms = 0.f;
    bool built_here = false;
    bool masks_fused = false;  // the last build wrote H itself: launch_transpose only derives the last-row summary
    // composition-count table of the direct pass (sst_direct.cuh): built on first use, dropped when the table is rebuilt
    uint32_t* d_cnt2d = nullptr;  // [R][Mcnt]
    int64_t Mcnt = 0;
    bool cnt_ready = false;
    float count_ms = 0.f;
};

struct sst_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;  // side stream: classification that overlaps an enumeration pass
    cudaDeviceProp prop{};
    char err[512] = {0};
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;    // user stopwatch
    cudaEvent_t ev_run = nullptr;                  // end of the device work of the last enumeration pass
    cudaEvent_t kev[2 * 32] = {nullptr};            // pooled per-launch brackets
    int pending_slot[32] = {0};
    int n_pending = 0;
    cudaEvent_t tev[2] = {nullptr, nullptr};        // table build stopwatch
    float k_ms[SST_K_COUNT_] = {0};
    uint64_t k_launches[SST_K_COUNT_] = {0};
    bool time_kernels = true;
    // staged peak batch
    int64_t P = 0;
    int R_staged = 0;
    DevBuf d_target, d_thr, d_maxmods, d_mode, d_ind, d_ismod, d_memo_peaks;
    int n_memo = 0;
    int64_t window_total = 0;  // sum of window sizes (upper bound for the number of roots)
    int64_t max_hi = 0;
    // results
    DevBuf d_status, d_cnt, d_peakoff, d_recs, d_blocksums;
    DevBuf d_scan, d_vmass, d_vthrf, d_chunk_k, d_chunk_r, d_tmprecs, d_tmppeak, d_lvlcnt, d_lvlA, d_ctalvl, d_nodemask;
    DevBuf d_item_m[2], d_item_peak[2], d_item_meta[2], d_item_all[2], d_item_ind[2], d_item_path[2];
    bool has_exact = false;
    int levels = 0;
    uint64_t widest_level = 0;      // nodes of the widest level of the last run (sizes the next grid)
    uint64_t item_capacity = 0, n_items = 0;
    bool valid_f64 = false;
    double v_precision = 1e-3, v_tolerance = 1e-5;
    int64_t deepest = 0;  // longest composition any staged window value can have
    DevBuf d_memo_keys, d_memo_alive, d_memo_top, d_memo_misc, d_flush;
    DevBuf d_vtarget, d_vthr, d_vout;  // staged validity probes
    DevBuf d_cobs, d_coff, d_cout;     // staged classification batch
    DevBuf d_bkeys, d_btop, d_blower, d_bupper, d_bout;  // sequence-length bounds
    int64_t CF = 0;
    int CB = 0;
    const uint8_t* c_async_out = nullptr;  // where the last asynchronous classification lands (sst_classify_wait looks at its first row)
    bool c_async_pack4 = false;
    int64_t VP = 0;
    int* h_misc = nullptr;             // pinned: run summary read back with one copy
    unsigned long long* h_run = nullptr;  // pinned + mapped: the enumeration pass writes its summary here itself
    unsigned long long* h_run_dev = nullptr;
    unsigned long long* d_run = nullptr;  // the same summary in device memory (pipelined submissions copy it out)
    unsigned int* d_bar = nullptr;     // two grid-barrier counters (64 words apart) that alternate between launches
    unsigned run_parity = 0;
    uint64_t n_roots = 0, n_comps = 0;
    int rec_width = 0;
    bool have_result = false;
    uint64_t item_limit = ~0ULL;     // blow-up guard (items per level)
    int pass_grid_max[3] = {0, 0, 0};  // co-resident CTAs of the k_explain_pass instances on this device
    int dfs_grid_max[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // the same for the k_explain_dfs instances (two CTA shapes)
    int pass_grid_cap = 512;           // CTA-total scratch is sized for the largest grid of either pass
    int pass_choice = 0;               // sst_set_pass: 0 automatic, 1 level-synchronous, 2 depth-first (items), 3 direct (count table)
    int dir_grid_max[2] = {0, 0};      // co-resident CTAs of the k_explain_direct instances
    DevBuf d_bag_m, d_bag_meta, d_bag_cnt, d_bag_off, d_bag_path;  // CTA-private bags of the direct pass
    DevBuf d_roots, d_tiles;           // its window roots and per-tile descriptors
    uint64_t root_capacity = 0;
    int last_pass = 0;                 // which pass produced the last result
    DevBuf d_rootvp, d_rootcnt, d_tilebase, d_ctans;
    uint64_t pool_capacity = 0;        // items of the depth-first pass's pool
    DevBuf d_peakcost, d_blkcost;      // scheduling costs of the staged batch (peak_cost(), sums per kCostBlock peaks)
    DevBuf d_peakoff32;                // peak offsets as uint32 (what the asynchronous entry copies back)
    // asynchronous whole call (sst_explain_submit_f64 / sst_explain_collect)
    struct Pending {
        bool active = false, done = false;
        const double* mass = nullptr;
        const double* thr = nullptr;
        int32_t max_mods = 0;
        int64_t P = 0;
        const int32_t* ind = nullptr;
        const uint8_t* is_mod = nullptr;
        double precision = 0, tolerance = 0;
        int with_memo = 0, rec_width = 0;
        bool direct = false;
        uint8_t* status = nullptr;
        uint32_t* off32 = nullptr;
        uint8_t* recs = nullptr;
        uint64_t recs_bytes = 0, copied = 0;
        uint8_t* block = nullptr;      // the caller's result block (layout: BlockLayout)
        bool split = false;            // records as planes: uint32 lo[capN], then hp byte planes of capN each
        uint64_t capN = 0, copiedN = 0;
        int hp = 0;
    } pend;
    // N3 / N4: the classified-fragment frame that stays on the device between the rounds of the alphabet reduction
    DevBuf d_lsu, d_lobs, d_lflags, d_lalive, d_lidx, d_lreach, d_lfirst, d_lhdr, d_lkeys, d_llast, d_lcall;
    int64_t LF = -1;                   // fragments of the staged frame (-1: none)
    uint64_t l_calls = 0;              // calls of the last round
    DevBuf d_block;                    // the same block on the device: the pass writes into it, ONE copy brings it back
    uint64_t last_comps = 0;           // compositions of the last batch: sizes the speculative copy of the next one
    bool replay_attr_set = false;      // k_memo_phase_a's dynamic shared-memory limit has been raised on this device
    bool split_records = false;        // sst_set_record_split: pipelined submissions bring 8-byte records back as 4 + k byte planes
    int spec_hi_planes = 4;            // k of the next submission: the previous batch's longest composition - 4 (0 .. 4)
    bool spec_ok = true;               // sst_explain_submit_f64 queues blindly (the last batch it had to redo would have fitted, or none yet)
    int spec_rec_width = 8;            // with this record width
    cudaEvent_t trace_ev[8] = {nullptr};  // diagnostics (sst_trace_ms): device timeline of the last submitted batch
    bool trace_on = false;
    int spec_margin_pct = 2;           // records copied back blindly: the previous batch's plus this much (SST_SPEC_MARGIN_PCT);
                                       // every percent is bus time of every batch, a larger batch costs one more small copy
    bool last_split = false;           // layout of the last collected submission's records (sst_explain_rec_layout)
    uint64_t last_capN = 0;
    int last_hp = 0;
    uint64_t last_d2h_bytes = 0;       // bytes the last collected submission brought back
    int32_t up_ind[128] = {0};         // what d_ind / d_ismod hold (the per-row budgets rarely change between batches)
    uint8_t up_ismod[128] = {0};
    int up_R = -1;
    unsigned long long* h_blk = nullptr;  // pinned: block sums of the staged batch's scheduling costs
    size_t h_blk_cap = 0;
    double est_cost_per_peak = 0.0;    // mean peak_cost() of the staged batch (the automatic choice of the pass looks at it)
    bool want_cta_ns = false;          // diagnostics: per-CTA phase timestamps of the depth-first pass
    int cta_ns_grid = 0;
    uint64_t phase_ns[32] = {0};
};

namespace {

int fail(sst_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? SST_ERR_NOMEM : SST_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                      \
    } while (0)

// Non-finite inputs are found by the kernels that integerise (non_finite() in sst_explain.cuh) and reported here: the
// reference's int(round(x)) / int(np.ceil(x)) raise ValueError for NaN, OverflowError for an infinity
// (mass_explanation.py:51-58,107-114).  No pass over the host arrays: at 10^5 peaks that cost more than every CUDA call
// of a submission together.
// Result block of the asynchronous entry, the same on the device and in the caller's (pinned) memory, so that one copy
// brings everything back: [0, 352) run summary of the pass, [384, 448) summary of the staging kernel, then status[P],
// off32[P + 1] and the records, each 16-byte aligned.
struct BlockLayout {
    uint64_t status, off32, recs;
    explicit BlockLayout(int64_t P) {
        status = 512;
        off32 = status + (((uint64_t)P + 15) & ~15ULL);
        recs = off32 + ((4 * ((uint64_t)P + 1) + 15) & ~15ULL);
    }
};
constexpr uint64_t kBlockStageSummary = 384;

int nf_error(sst_ctx* ctx, unsigned long long bits) {
    if (bits & NF_NAN) return fail(ctx, SST_ERR_NAN, "cannot convert float NaN to integer");
    if (bits & NF_INF) return fail(ctx, SST_ERR_INF, "cannot convert float infinity to integer");
    return SST_OK;
}
// the same from a row of result codes: `nan_code` / `inf_code` in a byte (or, packed, in either nibble)
int nf_error_codes(sst_ctx* ctx, const uint8_t* codes, int64_t n, int nan_code, int inf_code, bool nibbles) {
    // both codes have bit 3 set, which no answer has: one OR over the row, 32 bytes at a time, says whether there is
    // anything to look for
    {
        uint64_t acc[4] = {0, 0, 0, 0};
        int64_t i = 0;
        for (; i + 32 <= n; i += 32) {
            uint64_t w[4];
            memcpy(w, codes + i, 32);
            acc[0] |= w[0]; acc[1] |= w[1]; acc[2] |= w[2]; acc[3] |= w[3];
        }
        uint64_t any = acc[0] | acc[1] | acc[2] | acc[3];
        for (; i < n; i++) any |= (uint64_t)codes[i];
        if (!(any & 0x8888888888888888ULL)) return SST_OK;
    }
    unsigned seen = 0;
    if (nibbles) {
        for (int64_t i = 0; i < n; i++) seen |= 1u << (codes[i] & 15) | 1u << (codes[i] >> 4);
    } else {
        for (int64_t i = 0; i < n; i++) seen |= 1u << (codes[i] & 15);
    }
    return nf_error(ctx, ((seen >> nan_code) & 1u ? NF_NAN : 0) | ((seen >> inf_code) & 1u ? NF_INF : 0));
}

int reserve(sst_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return SST_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes < 256 ? 256 : bytes;
    want += want / 8;  // a little headroom so that slowly growing batches do not reallocate each time
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        cudaGetLastError();
        return fail(ctx, SST_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
    }
    b.cap = want;
    return SST_OK;
}

struct KTimer {  // brackets one kernel family with a pair of pooled events; times are read in flush_timers
    sst_ctx* ctx;
    int slot, idx;
    KTimer(sst_ctx* c, int s) : ctx(c), slot(s), idx(-1) {
        if (ctx->time_kernels && ctx->n_pending < kMaxPending) {
            idx = ctx->n_pending++;
            ctx->pending_slot[idx] = slot;
            cudaEventRecord(ctx->kev[2 * idx], ctx->stream);
        }
    }
    void stop(uint64_t launches = 1) {
        ctx->k_launches[slot] += launches;
        if (idx >= 0) cudaEventRecord(ctx->kev[2 * idx + 1], ctx->stream);
    }
};

inline void trace_mark(sst_ctx* ctx, int k) {
    if (ctx->trace_on) cudaEventRecord(ctx->trace_ev[k], ctx->stream);
}

// call after the stream has been synchronised
void flush_timers(sst_ctx* ctx) {
    for (int i = 0; i < ctx->n_pending; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->kev[2 * i], ctx->kev[2 * i + 1]) == cudaSuccess) ctx->k_ms[ctx->pending_slot[i]] += ms;
    }
    ctx->n_pending = 0;
    cudaGetLastError();
}

constexpr bool kFuseMasksDefault = false;  // (set by measurement: see DESIGN.md, K1)

int launch_build(sst_ctx* ctx, sst_table* t) {
    t->masks_fused = false;
    KTimer kt(ctx, SST_K_BUILD);
    cudaEvent_t e0 = ctx->tev[0], e1 = ctx->tev[1];
    CK(cudaEventRecord(e0, ctx->stream));
    if (t->step_min >= kTileWords) {
        CK(cudaMemsetAsync(t->d_flags, 0, (size_t)t->n_tiles * kBuildMaxWarps * sizeof(int), ctx->stream));
        // rows per warp: the smallest of {1,2,4,8} that covers the rows with at most 16 warps
        int rpw = 1;
        while ((t->R - 1 + rpw - 1) / rpw > kBuildMaxWarps) rpw *= 2;
        int nwarps = (t->R - 1 + rpw - 1) / rpw;
        if (nwarps < 1) nwarps = 1;
        // tiles further apart than this never wait on each other
        int64_t indep = (t->step_min - (kTileWords - 1)) / kTileWords;
        if (indep < 1) indep = 1;
        auto launch = [&](auto kern) -> cudaError_t {
            int occ = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nwarps * 32, 0);
            if (occ < 1) occ = 1;
            int64_t grid = (int64_t)ctx->prop.multiProcessorCount * occ;  // all CTAs co-resident (cooperative launch)
            if (grid > indep) grid = indep;
            if (grid > t->n_tiles) grid = t->n_tiles;
            if (grid < 1) grid = 1;
            void* args[] = {(void*)&t->tbl, (void*)&t->R, (void*)&t->C, (void*)&t->d_step, (void*)&t->d_shift,
                            (void*)&t->last_mask, (void*)&t->n_tiles, (void*)&t->d_flags, (void*)&t->H};
            return cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(nwarps * 32), args, 0, ctx->stream);
        };
        // the build can write the mass-major row masks itself (the tile body holds every row's bit1 word): no second kernel
        // that reads the whole table back.  SST_FUSE_MASKS=0 / 1 overrides the default.
        bool fuse = t->H != nullptr && kFuseMasksDefault;
        if (const char* ev = getenv("SST_FUSE_MASKS")) fuse = t->H != nullptr && atoi(ev) != 0;
        cudaError_t e;
        if (fuse) {
            if (rpw == 1) e = launch(k_build_table<1, 0, true>);
            else if (rpw == 2) e = launch(k_build_table<2, 0, true>);
            else if (rpw == 4) e = launch(k_build_table<4, 0, true>);
            else e = launch(k_build_table<8, 0, true>);
        } else {
            if (rpw == 1) e = launch(k_build_table<1>);
            else if (rpw == 2) e = launch(k_build_table<2>);
            else if (rpw == 4) e = launch(k_build_table<4>);
            else e = launch(k_build_table<8>);
        }
        CK(e);
        t->masks_fused = fuse;
    } else {
        k_build_table_small<<<1, 1024, 0, ctx->stream>>>(t->tbl, t->R, t->C, t->d_step, t->d_shift, t->last_mask);
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(e1, ctx->stream));
    kt.stop(1);
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&t->build_ms, e0, e1));
    flush_timers(ctx);
    return SST_OK;
}

int launch_transpose(sst_ctx* ctx, sst_table* t) {
    // every table path (build, upload, rebuild) comes through here: the last-row summary for the classification
    // probes is derived first and timed with the row masks
    auto summary = [&]() {
        k_last_row_summary<<<(unsigned)((t->C + 255) / 256), 256, 0, ctx->stream>>>(t->tbl + (int64_t)(t->R - 1) * t->C, t->C, t->d_any);
    };
    if (!t->H) {
        summary();
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(ctx->stream));
        return SST_OK;
    }
    KTimer kt(ctx, SST_K_TRANSPOSE);
    cudaEvent_t e0 = ctx->tev[0], e1 = ctx->tev[1];
    CK(cudaEventRecord(e0, ctx->stream));
    summary();
    if (!t->masks_fused)  // (the build has written them already)
        k_transpose_masks<<<(unsigned)t->n_tiles, 256, 0, ctx->stream>>>(t->tbl, t->R, t->C, t->H);
    t->masks_fused = false;
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, ctx->stream));
    kt.stop(2);
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&t->transpose_ms, e0, e1));
    flush_timers(ctx);
    return SST_OK;
}

// multiplier m such that (w * m) >> 20 (32-bit arithmetic, 4096 slots) is distinct for all row weights
uint32_t find_leaf_hash(const int64_t* weights, int R) {
    uint64_t x = 0x9E3779B97F4A7C15ULL;
    for (int trial = 0; trial < 20000; trial++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;  // xorshift64
        const uint32_t mul = (uint32_t)(x >> 16) | 1u;
        bool used[kLeafSlots] = {false};
        bool ok = true;
        for (int r = 1; r < R && ok; r++) {
            const uint32_t slot = ((uint32_t)weights[r] * mul) >> 20;
            ok = !used[slot];
            used[slot] = true;
        }
        if (ok) return mul;
    }
    return 0;
}

int alloc_table(sst_ctx* ctx, sst_table* t, const int64_t* weights, int R, int64_t C, bool with_masks) {
    if (R < 1 || R > kMaxRows) return fail(ctx, SST_ERR_TOO_MANY_ROWS, "table has %d rows; at most %d are supported", R, kMaxRows);
    if (weights[0] != 0) return fail(ctx, SST_ERR_BAD_ARG, "weights[0] must be 0");
    for (int i = 1; i < R; i++)
        if (weights[i] <= weights[i - 1]) return fail(ctx, SST_ERR_BAD_ARG, "weights must be strictly ascending");
    if (C < 1 || C * 32 >= ((int64_t)1 << 31)) return fail(ctx, SST_ERR_BAD_ARG, "table width %lld words out of range", (long long)C);
    if ((int64_t)R * C >= ((int64_t)1 << 31)) return fail(ctx, SST_ERR_NOMEM, "table of %d x %lld words exceeds the 2^31-word limit of the build kernel", R, (long long)C);
    t->R = R;
    t->C = C;
    t->n_tiles = (int)((C + kTileWords - 1) / kTileWords);
    std::vector<int32_t> w(R), st(R), sh(R);
    int64_t step_min = C + 1, w_min = 0;
    for (int i = 0; i < R; i++) {
        w[i] = (int32_t)weights[i];
        t->w_host[i] = weights[i];
        st[i] = (int32_t)(weights[i] / 32);
        sh[i] = (int32_t)(weights[i] % 32);
        if (i >= 1 && st[i] < step_min) step_min = st[i];
        if (i == 1) w_min = weights[i];
    }
    if (R == 1) step_min = C + 1;
    t->step_min = step_min;
    t->w_min = w_min;
    t->leaf_mul = find_leaf_hash(weights, R);
    CK(cudaMalloc(&t->tbl, (size_t)R * (size_t)C * 8));
    CK(cudaMalloc(&t->d_weights, (size_t)kMaxRows * 4));
    CK(cudaMalloc(&t->d_step, (size_t)kMaxRows * 4));
    CK(cudaMalloc(&t->d_shift, (size_t)kMaxRows * 4));
    CK(cudaMalloc(&t->d_flags, (size_t)(t->n_tiles + 1) * kBuildMaxWarps * sizeof(int)));
    CK(cudaMalloc(&t->d_any, (size_t)(C / 32 + 2) * sizeof(uint32_t)));
    std::vector<uint8_t> wb((size_t)(weights[R - 1] >> kWeightBucketShift) + 2);
    for (size_t k = 0, r = 0; k < wb.size(); k++) {
        while (r < (size_t)R && weights[r] < (int64_t)(k << kWeightBucketShift)) r++;
        wb[k] = (uint8_t)r;  // R <= 128
    }
    {   // coarse coin-change count over the row weights: how many compositions have a mass near k * width
        const int64_t limit = C * 32;
        int64_t width = w_min / 8 > 0 ? w_min / 8 : 1;
        if (limit / width > 65536) width = limit / 65536 + 1;
        const int64_t K = limit / width + 2;
        std::vector<double> cnt((size_t)K + 1, 0.0);
        cnt[0] = 1.0;
        for (int r = 1; r < R; r++) {
            int64_t cw = (weights[r] + width / 2) / width;
            if (cw < 1) cw = 1;
            for (int64_t k = cw; k <= K; k++) cnt[(size_t)k] += cnt[(size_t)(k - cw)];
        }
        std::vector<uint32_t> lamq((size_t)K);
        for (int64_t k = 0; k < K; k++) {
            const double v = cnt[(size_t)k] / (double)width * 65536.0 * 4.0;  // x4: real differences sit where compositions cluster
            lamq[(size_t)k] = v < 1073741824.0 ? (uint32_t)v : 1073741824u;
        }
        t->lam_width = (uint32_t)width;
        t->lam_K = (uint32_t)K;
        t->h_lamq = lamq;
        CK(cudaMalloc(&t->d_lamq, (size_t)K * 4));
        CK(cudaMemcpyAsync(t->d_lamq, lamq.data(), (size_t)K * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    CK(cudaMalloc(&t->d_wbucket, wb.size()));
    CK(cudaMemcpyAsync(t->d_wbucket, wb.data(), wb.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(t->d_weights, w.data(), (size_t)R * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(t->d_step, st.data(), (size_t)R * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(t->d_shift, sh.data(), (size_t)R * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));  // the host vectors go out of scope
    if (with_masks) CK(cudaMalloc(&t->H, (size_t)C * 32 * sizeof(uint4)));
    return SST_OK;
}

void free_table(sst_table* t) {
    if (!t) return;
    cudaFree(t->tbl);
    cudaFree(t->H);
    cudaFree(t->d_any);
    cudaFree(t->d_wbucket);
    cudaFree(t->d_lamq);
    cudaFree(t->d_weights);
    cudaFree(t->d_step);
    cudaFree(t->d_shift);
    cudaFree(t->d_flags);
    cudaFree(t->d_cnt2d);
    delete t;
}

TableView view_of(const sst_table* t) {
    TableView tv;
    tv.tbl = t->tbl;
    tv.H = t->H;
    tv.any = t->d_any;
    tv.wbucket = t->d_wbucket;
    tv.weights = t->d_weights;
    tv.R = t->R;
    tv.C = t->C;
    return tv;
}

// the composition-count table of the direct pass: cnt(r, m) for every row and every mass below Mcnt, one launch per
// row (row r needs row r - 1 at the same mass and itself w_r earlier).  Built on the first enumeration that can use it.
int ensure_counts(sst_ctx* ctx, sst_table* t) {
    if (t->cnt_ready) return SST_OK;
    const int64_t M = t->C * 32 < kCountMasses ? t->C * 32 : kCountMasses;
    if (!t->d_cnt2d || t->Mcnt != M) {
        if (t->d_cnt2d) CK(cudaFree(t->d_cnt2d));
        t->d_cnt2d = nullptr;
        cudaError_t e = cudaMalloc(&t->d_cnt2d, (size_t)t->R * (size_t)M * 4);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, SST_ERR_NOMEM, "cudaMalloc of the %d x %lld count table failed: %s", t->R, (long long)M, cudaGetErrorString(e));
        }
        t->Mcnt = M;
    }
    cudaEvent_t e0 = ctx->tev[0], e1 = ctx->tev[1];
    CK(cudaEventRecord(e0, ctx->stream));
    k_count_row0<<<(unsigned)((M + 255) / 256), 256, 0, ctx->stream>>>(t->d_cnt2d, M);
    for (int r = 1; r < t->R; r++) {
        const int64_t w = t->w_host[r];
        const int64_t n = w < M ? w : M;
        k_count_row<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(t->tbl + (int64_t)r * t->C, t->d_cnt2d + (int64_t)(r - 1) * M,
                                                                          t->d_cnt2d + (int64_t)r * M, w, M);
    }
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&t->count_ms, e0, e1));
    ctx->k_launches[SST_K_BUILD] += (uint64_t)t->R;
    t->cnt_ready = true;
    return SST_OK;
}

// mean scheduling cost of the staged batch (block sums copied back; call before the stream is synchronised, read after)
int fetch_costs_enqueue(sst_ctx* ctx, int64_t P) {
    const size_t n = (size_t)((P + kCostBlock - 1) / kCostBlock);
    if (n > ctx->h_blk_cap) {
        if (ctx->h_blk) cudaFreeHost(ctx->h_blk);
        ctx->h_blk = nullptr;
        ctx->h_blk_cap = 0;
        CK(cudaMallocHost(&ctx->h_blk, (n + n / 2 + 64) * 8));
        ctx->h_blk_cap = n + n / 2 + 64;
    }
    if (n) CK(cudaMemcpyAsync(ctx->h_blk, ctx->d_blkcost.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return SST_OK;
}
void fetch_costs_finish(sst_ctx* ctx, int64_t P) {
    const size_t n = (size_t)((P + kCostBlock - 1) / kCostBlock);
    double sum = 0.0;
    for (size_t i = 0; i < n; i++) sum += (double)ctx->h_blk[i];
    ctx->est_cost_per_peak = P > 0 ? sum / (double)P : 0.0;
}

}  // namespace

extern "C" {

int sst_ctx_create(int device, sst_ctx** out) {
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return SST_ERR_NO_DEVICE;
    }
    sst_ctx* ctx = new (std::nothrow) sst_ctx();
    if (!ctx) return SST_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&ctx->prop, device) != cudaSuccess) {
        delete ctx;
        return SST_ERR_NO_DEVICE;
    }
    if (ctx->prop.major != 10) {  // the fatbin only holds sm_100a code
        delete ctx;
        return SST_ERR_NO_DEVICE;
    }
    // Every kernel's per-thread stack is at most ~1.4 KB (the depth-first walk of the enumeration pass); asking for that
    // now makes the driver size its local-memory pool here, once, instead of inside the first launch that needs it.
    cudaDeviceSetLimit(cudaLimitStackSize, 2048);
    cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    cudaEventCreate(&ctx->ev_a);
    cudaEventCreate(&ctx->ev_b);
    cudaEventCreate(&ctx->ev_run);
    for (auto& e : ctx->kev) cudaEventCreate(&e);
    for (auto& e : ctx->tev) cudaEventCreate(&e);
    cudaHostAlloc((void**)&ctx->h_misc, 512, cudaHostAllocDefault);
    cudaHostAlloc((void**)&ctx->h_run, 512, cudaHostAllocMapped);
    if (ctx->h_run) cudaHostGetDevicePointer((void**)&ctx->h_run_dev, ctx->h_run, 0);
    if (cudaMalloc(&ctx->d_bar, 512) == cudaSuccess) cudaMemset(ctx->d_bar, 0, 512);
    if (const char* e = getenv("SST_SPEC_MARGIN_PCT")) {
        const int v = atoi(e);
        if (v >= -100 && v <= 1000) ctx->spec_margin_pct = v;
    }
    if (cudaMalloc(&ctx->d_run, 512) == cudaSuccess) cudaMemset(ctx->d_run, 0, 512);
    *out = ctx;
    return SST_OK;
}

void sst_ctx_destroy(sst_ctx* ctx) {
    if (ctx && ctx->h_blk) cudaFreeHost(ctx->h_blk);
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->d_target, &ctx->d_thr, &ctx->d_maxmods, &ctx->d_mode, &ctx->d_ind, &ctx->d_ismod,
                      &ctx->d_memo_peaks, &ctx->d_status, &ctx->d_cnt, &ctx->d_peakoff, &ctx->d_recs,
                      &ctx->d_blocksums, &ctx->d_memo_keys, &ctx->d_memo_alive, &ctx->d_memo_top,
                      &ctx->d_memo_misc, &ctx->d_flush, &ctx->d_vtarget, &ctx->d_vthr, &ctx->d_vout,
                      &ctx->d_scan, &ctx->d_vmass, &ctx->d_vthrf, &ctx->d_chunk_k, &ctx->d_chunk_r, &ctx->d_tmprecs, &ctx->d_tmppeak, &ctx->d_lvlcnt, &ctx->d_lvlA, &ctx->d_ctalvl, &ctx->d_nodemask, &ctx->d_cobs, &ctx->d_coff, &ctx->d_cout, &ctx->d_bkeys, &ctx->d_btop, &ctx->d_blower, &ctx->d_bupper, &ctx->d_bout, &ctx->d_rootvp, &ctx->d_rootcnt, &ctx->d_tilebase, &ctx->d_ctans, &ctx->d_peakcost, &ctx->d_blkcost, &ctx->d_peakoff32, &ctx->d_bag_m, &ctx->d_bag_meta, &ctx->d_bag_cnt, &ctx->d_bag_off, &ctx->d_bag_path, &ctx->d_roots, &ctx->d_tiles, &ctx->d_block, &ctx->d_lsu, &ctx->d_lobs, &ctx->d_lflags, &ctx->d_lalive, &ctx->d_lidx, &ctx->d_lreach, &ctx->d_lfirst, &ctx->d_lhdr, &ctx->d_lkeys, &ctx->d_llast, &ctx->d_lcall,
                      &ctx->d_item_m[0], &ctx->d_item_m[1], &ctx->d_item_peak[0], &ctx->d_item_peak[1],
                      &ctx->d_item_meta[0], &ctx->d_item_meta[1], &ctx->d_item_all[0], &ctx->d_item_all[1],
                      &ctx->d_item_ind[0], &ctx->d_item_ind[1], &ctx->d_item_path[0], &ctx->d_item_path[1]};
    if (ctx->h_misc) cudaFreeHost(ctx->h_misc);
    if (ctx->h_run) cudaFreeHost(ctx->h_run);
    cudaFree(ctx->d_bar);
    cudaFree(ctx->d_run);
    for (DevBuf* b : bufs) cudaFree(b->p);
    cudaEventDestroy(ctx->ev_a);
    cudaEventDestroy(ctx->ev_b);
    cudaEventDestroy(ctx->ev_run);
    for (auto& e : ctx->kev) cudaEventDestroy(e);
    for (auto& e : ctx->tev) cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->stream2);
    cudaStreamDestroy(ctx->stream2);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sst_last_error(const sst_ctx* ctx) { return ctx ? ctx->err : "no context (no usable sm_100 GPU?)"; }

int sst_device_info(sst_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, uint64_t* free_bytes, uint64_t* total_bytes) {
    CK(cudaSetDevice(ctx->device));
    size_t f = 0, t = 0;
    CK(cudaMemGetInfo(&f, &t));
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return SST_OK;
}

void* sst_host_alloc(sst_ctx* ctx, size_t bytes) {
    void* p = nullptr;
    cudaSetDevice(ctx->device);
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

// page-lock memory the caller owns (a POSIX shared-memory segment: results that land there are gathered without a copy)
int sst_host_register(sst_ctx* ctx, void* p, size_t bytes) {
    CK(cudaSetDevice(ctx->device));
    if (!p || !bytes) return fail(ctx, SST_ERR_BAD_ARG, "nothing to register");
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return SST_OK;
}

int sst_host_unregister(sst_ctx* ctx, void* p) {
    CK(cudaSetDevice(ctx->device));
    if (p) CK(cudaHostUnregister(p));
    return SST_OK;
}

void sst_host_free(sst_ctx* ctx, void* p) {
    (void)ctx;
    if (p) cudaFreeHost(p);
}

int sst_timer_start(sst_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaEventRecord(ctx->ev_a, ctx->stream));
    return SST_OK;
}

int sst_timer_stop(sst_ctx* ctx, float* ms) {
    CK(cudaEventRecord(ctx->ev_b, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev_b));
    CK(cudaEventElapsedTime(ms, ctx->ev_a, ctx->ev_b));
    return SST_OK;
}

int sst_timer_stop_at_run(sst_ctx* ctx, float* ms) {
    CK(cudaEventSynchronize(ctx->ev_run));
    CK(cudaEventElapsedTime(ms, ctx->ev_a, ctx->ev_run));
    return SST_OK;
}

int sst_stats_reset(sst_ctx* ctx) {
    memset(ctx->k_ms, 0, sizeof ctx->k_ms);
    memset(ctx->k_launches, 0, sizeof ctx->k_launches);
    return SST_OK;
}

int sst_kernel_ms(sst_ctx* ctx, float* ms, uint64_t* launches) {
    for (int i = 0; i < SST_K_COUNT_; i++) {
        if (ms) ms[i] = ctx->k_ms[i];
        if (launches) launches[i] = ctx->k_launches[i];
    }
    return SST_OK;
}

int sst_set_item_limit(sst_ctx* ctx, uint64_t limit) {
    ctx->item_limit = limit ? limit : ~0ULL;
    return SST_OK;
}

int sst_flush_l2(sst_ctx* ctx, size_t bytes) {
    CK(cudaSetDevice(ctx->device));
    int rc = reserve(ctx, ctx->d_flush, bytes);
    if (rc) return rc;
    CK(cudaMemsetAsync(ctx->d_flush.p, 0x5a, bytes, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SST_OK;
}

int sst_table_build(sst_ctx* ctx, const int64_t* weights, int R, int64_t max_mass, int compression,
                    uint64_t last_col_mask, int with_masks, sst_table** out) {
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    if (compression != 32)
        return fail(ctx, SST_ERR_COMPRESSION, "compression %d: the device table packs 32 masses per uint64 cell only", compression);
    if (max_mass < 0) return fail(ctx, SST_ERR_BAD_ARG, "max_mass must be >= 0");
    for (int i = 1; i < R; i++)
        if (weights[i] < 32) return fail(ctx, SST_ERR_BAD_ARG, "weight %lld < 32: the in-place word loop of the reference is not a closed form there", (long long)weights[i]);
    const int64_t C = (max_mass + 1 + 31) / 32;
    sst_table* t = new (std::nothrow) sst_table();
    if (!t) return fail(ctx, SST_ERR_NOMEM, "host allocation failed");
    int rc = alloc_table(ctx, t, weights, R, C, with_masks != 0);
    if (rc) {
        free_table(t);
        return rc;
    }
    t->last_mask = last_col_mask;
    t->built_here = true;
    rc = launch_build(ctx, t);
    if (!rc) rc = launch_transpose(ctx, t);
    if (rc) {
        free_table(t);
        return rc;
    }
    *out = t;
    return SST_OK;
}

int sst_table_upload(sst_ctx* ctx, const uint64_t* host_table, const int64_t* weights, int R, int64_t C, sst_table** out) {
    *out = nullptr;
    CK(cudaSetDevice(ctx->device));
    sst_table* t = new (std::nothrow) sst_table();
    if (!t) return fail(ctx, SST_ERR_NOMEM, "host allocation failed");
    int rc = alloc_table(ctx, t, weights, R, C, true);
    if (rc) {
        free_table(t);
        return rc;
    }
    cudaError_t e = cudaMemcpyAsync(t->tbl, host_table, (size_t)R * (size_t)C * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) {
        free_table(t);
        return fail(ctx, SST_ERR_CUDA, "table upload: %s", cudaGetErrorString(e));
    }
    rc = launch_transpose(ctx, t);
    if (rc) {
        free_table(t);
        return rc;
    }
    *out = t;
    return SST_OK;
}

int sst_table_rebuild(sst_ctx* ctx, sst_table* t) {
    CK(cudaSetDevice(ctx->device));
    if (!t->built_here) return fail(ctx, SST_ERR_STATE, "an uploaded table cannot be rebuilt");
    t->cnt_ready = false;  // the count table follows the table bits: rebuilt on the next enumeration that uses it
    int rc = launch_build(ctx, t);
    if (!rc) rc = launch_transpose(ctx, t);
    return rc;
}

int sst_table_info(const sst_table* t, int* R, int64_t* C, float* build_ms, float* transpose_ms) {
    if (R) *R = t->R;
    if (C) *C = t->C;
    if (build_ms) *build_ms = t->build_ms;
    if (transpose_ms) *transpose_ms = t->transpose_ms;
    return SST_OK;
}

int sst_table_download(sst_ctx* ctx, const sst_table* t, uint64_t* host_out) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host_out, t->tbl, (size_t)t->R * (size_t)t->C * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SST_OK;
}

int sst_table_download_masks(sst_ctx* ctx, const sst_table* t, int64_t first_mass, int64_t n, uint32_t* host_out) {
    CK(cudaSetDevice(ctx->device));
    if (!t->H) return fail(ctx, SST_ERR_STATE, "table was built without row masks");
    if (first_mass < 0 || n < 0 || first_mass + n > t->C * 32) return fail(ctx, SST_ERR_BAD_ARG, "mask range out of table");
    CK(cudaMemcpyAsync(host_out, t->H + first_mass, (size_t)n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SST_OK;
}

void sst_table_destroy(sst_ctx* ctx, sst_table* t) {
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream2);
        cudaStreamSynchronize(ctx->stream);
    }
    free_table(t);
}

int sst_valid_stage(sst_ctx* ctx, const int64_t* target, const int64_t* thr, int64_t P) {
    CK(cudaSetDevice(ctx->device));
    if (P < 0) return fail(ctx, SST_ERR_BAD_ARG, "negative probe count");
    int rc;
    if ((rc = reserve(ctx, ctx->d_vtarget, (size_t)(P ? P : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vthr, (size_t)(P ? P : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vout, (size_t)(P ? P : 1)))) return rc;
    if (P) {
        CK(cudaMemcpyAsync(ctx->d_vtarget.p, target, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_vthr.p, thr, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->VP = P;
    ctx->valid_f64 = false;
    return SST_OK;
}

int sst_valid_stage_f64(sst_ctx* ctx, const double* mass, const double* thr, int64_t P, double precision, double tolerance) {
    CK(cudaSetDevice(ctx->device));
    if (P < 0) return fail(ctx, SST_ERR_BAD_ARG, "negative probe count");
    int rc;
    if ((rc = reserve(ctx, ctx->d_vmass, (size_t)(P ? P : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vthrf, (size_t)(P ? P : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vout, (size_t)(P ? P : 1)))) return rc;
    if (P) {
        CK(cudaMemcpyAsync(ctx->d_vmass.p, mass, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (thr) CK(cudaMemcpyAsync(ctx->d_vthrf.p, thr, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        else CK(cudaMemsetAsync(ctx->d_vthrf.p, 0xFF, (size_t)P * 8, ctx->stream));  // all-ones = NaN = "relative"
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->VP = P;
    ctx->valid_f64 = true;
    ctx->v_precision = precision;
    ctx->v_tolerance = tolerance;
    return SST_OK;
}

int sst_valid_run(sst_ctx* ctx, const sst_table* t) {
    CK(cudaSetDevice(ctx->device));
    const int64_t P = ctx->VP;
    if (P) {
        KTimer kt(ctx, SST_K_IS_VALID);
        if (ctx->valid_f64)
            k_is_valid_f64<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(view_of(t), (const double*)ctx->d_vmass.p, (const double*)ctx->d_vthrf.p,
                                                                                 ctx->v_precision, ctx->v_tolerance, P, (uint8_t*)ctx->d_vout.p);
        else
            k_is_valid<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(view_of(t), (const int64_t*)ctx->d_vtarget.p,
                                                                             (const int64_t*)ctx->d_vthr.p, P, (uint8_t*)ctx->d_vout.p);
        kt.stop(1);
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(ctx->stream));
    flush_timers(ctx);
    return SST_OK;
}

int sst_valid_fetch(sst_ctx* ctx, uint8_t* out) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->VP) CK(cudaMemcpyAsync(out, ctx->d_vout.p, (size_t)ctx->VP, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->valid_f64) return nf_error_codes(ctx, out, ctx->VP, VALID_CODE_NAN, VALID_CODE_INF, false);
    return SST_OK;
}

// the three calls above in one, with ONE synchronisation: what a reference-shaped is_valid_mass (a batch of one) costs is
// the waiting, not the kernel
int sst_is_valid_f64(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, int64_t P, double precision, double tolerance,
                     uint8_t* out) {
    CK(cudaSetDevice(ctx->device));
    if (P < 0) return fail(ctx, SST_ERR_BAD_ARG, "negative probe count");
    if (!P) return SST_OK;
    int rc;
    if ((rc = reserve(ctx, ctx->d_vmass, (size_t)P * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vthrf, (size_t)P * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vout, (size_t)P))) return rc;
    CK(cudaMemcpyAsync(ctx->d_vmass.p, mass, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (thr) CK(cudaMemcpyAsync(ctx->d_vthrf.p, thr, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
    {
        KTimer kt(ctx, SST_K_IS_VALID);
        k_is_valid_f64<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(view_of(t), (const double*)ctx->d_vmass.p,
                                                                             thr ? (const double*)ctx->d_vthrf.p : nullptr, precision, tolerance, P,
                                                                             (uint8_t*)ctx->d_vout.p);
        kt.stop(1);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->d_vout.p, (size_t)P, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    flush_timers(ctx);
    ctx->VP = 0;  // (nothing stays staged)
    return nf_error_codes(ctx, out, P, VALID_CODE_NAN, VALID_CODE_INF, false);
}

int sst_is_valid(sst_ctx* ctx, const sst_table* t, const int64_t* target, const int64_t* thr, int64_t P, uint8_t* out) {
    int rc = sst_valid_stage(ctx, target, thr, P);
    if (!rc) rc = sst_valid_run(ctx, t);
    if (!rc) rc = sst_valid_fetch(ctx, out);
    return rc;
}

int sst_classify_stage(sst_ctx* ctx, const double* observed, int64_t F, const double* offsets, int B) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream2));  // an asynchronous classification may still own the buffers
    if (F < 0 || B < 0 || B > 65535) return fail(ctx, SST_ERR_BAD_ARG, "fragment / breakage count out of range");
    int rc;
    if ((rc = reserve(ctx, ctx->d_cobs, (size_t)(F ? F : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_coff, (size_t)(B ? B : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_cout, (size_t)(F * B ? F * B : 1)))) return rc;
    if (F) CK(cudaMemcpyAsync(ctx->d_cobs.p, observed, (size_t)F * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (B) CK(cudaMemcpyAsync(ctx->d_coff.p, offsets, (size_t)B * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->CF = F;
    ctx->CB = B;
    return SST_OK;
}

int sst_classify_launch(sst_ctx* ctx, const sst_table* t, double precision, double tolerance) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream2));
    if (ctx->CF && ctx->CB) {
        KTimer kt(ctx, SST_K_CLASSIFY);
        k_classify<<<dim3((unsigned)((ctx->CF + 255) / 256), (unsigned)((ctx->CB + kClassifyPerThread - 1) / kClassifyPerThread)), 256, 0, ctx->stream>>>(view_of(t), (const double*)ctx->d_cobs.p, ctx->CF,
                                                                              (const double*)ctx->d_coff.p, ctx->CB, precision, tolerance,
                                                                              (uint8_t*)ctx->d_cout.p, 0);
        kt.stop(1);
        CK(cudaGetLastError());
    }
    return SST_OK;
}

int sst_classify_run(sst_ctx* ctx, const sst_table* t, double precision, double tolerance) {
    int rc = sst_classify_launch(ctx, t, precision, tolerance);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    flush_timers(ctx);
    return SST_OK;
}

// the whole classification on the context's SIDE stream, without waiting: stage (pinned `observed` / `offsets`
// recommended), kernel, copy of the flags into `out` (pinned).  sst_classify_wait completes it.  An enumeration pass
// issued in between runs concurrently on the main stream.
static int classify_async(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B,
                          double precision, double tolerance, uint8_t* out, int pack4);

int sst_classify_async(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B,
                       double precision, double tolerance, uint8_t* out) {
    return classify_async(ctx, t, observed, F, offsets, B, precision, tolerance, out, 0);
}

int sst_classify_async_packed(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B,
                              double precision, double tolerance, uint8_t* out) {
    return classify_async(ctx, t, observed, F, offsets, B, precision, tolerance, out, 1);
}

static int classify_async(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B,
                          double precision, double tolerance, uint8_t* out, int pack4) {
    hp_begin();
    CK(cudaSetDevice(ctx->device));
    if (F < 0 || B < 0 || B > 65535) return fail(ctx, SST_ERR_BAD_ARG, "fragment / breakage count out of range");
    CK(cudaStreamSynchronize(ctx->stream2));  // an earlier asynchronous classification still owns the buffers
    hp_mark(16);
    int rc;
    if ((rc = reserve(ctx, ctx->d_cobs, (size_t)(F ? F : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_coff, (size_t)(B ? B : 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_cout, (size_t)(F * B ? F * B : 1)))) return rc;
    ctx->CF = F;
    ctx->CB = B;
    if (!F || !B) return SST_OK;
    CK(cudaMemcpyAsync(ctx->d_cobs.p, observed, (size_t)F * 8, cudaMemcpyHostToDevice, ctx->stream2));
    CK(cudaMemcpyAsync(ctx->d_coff.p, offsets, (size_t)B * 8, cudaMemcpyHostToDevice, ctx->stream2));
    hp_mark(18);
    k_classify<<<dim3((unsigned)((F + 255) / 256), (unsigned)((B + kClassifyPerThread - 1) / kClassifyPerThread)), 256, 0, ctx->stream2>>>(view_of(t), (const double*)ctx->d_cobs.p, F,
                                                                                       (const double*)ctx->d_coff.p, B, precision, tolerance,
                                                                                       (uint8_t*)ctx->d_cout.p, pack4);
    CK(cudaGetLastError());
    hp_mark(19);
    ctx->k_launches[SST_K_CLASSIFY] += 1;
    CK(cudaMemcpyAsync(out, ctx->d_cout.p, pack4 ? (size_t)(((F + 1) & ~1LL) / 2) * B : (size_t)F * B, cudaMemcpyDeviceToHost, ctx->stream2));
    hp_mark(20);
    ctx->c_async_out = out;
    ctx->c_async_pack4 = pack4 != 0;
    return SST_OK;
}

int sst_classify_wait(sst_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream2));
    if (ctx->c_async_out && ctx->CF && ctx->CB) {  // a non-finite observed mass marks every breakage row: the first one tells
        const uint8_t* row0 = ctx->c_async_out;
        ctx->c_async_out = nullptr;
        return nf_error_codes(ctx, row0, ctx->c_async_pack4 ? (ctx->CF + 1) / 2 : ctx->CF, CLASS_CODE_NAN, CLASS_CODE_INF, ctx->c_async_pack4);
    }
    ctx->c_async_out = nullptr;
    return SST_OK;
}

int sst_classify_fetch(sst_ctx* ctx, uint8_t* out) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->CF && ctx->CB) CK(cudaMemcpyAsync(out, ctx->d_cout.p, (size_t)ctx->CF * ctx->CB, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->CF && ctx->CB) return nf_error_codes(ctx, out, ctx->CF, CLASS_CODE_NAN, CLASS_CODE_INF, false);
    return SST_OK;
}

int sst_classify(sst_ctx* ctx, const sst_table* t, const double* observed, int64_t F, const double* offsets, int B, double precision,
                 double tolerance, uint8_t* out) {
    int rc = sst_classify_stage(ctx, observed, F, offsets, B);
    if (!rc) rc = sst_classify_run(ctx, t, precision, tolerance);
    if (!rc) rc = sst_classify_fetch(ctx, out);
    return rc;
}

int sst_length_bounds(sst_ctx* ctx, const sst_table* t, int64_t target, int64_t thr, int32_t max_mods, int32_t max_len,
                      const int32_t* ind, const uint8_t* is_mod, uint64_t memo_capacity, int64_t* lower, int64_t* upper) {
    CK(cudaSetDevice(ctx->device));
    if (!t->H) return fail(ctx, SST_ERR_STATE, "table was built without row masks");
    if (max_len < 0 || max_len > 120) return fail(ctx, SST_ERR_TOO_DEEP, "max_len %d outside [0, 120]", (int)max_len);
    if (t->w_min > 0 && (target + thr) / t->w_min > kMaxDepth)
        return fail(ctx, SST_ERR_TOO_DEEP, "a composition may need %lld nucleotides (limit %d)", (long long)((target + thr) / t->w_min), kMaxDepth);
    uint64_t cap = memo_capacity ? memo_capacity : ((uint64_t)1 << 16);
    uint64_t pow2 = 1024;
    while (pow2 < cap) pow2 <<= 1;
    if (pow2 > ((uint64_t)1 << 26)) return fail(ctx, SST_ERR_NOMEM, "memo capacity %llu too large", (unsigned long long)cap);
    int rc;
    if ((rc = reserve(ctx, ctx->d_bkeys, pow2 * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_btop, pow2))) return rc;
    if ((rc = reserve(ctx, ctx->d_blower, pow2 * kMaxRows))) return rc;
    if ((rc = reserve(ctx, ctx->d_bupper, pow2 * kMaxRows))) return rc;
    if ((rc = reserve(ctx, ctx->d_bout, 64))) return rc;
    if ((rc = reserve(ctx, ctx->d_ind, (size_t)kMaxRows * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_ismod, (size_t)kMaxRows))) return rc;
    CK(cudaMemsetAsync(ctx->d_bkeys.p, 0, pow2 * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_btop.p, 0, pow2, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_bout.p, 0, 64, ctx->stream));
    ctx->up_R = -1;  // (d_ind / d_ismod no longer hold what stage_f64_enqueue remembers)
    CK(cudaMemcpyAsync(ctx->d_ind.p, ind, (size_t)t->R * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ismod.p, is_mod, (size_t)t->R, cudaMemcpyHostToDevice, ctx->stream));
    ctx->have_result = false;  // d_ind / d_ismod of a staged enumeration batch are gone: it has to be staged again
    ctx->R_staged = -1;
    BoundMap mp{(uint32_t*)ctx->d_bkeys.p, (uint8_t*)ctx->d_btop.p, (int8_t*)ctx->d_blower.p, (int8_t*)ctx->d_bupper.p,
                (uint32_t)(pow2 - 1), (int*)((char*)ctx->d_bout.p + 32)};
    {
        KTimer kt(ctx, SST_K_LENGTH_BOUND);
        k_length_bounds<<<1, 32, 0, ctx->stream>>>(view_of(t), RowMeta{(const int32_t*)ctx->d_ind.p, (const uint8_t*)ctx->d_ismod.p}, target, thr,
                                                   max_mods, max_len, mp, (int64_t*)ctx->d_bout.p);
        kt.stop(1);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(ctx->h_misc, ctx->d_bout.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    flush_timers(ctx);
    const int64_t* h = (const int64_t*)ctx->h_misc;
    if (ctx->h_misc[8]) return fail(ctx, SST_ERR_MEMO_FULL, "memo of the length-bound walk is too small (%llu slots)", (unsigned long long)pow2);
    if (h[2]) return fail(ctx, SST_ERR_OUT_OF_TABLE, "a value of the mass window is not in the DP table");
    if (lower) *lower = h[0];
    if (upper) *upper = h[1];
    return SST_OK;
}

int sst_explain_stage(sst_ctx* ctx, const sst_table* t, const int64_t* target, const int64_t* thr, const int32_t* max_mods,
                      const uint8_t* mode, int64_t P, const int32_t* ind, const uint8_t* is_mod) {
    CK(cudaSetDevice(ctx->device));
    ctx->have_result = false;
    if (P < 0) return fail(ctx, SST_ERR_BAD_ARG, "negative peak count");
    if (!t->H) return fail(ctx, SST_ERR_STATE, "table was built without row masks");
    int rc;
    const size_t p8 = (size_t)(P ? P : 1) * 8;
    if ((rc = reserve(ctx, ctx->d_target, p8))) return rc;
    if ((rc = reserve(ctx, ctx->d_thr, p8))) return rc;
    if ((rc = reserve(ctx, ctx->d_maxmods, p8 / 2))) return rc;
    if ((rc = reserve(ctx, ctx->d_mode, p8 / 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_ind, (size_t)kMaxRows * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_ismod, (size_t)kMaxRows))) return rc;
    std::vector<uint32_t> memo_peaks;
    int64_t window_total = 0, max_hi = 0;
    bool has_exact = false;
    for (int64_t p = 0; p < P; p++) {
        if (mode[p] == SST_MODE_MEMO) memo_peaks.push_back((uint32_t)p);
        else if (mode[p] == SST_MODE_EXACT) has_exact = true;
        else if (mode[p] != SST_MODE_FREE) return fail(ctx, SST_ERR_BAD_ARG, "peak %lld: unknown mode %d", (long long)p, (int)mode[p]);
        if (thr[p] >= 0) {
            int64_t hi = target[p] + thr[p], lo = target[p] - thr[p];
            if (hi > max_hi) max_hi = hi;
            int64_t a = lo < 1 ? 1 : lo, b = hi < t->C * 32 - 1 ? hi : t->C * 32 - 1;
            if (b >= a) window_total += b - a + 1;
        }
    }
    if (P) {
        CK(cudaMemcpyAsync(ctx->d_target.p, target, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_thr.p, thr, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_maxmods.p, max_mods, (size_t)P * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_mode.p, mode, (size_t)P, cudaMemcpyHostToDevice, ctx->stream));
    }
    ctx->up_R = -1;  // (d_ind / d_ismod no longer hold what stage_f64_enqueue remembers)
    CK(cudaMemcpyAsync(ctx->d_ind.p, ind, (size_t)t->R * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_ismod.p, is_mod, (size_t)t->R, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_memo = (int)memo_peaks.size();
    if (ctx->n_memo) {
        if ((rc = reserve(ctx, ctx->d_memo_peaks, memo_peaks.size() * 4))) return rc;
        CK(cudaMemcpyAsync(ctx->d_memo_peaks.p, memo_peaks.data(), memo_peaks.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = reserve(ctx, ctx->d_peakcost, (size_t)(P + 1) * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_blkcost, (size_t)((P + kCostBlock - 1) / kCostBlock + 1) * 8))) return rc;
    if (P) {
        k_peak_costs<<<(unsigned)((P + kCostBlock - 1) / kCostBlock), kCostBlock, 0, ctx->stream>>>(
            (const int64_t*)ctx->d_target.p, (const int64_t*)ctx->d_thr.p, P, t->C * 32, CostModel{t->d_lamq, t->lam_width, t->lam_K},
            (uint32_t*)ctx->d_peakcost.p, (unsigned long long*)ctx->d_blkcost.p);
        CK(cudaGetLastError());
    }
    if ((rc = fetch_costs_enqueue(ctx, P))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    fetch_costs_finish(ctx, P);
    ctx->P = P;
    ctx->R_staged = t->R;
    ctx->window_total = window_total;
    ctx->has_exact = has_exact;
    ctx->max_hi = max_hi;
    {
        const int64_t cap = t->C * 32 - 1;
        ctx->deepest = t->w_min > 0 ? (max_hi < cap ? max_hi : cap) / t->w_min : 0;
    }
    return SST_OK;
}

}  // extern "C"

namespace {
// max_mods == nullptr: every peak has the budget `uniform_mods` (the array is filled on the device)
// everything of the staging that can be queued without waiting: copies in, k_stage_f64, the summary on its way to h_misc
int stage_f64_enqueue(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, const int32_t* max_mods, int32_t uniform_mods,
                      int64_t P, const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo,
                      unsigned long long* summary_at = nullptr, bool inputs_on_device = false) {
    const bool summary_later = summary_at != nullptr;  // (part of a result block that is copied out as a whole)
    CK(cudaSetDevice(ctx->device));
    ctx->have_result = false;
    if (P < 0) return fail(ctx, SST_ERR_BAD_ARG, "negative peak count");
    if (!t->H) return fail(ctx, SST_ERR_STATE, "table was built without row masks");
    // Integerisation (the float operations of mass_explanation.py:107-114), the choice of the budget mode and
    // the batch summary all run on the device (k_stage_f64): the host only derives the two mode thresholds
    // from the per-row budgets.  FREE when no composition inside the window can exhaust a budget.
    int64_t w_min_mod = 0, hi_limit = INT64_MAX;
    for (int r = 1; r < t->R; r++)
        if (is_mod[r]) {
            const int64_t w = t->w_host[r];
            if (!w_min_mod || w < w_min_mod) w_min_mod = w;
            const int64_t lim = ((int64_t)ind[r] + 1) * w;  // ind[r] >= hi / w  <=>  hi < (ind[r]+1) * w
            if (lim < hi_limit) hi_limit = lim;
        }
    int rc;
    const size_t p8 = (size_t)(P ? P : 1) * 8;
    if ((rc = reserve(ctx, ctx->d_target, p8))) return rc;
    if ((rc = reserve(ctx, ctx->d_thr, p8))) return rc;
    if ((rc = reserve(ctx, ctx->d_maxmods, p8 / 2))) return rc;
    if ((rc = reserve(ctx, ctx->d_mode, p8 / 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_ind, (size_t)kMaxRows * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_ismod, (size_t)kMaxRows))) return rc;
    if ((rc = reserve(ctx, ctx->d_vmass, p8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vthrf, p8))) return rc;
    if ((rc = reserve(ctx, ctx->d_memo_peaks, p8 / 2))) return rc;
    if ((rc = reserve(ctx, ctx->d_scan, 512))) return rc;
    if ((rc = reserve(ctx, ctx->d_peakcost, (size_t)(P + 1) * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_blkcost, (size_t)((P + kCostBlock - 1) / kCostBlock + 1) * 8))) return rc;
    hp_mark(2);
    trace_mark(ctx, 0);
    unsigned long long* const summary = summary_at ? summary_at : (unsigned long long*)ctx->d_scan.p;
    CK(cudaMemsetAsync(summary, 0, 64, ctx->stream));
    hp_mark(3);
    if (P && !inputs_on_device) {  // (the ladder generator writes masses and thresholds into d_vmass / d_vthrf itself)
        CK(cudaMemcpyAsync(ctx->d_vmass.p, mass, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (thr) CK(cudaMemcpyAsync(ctx->d_vthrf.p, thr, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (P && max_mods) CK(cudaMemcpyAsync(ctx->d_maxmods.p, max_mods, (size_t)P * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->up_R != t->R || memcmp(ctx->up_ind, ind, (size_t)t->R * 4) || memcmp(ctx->up_ismod, is_mod, (size_t)t->R)) {
        memcpy(ctx->up_ind, ind, (size_t)t->R * 4);  // (the copies below read the context's arrays: the caller's may go away)
        memcpy(ctx->up_ismod, is_mod, (size_t)t->R);
        ctx->up_R = t->R;
        CK(cudaMemcpyAsync(ctx->d_ind.p, ctx->up_ind, (size_t)t->R * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_ismod.p, ctx->up_ismod, (size_t)t->R, cudaMemcpyHostToDevice, ctx->stream));
    }
    hp_mark(4);
    trace_mark(ctx, 1);
    if (P) {
        k_stage_f64<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(
            (const double*)ctx->d_vmass.p, thr ? (const double*)ctx->d_vthrf.p : nullptr, (int32_t*)ctx->d_maxmods.p, uniform_mods,
            max_mods ? 0 : 1, P, precision,
            tolerance, w_min_mod, hi_limit, with_memo ? SST_MODE_MEMO : SST_MODE_EXACT, t->C * 32, (int64_t*)ctx->d_target.p,
            (int64_t*)ctx->d_thr.p, (uint8_t*)ctx->d_mode.p, (uint32_t*)ctx->d_memo_peaks.p, summary,
            CostModel{t->d_lamq, t->lam_width, t->lam_K}, (uint32_t*)ctx->d_peakcost.p, (unsigned long long*)ctx->d_blkcost.p);
        CK(cudaGetLastError());
    }
    hp_mark(5);
    // (a pipelined submission gets the summary with the result block: a copy between staging and pass would make the pass
    //  wait for the copy engine, which is still busy with the previous batch's records)
    if (!summary_later) CK(cudaMemcpyAsync(ctx->h_misc, ctx->d_scan.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    hp_mark(6);
    trace_mark(ctx, 2);
    ctx->P = P;
    ctx->R_staged = t->R;
    return SST_OK;
}

int stage_f64(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, const int32_t* max_mods, int32_t uniform_mods,
              int64_t P, const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo,
              bool inputs_on_device = false) {
    int rc = stage_f64_enqueue(ctx, t, mass, thr, max_mods, uniform_mods, P, ind, is_mod, precision, tolerance, with_memo, nullptr,
                               inputs_on_device);
    if (rc) return rc;
    if ((rc = fetch_costs_enqueue(ctx, P))) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    fetch_costs_finish(ctx, P);
    const unsigned long long* h = (const unsigned long long*)ctx->h_misc;
    if ((rc = nf_error(ctx, h[4]))) return rc;
    ctx->window_total = (int64_t)h[0];
    ctx->max_hi = (int64_t)h[1];
    ctx->n_memo = (int)h[2];
    ctx->has_exact = h[3] != 0;
    {
        const int64_t cap = t->C * 32 - 1;
        ctx->deepest = t->w_min > 0 ? (ctx->max_hi < cap ? ctx->max_hi : cap) / t->w_min : 0;
    }
    return SST_OK;
}
}  // namespace

extern "C" {

int sst_explain_stage_f64(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, const int32_t* max_mods,
                          int64_t P, const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo) {
    if (!max_mods && P) return fail(ctx, SST_ERR_BAD_ARG, "max_mods is null");
    return stage_f64(ctx, t, mass, thr, max_mods, 0, P, ind, is_mod, precision, tolerance, with_memo);
}

int sst_explain_stage_f64_uniform(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, int32_t max_mods,
                                  int64_t P, const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo) {
    return stage_f64(ctx, t, mass, thr, nullptr, max_mods, P, ind, is_mod, precision, tolerance, with_memo);
}

}  // extern "C"

namespace {

// MEMO mode: the first-visit replay (k_memo_phase_a) fills the hash map the enumeration reads its edges from.  Launched
// once per run, before the pass; memo_check looks at its counters after the stream has been synchronised.
int memo_launch(sst_ctx* ctx, const sst_table* t, uint64_t memo_capacity, MemoMap& mp) {
    const int64_t P = ctx->P;
    uint64_t mcap = memo_capacity ? memo_capacity : ((uint64_t)1 << 20);
    uint64_t pow2 = 1024;
    while (pow2 < mcap) pow2 <<= 1;
    if (pow2 > ((uint64_t)1 << 31)) return fail(ctx, SST_ERR_NOMEM, "memo capacity %llu too large", (unsigned long long)mcap);
    int rc;
    if ((rc = reserve(ctx, ctx->d_memo_keys, pow2 * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_memo_alive, pow2 * 16))) return rc;
    if ((rc = reserve(ctx, ctx->d_memo_top, pow2 * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_memo_misc, 64))) return rc;
    CK(cudaMemsetAsync(ctx->d_memo_keys.p, 0, pow2 * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_memo_alive.p, 0, pow2 * 16, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_memo_top.p, 0, pow2 * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_memo_misc.p, 0, 64, ctx->stream));
    mp.keys = (unsigned long long*)ctx->d_memo_keys.p;
    mp.alive = (uint4*)ctx->d_memo_alive.p;
    mp.top = (uint32_t*)ctx->d_memo_top.p;
    mp.cap_mask = (uint32_t)(pow2 - 1);
    mp.fill = (unsigned int*)ctx->d_memo_misc.p;
    mp.overflow = (int*)ctx->d_memo_misc.p + 1;
    KTimer kt(ctx, SST_K_PHASE_A);
    // the replay stack lives in shared memory: frames for the deepest composition, as many threads per CTA as fit
    const int64_t frames = ctx->deepest + 3;
    const int threads = replay_threads(frames, kReplaySmemLimit);
    if (!threads) return fail(ctx, SST_ERR_TOO_DEEP, "a composition may need %lld nucleotides: the replay stack does not fit in shared memory", (long long)ctx->deepest);
    if (!ctx->replay_attr_set) {
        CK(cudaFuncSetAttribute((const void*)k_memo_phase_a, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kReplaySmemLimit));
        ctx->replay_attr_set = true;
    }
    k_memo_phase_a<<<(unsigned)((ctx->n_memo + threads - 1) / threads), threads, (size_t)frames * threads * kReplayFrameBytes, ctx->stream>>>(
        view_of(t), RowMeta{(const int32_t*)ctx->d_ind.p, (const uint8_t*)ctx->d_ismod.p},
        PeakBatch{(const int64_t*)ctx->d_target.p, (const int64_t*)ctx->d_thr.p, (const int32_t*)ctx->d_maxmods.p,
                  (const uint8_t*)ctx->d_mode.p, P},
        (const uint32_t*)ctx->d_memo_peaks.p, ctx->n_memo, mp, (int)frames);
    kt.stop(1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_misc + 100, ctx->d_memo_misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    return SST_OK;
}

// too small = an insertion failed, or the load factor passed 3/4 (the replay still finished, but probing a map that
// full is slow: the caller grows it)
int memo_check(sst_ctx* ctx, const MemoMap& mp) {
    if (ctx->h_misc[101] || (unsigned)ctx->h_misc[100] > (mp.cap_mask >> 1) + (mp.cap_mask >> 2))
        return fail(ctx, SST_ERR_MEMO_FULL, "first-visit map is too small (%d slots used)", ctx->h_misc[100]);
    return SST_OK;
}

int grow_records(sst_ctx* ctx, unsigned long long comps, int rec_width) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const unsigned long long need = comps * (unsigned long long)rec_width;
    if (need > (unsigned long long)free_b + ctx->d_recs.cap)
        return fail(ctx, SST_ERR_NOMEM, "%llu compositions x %d bytes do not fit in device memory (%zu bytes free)", comps, rec_width, free_b);
    return reserve(ctx, ctx->d_recs, (size_t)need + (need >> 2) + 8);
}

enum { PASS_DONE = 0, PASS_FALLBACK = -1 };
constexpr double kHeavyCostPerPeak = 4000.0;  // mean peak_cost() above which the automatic choice is the level-synchronous pass

// Can the direct pass (sst_direct.cuh) take the staged batch?  Every peak FREE, compositions of at most kDirDepth
// nucleotides in 8- or 16-byte records, every window inside the count table.
bool direct_eligible(const sst_ctx* ctx, const sst_table* t, int rec_width) {
    return t->H && !ctx->n_memo && !ctx->has_exact && ctx->deepest <= kDirDepth && rec_width <= 16 &&
           (t->C * 32 <= kCountMasses || ctx->max_hi < kCountMasses);
}

// Direct pass: counts from the table, two grid barriers, output-balanced fill.  One cooperative launch.
int direct_enqueue(sst_ctx* ctx, sst_table* t, int rec_width) {
    const int64_t P = ctx->P;
    const int nw = rec_width / 8;
    int rc;
    if ((rc = ensure_counts(ctx, t))) return rc;
    auto kern = nw == 1 ? k_explain_direct<1> : k_explain_direct<2>;
    int& grid_max = ctx->dir_grid_max[nw == 1 ? 0 : 1];
    if (!grid_max) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kDirThreads, 0));
        if (occ < 1) return fail(ctx, SST_ERR_CUDA, "k_explain_direct does not fit on an SM");
        grid_max = occ * ctx->prop.multiProcessorCount;
        if (grid_max > 512) grid_max = 512;
    }
    // a CTA per 128 peaks, at most one co-resident wave: a single-peak call is one CTA and never waits at a barrier
    int64_t want = (P + 127) / 128;
    if (want < 1) want = 1;
    const unsigned grid = (unsigned)(want < grid_max ? want : grid_max);
    if ((rc = reserve(ctx, ctx->d_status, (size_t)(P ? P : 1)))) return rc;
    if ((rc = reserve(ctx, ctx->d_peakoff, (size_t)(P + 2) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_peakoff32, (size_t)(P + 2) * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_blocksums, (size_t)3 * (grid_max > ctx->pass_grid_cap ? grid_max : ctx->pass_grid_cap) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_cnt, (size_t)(P + 1) * 4))) return rc;
    if (!ctx->d_recs.p && (rc = reserve(ctx, ctx->d_recs, (size_t)64 << 20))) return rc;
    if (!ctx->h_run_dev || !ctx->d_bar) return fail(ctx, SST_ERR_NOMEM, "run summary buffers are missing");
    const size_t bag = (size_t)grid_max * kBagItems;
    if ((rc = reserve(ctx, ctx->d_bag_m, bag * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_bag_meta, bag * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_bag_cnt, bag * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_bag_off, bag * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_bag_path, bag * 16))) return rc;
    // the root pool keeps its size between runs and grows when a run reports that it was too small
    // peaks per tile: every CTA of the grid gets a tile even when the batch is small (a multiple of 32, at most a CTA's threads)
    int64_t tile = ((P + grid - 1) / grid + 31) / 32 * 32;
    if (tile < 32) tile = 32;
    if (tile > kDirThreads) tile = kDirThreads;
    const int64_t n_tiles = (P + tile - 1) / tile;
    if ((rc = reserve(ctx, ctx->d_tiles, (size_t)(n_tiles + 2) * 20))) return rc;
    if (ctx->root_capacity < (uint64_t)(8 * P + 65536) + (uint64_t)grid_max * kRootGranule) ctx->root_capacity = (uint64_t)(8 * P + 65536) + (uint64_t)grid_max * kRootGranule;
    if ((rc = reserve(ctx, ctx->d_roots, (size_t)ctx->root_capacity * 12))) return rc;
    DirArgs a{};
    a.tile_size = (int)tile;
    a.root_m = (uint32_t*)ctx->d_roots.p;
    a.root_pre = a.root_m + ctx->root_capacity;
    a.root_n = a.root_pre + ctx->root_capacity;
    a.root_cap = ctx->root_capacity;
    a.tile_start = (unsigned long long*)ctx->d_tiles.p;
    a.tile_base = a.tile_start + (n_tiles + 1);
    a.tile_nroots = (uint32_t*)(a.tile_base + (n_tiles + 1));
    a.tv = view_of(t);
    a.cv = CountView{t->d_cnt2d, t->Mcnt};
    a.pk = PeakBatch{(const int64_t*)ctx->d_target.p, (const int64_t*)ctx->d_thr.p, (const int32_t*)ctx->d_maxmods.p,
                     (const uint8_t*)ctx->d_mode.p, P};
    a.status = (uint8_t*)ctx->d_status.p;
    a.rel = (uint32_t*)ctx->d_cnt.p;
    a.recs = (unsigned long long*)ctx->d_recs.p;
    a.rec_capacity = (unsigned long long)(ctx->d_recs.cap / rec_width);
    a.peak_off = (unsigned long long*)ctx->d_peakoff.p;
    a.peak_off32 = (uint32_t*)ctx->d_peakoff32.p;
    a.cta_tot = (unsigned long long*)ctx->d_blocksums.p;
    a.bag_m = (uint32_t*)ctx->d_bag_m.p;
    a.bag_meta = (uint32_t*)ctx->d_bag_meta.p;
    a.bag_cnt = (uint32_t*)ctx->d_bag_cnt.p;
    a.bag_off = (unsigned long long*)ctx->d_bag_off.p;
    a.bag_path = (unsigned long long*)ctx->d_bag_path.p;
    a.sync = ctx->d_bar + 64 * (ctx->run_parity & 1);
    a.sync_next = ctx->d_bar + 64 * ((ctx->run_parity + 1) & 1);
    a.host_out = ctx->h_run_dev;
    a.leaf = LeafHash{t->leaf_mul};
    a.cta_ns = nullptr;
    if (ctx->want_cta_ns) {
        if ((rc = reserve(ctx, ctx->d_ctans, (size_t)grid * 64))) return rc;
        CK(cudaMemsetAsync(ctx->d_ctans.p, 0, (size_t)grid * 64, ctx->stream));
        a.cta_ns = (unsigned long long*)ctx->d_ctans.p;
        ctx->cta_ns_grid = (int)grid;
    }
    {
        KTimer kt(ctx, SST_K_EXPLAIN_PASS);
        void* args[] = {(void*)&a};
        CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kDirThreads), args, 0, ctx->stream));
        ctx->run_parity++;  // only a launch that really started clears the other set
        kt.stop(1);
    }
    CK(cudaEventRecord(ctx->ev_run, ctx->stream));  // the device is done here; what follows is the host waking up
    return SST_OK;
}

// Depth-first pass (sst_enum.cuh): one cooperative launch, one grid barrier.  Returns PASS_FALLBACK when a root's
// subtree is too large for one thread (the level-synchronous pass balances such batches across the machine).
struct OutBlock {  // where a pipelined submission wants status, 32-bit offsets, records and the run summary
    uint8_t* base;
    BlockLayout at;
    uint64_t rec_bytes;
    uint64_t capN;  // > 0: split records — uint32 lo[capN] at at.recs, then hp byte planes of capN each
    int hp;
};
int dfs_enqueue(sst_ctx* ctx, const sst_table* t, int rec_width, const MemoMap& mp, const SpecGuard* guard = nullptr,
                const OutBlock* ob = nullptr) {
    const int64_t P = ctx->P;
    const int nw = rec_width / 8;
    int rc;
    const bool wide = P <= kDfsWideMaxPeaks;  // one 1024-thread CTA per SM for batches in the latency regime
    const int threads = wide ? kDfsThreadsWide : kDfsThreads;
    auto kern = wide ? (nw == 1 ? (ctx->has_exact ? k_explain_dfs<1, true, kDfsThreadsWide> : k_explain_dfs<1, false, kDfsThreadsWide>)
                                : (ctx->has_exact ? k_explain_dfs<2, true, kDfsThreadsWide> : k_explain_dfs<2, false, kDfsThreadsWide>))
                     : (nw == 1 ? (ctx->has_exact ? k_explain_dfs<1, true, kDfsThreads> : k_explain_dfs<1, false, kDfsThreads>)
                                : (ctx->has_exact ? k_explain_dfs<2, true, kDfsThreads> : k_explain_dfs<2, false, kDfsThreads>));
    int& grid_max = ctx->dfs_grid_max[(wide ? 4 : 0) + (nw == 1 ? 0 : 2) + (ctx->has_exact ? 1 : 0)];
    if (!grid_max) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0));
        if (occ < 1) return fail(ctx, SST_ERR_CUDA, "k_explain_dfs does not fit on an SM");
        grid_max = occ * ctx->prop.multiProcessorCount;
    }
    // a CTA per 128 peaks, at most one co-resident wave: a single-peak call is one CTA and never waits at the barrier
    int64_t want = (P + 127) / 128;
    if (want < 1) want = 1;
    const unsigned grid = (unsigned)(want < grid_max ? want : grid_max);
    if ((rc = reserve(ctx, ctx->d_status, (size_t)(P ? P : 1)))) return rc;
    if ((rc = reserve(ctx, ctx->d_peakoff, (size_t)(P + 2) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_blocksums, (size_t)3 * (grid_max > ctx->pass_grid_cap ? grid_max : ctx->pass_grid_cap) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_cnt, (size_t)(P + 1) * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_tilebase, (size_t)grid * 24 + (size_t)(P / 32 + grid + 8) * 4))) return rc;
    if (!ctx->d_recs.p && (rc = reserve(ctx, ctx->d_recs, (size_t)64 << 20))) return rc;
    if (!ctx->h_run_dev || !ctx->d_bar || !ctx->d_run) return fail(ctx, SST_ERR_NOMEM, "run summary buffers are missing");
    // the item pool holds every tile's window roots and the lists of its split rounds; it keeps its size between runs
    // and grows when a run reports that it was too small
    if (ctx->pool_capacity < (uint64_t)(32 * P + (1 << 20))) ctx->pool_capacity = ((uint64_t)(32 * P + (1 << 20)) + 63) & ~63ULL;
    if ((rc = reserve(ctx, ctx->d_peakoff32, (size_t)(P + 2) * 4))) return rc;
    {
        const size_t cap = (size_t)ctx->pool_capacity;
        if ((rc = reserve(ctx, ctx->d_item_m[0], cap * 4))) return rc;
        if ((rc = reserve(ctx, ctx->d_item_peak[0], cap * 4))) return rc;
        if ((rc = reserve(ctx, ctx->d_item_meta[0], cap * 4))) return rc;
        if ((rc = reserve(ctx, ctx->d_item_path[0], cap * 8 * (size_t)nw))) return rc;
        if ((rc = reserve(ctx, ctx->d_rootcnt, cap * 4))) return rc;
        if ((rc = reserve(ctx, ctx->d_rootvp, cap * 4))) return rc;
        if ((rc = reserve(ctx, ctx->d_nodemask, cap * 16))) return rc;
        if ((rc = reserve(ctx, ctx->d_chunk_k, (cap / 32 + 64) * 4))) return rc;
        if (ctx->has_exact) {
            if ((rc = reserve(ctx, ctx->d_item_all[0], cap * 4))) return rc;
            if ((rc = reserve(ctx, ctx->d_item_ind[0], cap * 4))) return rc;
        }
        DfsArgs a{};
        a.tv = view_of(t);
        a.meta = RowMeta{(const int32_t*)ctx->d_ind.p, (const uint8_t*)ctx->d_ismod.p};
        a.pk = PeakBatch{(const int64_t*)ctx->d_target.p, (const int64_t*)ctx->d_thr.p, (const int32_t*)ctx->d_maxmods.p,
                         (const uint8_t*)ctx->d_mode.p, P};
        a.mp = mp;
        a.status = (uint8_t*)ctx->d_status.p;
        a.peak_first = (uint32_t*)ctx->d_cnt.p;
        a.pool = ItemPool{(uint32_t*)ctx->d_item_m[0].p, (uint32_t*)ctx->d_item_peak[0].p, (uint32_t*)ctx->d_item_meta[0].p,
                          (unsigned long long*)ctx->d_item_path[0].p, (int32_t*)ctx->d_item_all[0].p, (int32_t*)ctx->d_item_ind[0].p,
                          (uint4*)ctx->d_nodemask.p, (uint32_t*)ctx->d_rootcnt.p, (uint32_t*)ctx->d_rootvp.p, (uint32_t*)ctx->d_chunk_k.p,
                          (unsigned long long)cap};
        a.peak_cost = (const uint32_t*)ctx->d_peakcost.p;
        a.blk_cost = (const unsigned long long*)ctx->d_blkcost.p;
        a.cta_info = (unsigned long long*)ctx->d_tilebase.p;
        a.peak_cpre = (uint32_t*)((char*)ctx->d_tilebase.p + (size_t)grid * 24);
        a.recs = (unsigned long long*)ctx->d_recs.p;
        a.rec_capacity = (unsigned long long)(ctx->d_recs.cap / rec_width);
        a.peak_off = (unsigned long long*)ctx->d_peakoff.p;
        a.peak_off32 = (uint32_t*)ctx->d_peakoff32.p;
        a.cta_tot = (unsigned long long*)ctx->d_blocksums.p;
        a.sync = ctx->d_bar + 64 * (ctx->run_parity & 1);
        a.sync_next = ctx->d_bar + 64 * ((ctx->run_parity + 1) & 1);
        a.host_out = ctx->h_run_dev;
        a.leaf = LeafHash{t->leaf_mul};
        a.cta_ns = nullptr;
        if (ctx->want_cta_ns) {
            if ((rc = reserve(ctx, ctx->d_ctans, (size_t)grid * 64))) return rc;
            a.cta_ns = (unsigned long long*)ctx->d_ctans.p;
            ctx->cta_ns_grid = (int)grid;
        }
        a.guard = guard ? *guard : SpecGuard{};
        if (ob) {  // everything the host wants back lies in one block
            a.status = ob->base + ob->at.status;
            a.peak_off32 = (uint32_t*)(ob->base + ob->at.off32);
            a.recs = (unsigned long long*)(ob->base + ob->at.recs);
            a.rec_capacity = (unsigned long long)(ob->rec_bytes / rec_width);
            if (ob->capN) {
                a.rec_lo = (uint32_t*)(ob->base + ob->at.recs);
                a.rec_hi = ob->base + ob->at.recs + 4 * ob->capN;
                a.rec_hi_planes = ob->hp;
                a.rec_capacity = ob->capN;
            }
            a.host_out = (unsigned long long*)ob->base;
        }
        hp_mark(7);
        {
            KTimer kt(ctx, SST_K_EXPLAIN_PASS);
            hp_mark(8);
            void* args[] = {(void*)&a};
            CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(threads), args, 0, ctx->stream));
            hp_mark(9);
            ctx->run_parity++;  // only a launch that really started clears the other set
            kt.stop(1);
        }
        if (guard) trace_mark(ctx, 3);
        CK(cudaEventRecord(ctx->ev_run, ctx->stream));  // the device is done here; what follows is the host waking up
        hp_mark(10);
    }
    return SST_OK;
}

enum { DFS_OK = 0, DFS_RETRY = -2 };
// after the stream has been synchronised: what the pass left in the run summary
int dfs_evaluate(sst_ctx* ctx, int rec_width, const MemoMap& mp, bool memo_fresh, int attempt, unsigned long long* roots, unsigned long long* comps) {
    const int64_t P = ctx->P;
    const int nw = rec_width / 8;
    int rc;
    flush_timers(ctx);
    const unsigned long long* h_tot = ctx->h_run;
    const int* h_flags = reinterpret_cast<const int*>(ctx->h_run + 40);
    *roots = h_tot[0];
    *comps = h_tot[2];
    ctx->levels = 1;
    for (int i = 0; i < 32; i++) ctx->phase_ns[i] = h_tot[8 + i];
    if (memo_fresh && attempt == 0 && (rc = memo_check(ctx, mp))) return rc;
    if (h_flags[3]) return PASS_FALLBACK;
    if (h_flags[2]) {  // the item pool was too small: grow and run the pass again
        // a pool far beyond the batch size means combinatorial blow-up: that is the level-synchronous pass's job
        // (it spreads single huge subtrees over the machine and enforces sst_set_item_limit)
        unsigned long long most = 256ULL * (unsigned long long)(P > 0 ? P : 1);
        if (most < (32ULL << 20)) most = 32ULL << 20;
        if (ctx->item_limit < most) most = ctx->item_limit;
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const unsigned long long per_item = 37 + 8ULL * nw + (ctx->has_exact ? 8 : 0);
        if (attempt >= 10 || ctx->pool_capacity * 2 > most || ctx->pool_capacity * per_item > (unsigned long long)free_b) return PASS_FALLBACK;
        ctx->pool_capacity *= 2;
        return DFS_RETRY;
    }
    if (h_flags[1]) {  // records did not fit: grow and run the pass again
        if (attempt >= 3) return fail(ctx, SST_ERR_CUDA, "record buffer kept overflowing (%llu compositions)", *comps);
        if ((rc = grow_records(ctx, *comps, rec_width))) return rc;
        return DFS_RETRY;
    }
    return DFS_OK;
}

// Depth-first pass (sst_enum.cuh): one cooperative launch, one grid barrier.  Returns PASS_FALLBACK when the batch
// is not for it (a subtree too large for one thread, combinatorial blow-up of the item pool).
int run_dfs_pass(sst_ctx* ctx, const sst_table* t, int rec_width, const MemoMap& mp, bool memo_fresh, unsigned long long* roots,
                 unsigned long long* comps) {
    for (int attempt = 0;; attempt++) {
        int rc = dfs_enqueue(ctx, t, rec_width, mp);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        rc = dfs_evaluate(ctx, rec_width, mp, memo_fresh, attempt, roots, comps);
        if (rc == DFS_RETRY) continue;
        if (rc) return rc;
        break;
    }
    ctx->last_pass = 2;
    return PASS_DONE;
}

// Level-synchronous pass (sst_explain.cuh, k_explain_pass): one cooperative launch, two grid barriers per level.
// after the stream has been synchronised: what the direct pass left in the run summary
int direct_evaluate(sst_ctx* ctx, int rec_width, int attempt, unsigned long long* roots, unsigned long long* comps) {
    int rc;
    flush_timers(ctx);
    const unsigned long long* h_tot = ctx->h_run;
    const int* h_flags = reinterpret_cast<const int*>(ctx->h_run + 40);
    *roots = h_tot[0];
    *comps = h_tot[2];
    ctx->levels = 1;
    for (int i = 0; i < 32; i++) ctx->phase_ns[i] = h_tot[8 + i];
    if (h_flags[2]) {  // the root pool was too small: grow to what the run asked for and run again
        if (attempt >= 3) return PASS_FALLBACK;
        const unsigned long long want = h_tot[1] + h_tot[1] / 8 + 4096;
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        if (want * 12ULL > (unsigned long long)free_b + ctx->d_roots.cap) return PASS_FALLBACK;
        ctx->root_capacity = want;
        return DFS_RETRY;
    }
    if (h_flags[3]) return PASS_FALLBACK;
    if (h_flags[1]) {  // records did not fit: grow and run the pass again
        if (attempt >= 3) return fail(ctx, SST_ERR_CUDA, "record buffer kept overflowing (%llu compositions)", *comps);
        if ((rc = grow_records(ctx, *comps, rec_width))) return rc;
        return DFS_RETRY;
    }
    return DFS_OK;
}

// Returns PASS_FALLBACK when the batch is not for the direct pass after all (a saturated count, a table whose bits are
// not a consistent knapsack table, a bag overflow): the item pass walks such batches.
int run_direct_pass(sst_ctx* ctx, sst_table* t, int rec_width, unsigned long long* roots, unsigned long long* comps) {
    for (int attempt = 0;; attempt++) {
        int rc = direct_enqueue(ctx, t, rec_width);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        rc = direct_evaluate(ctx, rec_width, attempt, roots, comps);
        if (rc == DFS_RETRY) continue;
        if (rc) return rc;
        break;
    }
    ctx->last_pass = 3;
    return PASS_DONE;
}

int run_level_pass(sst_ctx* ctx, const sst_table* t, int rec_width, const MemoMap& mp, bool memo_fresh, unsigned long long* roots_out,
                   unsigned long long* comps_out) {
    const int64_t P = ctx->P;
    // Level-0 items are bounded by the summed window sizes (known to the host); the two item buffers and the
    // record buffer keep their capacity from earlier runs — a level that would overflow sets a flag and the pass
    // is repeated with larger buffers.
    const int64_t root_bound = ctx->window_total;
    const int nw = rec_width / 8;
    int rc;
    auto kern = nw == 1 ? k_explain_pass<1> : nw == 2 ? k_explain_pass<2> : k_explain_pass<0>;
    int& grid_max = ctx->pass_grid_max[nw == 1 ? 0 : nw == 2 ? 1 : 2];
    if (!grid_max) {
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kPassThreads, 0));
        if (occ < 1) return fail(ctx, SST_ERR_CUDA, "k_explain_pass does not fit on an SM");
        grid_max = occ * ctx->prop.multiProcessorCount;
        if (grid_max > kPassThreads) grid_max = kPassThreads;  // the slice totals are scanned by one CTA (balanced_range)
        if (grid_max > ctx->pass_grid_cap) ctx->pass_grid_cap = grid_max;
    }
    if ((rc = reserve(ctx, ctx->d_status, (size_t)(P ? P : 1)))) return rc;
    if ((rc = reserve(ctx, ctx->d_peakoff, (size_t)(P + 2) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_blocksums, (size_t)3 * ctx->pass_grid_cap * 8))) return rc;
    int lvl_cap = (int)ctx->deepest + 3;
    if (lvl_cap > kMaxLevels) lvl_cap = kMaxLevels;
    if ((rc = reserve(ctx, ctx->d_lvlcnt, (size_t)lvl_cap * (size_t)(P + 1) * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_lvlA, (size_t)lvl_cap * (size_t)(P + 1) * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_ctalvl, (size_t)lvl_cap * (size_t)grid_max * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_scan, 512))) return rc;
    if (!ctx->d_recs.p && (rc = reserve(ctx, ctx->d_recs, (size_t)64 << 20))) return rc;
    if (!ctx->item_capacity) ctx->item_capacity = (uint64_t)1 << 20;
    if ((int64_t)ctx->item_capacity < root_bound + 1) ctx->item_capacity = (uint64_t)root_bound + 1;

    unsigned long long roots = 0, items = 0, comps = 0;
    for (int attempt = 0;; attempt++) {
        const size_t cap = (size_t)ctx->item_capacity;
        for (int k = 0; k < 2; k++) {
            if ((rc = reserve(ctx, ctx->d_item_m[k], cap * 4))) return rc;
            if ((rc = reserve(ctx, ctx->d_item_peak[k], cap * 4))) return rc;
            if ((rc = reserve(ctx, ctx->d_item_meta[k], cap * 4))) return rc;
            if ((rc = reserve(ctx, ctx->d_item_path[k], cap * 8 * (size_t)nw))) return rc;
            if (ctx->has_exact) {
                if ((rc = reserve(ctx, ctx->d_item_all[k], cap * 4))) return rc;
                if ((rc = reserve(ctx, ctx->d_item_ind[k], cap * 4))) return rc;
            }
        }
        const size_t most = (size_t)P > cap ? (size_t)P : cap;
        if ((rc = reserve(ctx, ctx->d_cnt, (most + 1) * 4))) return rc;
        const size_t n_chunks = most / 32 + (size_t)grid_max + 64;
        if ((rc = reserve(ctx, ctx->d_nodemask, (cap + 1) * 16))) return rc;
        if ((rc = reserve(ctx, ctx->d_chunk_k, n_chunks * 4))) return rc;
        if ((rc = reserve(ctx, ctx->d_chunk_r, n_chunks * 4))) return rc;
        // the level-ordered record buffer mirrors the result buffer
        if ((rc = reserve(ctx, ctx->d_tmprecs, ctx->d_recs.cap))) return rc;
        if ((rc = reserve(ctx, ctx->d_tmppeak, (ctx->d_recs.cap / rec_width + 1) * 4))) return rc;
        // totals, timestamps and flags live in shared memory of CTA 0 during the pass, which stores them to pinned host
        // memory on its way out (h_run: [0,40) totals and timestamps, then the flags)
        if (!ctx->h_run_dev || !ctx->d_bar) return fail(ctx, SST_ERR_NOMEM, "run summary buffers are missing");

        PassArgs a{};
        a.tv = view_of(t);
        a.meta = RowMeta{(const int32_t*)ctx->d_ind.p, (const uint8_t*)ctx->d_ismod.p};
        a.pk = PeakBatch{(const int64_t*)ctx->d_target.p, (const int64_t*)ctx->d_thr.p, (const int32_t*)ctx->d_maxmods.p,
                         (const uint8_t*)ctx->d_mode.p, P};
        a.mp = mp;
        a.status = (uint8_t*)ctx->d_status.p;
        for (int k = 0; k < 2; k++)
            a.buf[k] = ItemBuf{(uint32_t*)ctx->d_item_m[k].p, (uint32_t*)ctx->d_item_peak[k].p, (uint32_t*)ctx->d_item_meta[k].p,
                               (int32_t*)ctx->d_item_all[k].p, (int32_t*)ctx->d_item_ind[k].p, (unsigned long long*)ctx->d_item_path[k].p};
        a.cap = (unsigned long long)cap;
        a.item_limit = ctx->item_limit;
        a.cnt = (uint32_t*)ctx->d_cnt.p;
        a.node_mask = (uint4*)ctx->d_nodemask.p;
        a.chunk_k = (uint32_t*)ctx->d_chunk_k.p;
        a.chunk_r = (uint32_t*)ctx->d_chunk_r.p;
        a.tmp_recs = (unsigned long long*)ctx->d_tmprecs.p;
        a.tmp_peak = (uint32_t*)ctx->d_tmppeak.p;
        a.lvl_cnt = (uint32_t*)ctx->d_lvlcnt.p;
        a.lvl_A = (unsigned long long*)ctx->d_lvlA.p;
        a.lvl_cap = lvl_cap;
        a.cta_lvl = (unsigned long long*)ctx->d_ctalvl.p;
        a.nw = nw;
        a.has_budget = ctx->has_exact ? 1 : 0;
        a.recs = (uint8_t*)ctx->d_recs.p;
        a.rec_capacity = (unsigned long long)(ctx->d_recs.cap / rec_width);
        a.peak_off = (unsigned long long*)ctx->d_peakoff.p;
        a.cta_tot = (unsigned long long*)ctx->d_blocksums.p;
        a.barrier = ctx->d_bar + 64 * (ctx->run_parity & 1);
        a.barrier_next = ctx->d_bar + 64 * ((ctx->run_parity + 1) & 1);
        a.host_out = ctx->h_run_dev;
        a.leaf = LeafHash{t->leaf_mul};
        // enough CTAs that every thread gets about one entity of the longest list we can foresee (peaks, window
        // values, the widest level of the previous run), at most one co-resident wave: a single-peak call must not
        // pay twelve grid barriers of a full-machine grid
        int64_t foresee = P > root_bound ? P : root_bound;
        if ((int64_t)ctx->widest_level > foresee) foresee = (int64_t)ctx->widest_level;
        int64_t want = (foresee + kPassThreads - 1) / kPassThreads;
        if (want < 1) want = 1;
        const unsigned grid = (unsigned)(want < grid_max ? want : grid_max);
        {
            KTimer kt(ctx, SST_K_EXPLAIN_PASS);
            void* args[] = {(void*)&a};
            CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kPassThreads), args, 0, ctx->stream));
            ctx->run_parity++;  // only a launch that really started clears the other counter (a failed one would leave it stale)
            kt.stop(1);
        }
        // totals + timestamps + flags arrive in h_run by the kernel's own stores
        CK(cudaEventRecord(ctx->ev_run, ctx->stream));  // the device is done here; what follows is the host waking up
        CK(cudaStreamSynchronize(ctx->stream));
        flush_timers(ctx);
        const unsigned long long* h_tot = ctx->h_run;
        const int* h_flags = reinterpret_cast<const int*>(ctx->h_run + 40);
        roots = h_tot[0];
        items = h_tot[1];
        comps = h_tot[2];
        ctx->levels = (int)h_tot[3];
        if (!h_flags[2]) ctx->widest_level = items;
        for (int i = 0; i < 32; i++) ctx->phase_ns[i] = h_tot[8 + i];
        if (memo_fresh && attempt == 0 && (rc = memo_check(ctx, mp))) return rc;
        if (h_flags[0])
            return fail(ctx, SST_ERR_NOMEM, "more than %llu partial compositions in one level (%llu): combinatorial blow-up (raise the limit with sst_set_item_limit)",
                        (unsigned long long)ctx->item_limit, items);
        if (h_flags[2]) {  // a level did not fit: grow (deeper levels are larger still) and run the pass again
            if (attempt >= 12) return fail(ctx, SST_ERR_CUDA, "item buffers kept overflowing (%llu items)", items);
            size_t free_b = 0, total_b = 0;
            CK(cudaMemGetInfo(&free_b, &total_b));
            const unsigned long long want_items = items * 2 + 1024;
            const unsigned long long per_item = 2ULL * (12 + 8ULL * nw + (ctx->has_exact ? 8 : 0)) + 4;
            if ((want_items - cap) * per_item > (unsigned long long)free_b)
                return fail(ctx, SST_ERR_NOMEM, "%llu partial compositions do not fit in device memory (%zu bytes free)", items, free_b);
            ctx->item_capacity = want_items;
            continue;
        }
        if (h_flags[1]) {  // records did not fit: grow and run the pass again
            if (attempt >= 14) return fail(ctx, SST_ERR_CUDA, "record buffer kept overflowing (%llu compositions)", comps);
            if ((rc = grow_records(ctx, comps, rec_width))) return rc;
            continue;
        }
        break;
    }
    ctx->n_items = items;
    ctx->last_pass = 1;
    *roots_out = roots;
    *comps_out = comps;
    return PASS_DONE;
}

}  // namespace

extern "C" {

int sst_set_pass(sst_ctx* ctx, int which) {
    if ((which < 0 || which > 3) && which != -3)
        return fail(ctx, SST_ERR_BAD_ARG, "pass %d: 0 = automatic, 1 = level-synchronous, 2 = depth-first items, 3 = direct (count table), -3 = automatic without the direct pass", which);
    ctx->pass_choice = which;
    return SST_OK;
}

int sst_last_pass(const sst_ctx* ctx) { return ctx->last_pass; }

int sst_explain_cta_ns(sst_ctx* ctx, int enable, uint64_t* out, int cap_ctas, int* n_ctas) {
    CK(cudaSetDevice(ctx->device));
    ctx->want_cta_ns = enable != 0;
    const int n = ctx->cta_ns_grid < cap_ctas ? ctx->cta_ns_grid : cap_ctas;
    if (out && n > 0 && ctx->d_ctans.p) {
        CK(cudaMemcpyAsync(out, ctx->d_ctans.p, (size_t)n * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (n_ctas) *n_ctas = out ? n : 0;
    return SST_OK;
}

int sst_explain_run(sst_ctx* ctx, const sst_table* t, int rec_width, uint64_t memo_capacity, uint64_t* n_roots,
                    uint64_t* n_comps) {
    CK(cudaSetDevice(ctx->device));
    ctx->have_result = false;
    if (ctx->R_staged != t->R) return fail(ctx, SST_ERR_STATE, "staged batch belongs to a table with %d rows", ctx->R_staged);
    if (ctx->deepest > kMaxDepth - 3) return fail(ctx, SST_ERR_TOO_DEEP, "a composition may need %lld nucleotides (limit %d)", (long long)ctx->deepest, kMaxDepth - 3);
    if (rec_width == 0) rec_width = (int)(8 * (ctx->deepest > 8 ? (ctx->deepest + 7) / 8 : 1));
    if (rec_width < 8 || rec_width % 8 || rec_width > kMaxDepth) return fail(ctx, SST_ERR_BAD_ARG, "rec_width %d must be a multiple of 8 in [8, %d]", rec_width, kMaxDepth);
    if (ctx->deepest > rec_width) return fail(ctx, SST_ERR_BAD_ARG, "a composition may need %lld nucleotides but rec_width is %d", (long long)ctx->deepest, rec_width);
    // Depth-first pass for everything up to kDfsDepth nucleotides (ladder differences, singletons), level-synchronous
    // pass for deeper batches and for those the depth-first pass gives back (a subtree too large for one thread).
    const bool dfs_ok = ctx->deepest <= kDfsDepth && rec_width <= 16;
    if (ctx->pass_choice == 2 && !dfs_ok)
        return fail(ctx, SST_ERR_TOO_DEEP, "the depth-first pass holds at most %d nucleotides per composition (batch: %lld)", kDfsDepth, (long long)ctx->deepest);
    const bool direct_ok = direct_eligible(ctx, t, rec_width);
    if (ctx->pass_choice == 3 && !direct_ok)
        return fail(ctx, SST_ERR_STATE, "the direct pass needs a batch without binding budgets, compositions of at most %d nucleotides and windows below %lld", kDirDepth, (long long)kCountMasses);
    // Automatic choice: a batch of ladder differences (a handful of compositions per peak) takes the depth-first item
    // pass; a batch whose peaks have thousands of compositions each (sparse ladders, wide windows: peak_cost() of the
    // staged batch) takes the level-synchronous pass, which spreads single huge subtrees over the whole machine —
    // measured on the C5 workload: 0.50 ms against 9.6 ms (item pass) and 14 ms (direct pass).  The direct pass is
    // there for the asking (sst_set_pass(ctx, 3)): its count phase is twice as fast as the item pass's, its fill slower.
    const bool heavy = ctx->est_cost_per_peak > kHeavyCostPerPeak;
    bool use_direct = ctx->pass_choice == 3;
    bool use_dfs = ctx->pass_choice == 2 || ((ctx->pass_choice == 0 || ctx->pass_choice == -3) && dfs_ok && !heavy);
    MemoMap mp{};
    int rc;
    if (ctx->n_memo && (rc = memo_launch(ctx, t, memo_capacity, mp))) return rc;
    unsigned long long roots = 0, comps = 0;
    bool memo_fresh = ctx->n_memo != 0;
    if (use_direct) {
        rc = run_direct_pass(ctx, const_cast<sst_table*>(t), rec_width, &roots, &comps);
        if (rc == PASS_FALLBACK && ctx->pass_choice == 3)
            return fail(ctx, SST_ERR_STATE, "the direct pass gave the batch back (a saturated count, inconsistent table bits or a full bag)");
        if (rc == PASS_FALLBACK) use_direct = false;
        else if (rc) return rc;
    }
    if (use_direct) use_dfs = false;
    if (use_dfs) {
        rc = run_dfs_pass(ctx, t, rec_width, mp, memo_fresh, &roots, &comps);
        if (rc == PASS_FALLBACK && ctx->pass_choice == 2)
            return fail(ctx, SST_ERR_NOMEM, "a window value has more than %u partial compositions: too large for the depth-first pass", kDfsNodeCap);
        if (rc == PASS_FALLBACK) {
            use_dfs = false;
            memo_fresh = false;  // already checked
        } else if (rc) {
            return rc;
        }
    }
    if (!use_dfs && !use_direct && (rc = run_level_pass(ctx, t, rec_width, mp, memo_fresh, &roots, &comps))) return rc;
    ctx->n_roots = roots;
    ctx->n_comps = comps;
    ctx->rec_width = rec_width;
    ctx->have_result = true;
    if (n_roots) *n_roots = roots;
    if (n_comps) *n_comps = comps;
    return SST_OK;
}

// ---- the whole call without waiting: inputs in, staging, pass, results out are queued on the context's stream ----
int sst_explain_submit_f64(sst_ctx* ctx, const sst_table* t, const double* mass, const double* thr, int32_t max_mods, int64_t P,
                           const int32_t* ind, const uint8_t* is_mod, double precision, double tolerance, int with_memo, uint8_t* out_block,
                           uint64_t block_bytes) {
    hp_begin();
    CK(cudaSetDevice(ctx->device));
    if (ctx->pend.active) return fail(ctx, SST_ERR_STATE, "a submitted batch has not been collected yet");
    if (P < 0) return fail(ctx, SST_ERR_BAD_ARG, "negative peak count");
    if (!t->H) return fail(ctx, SST_ERR_STATE, "table was built without row masks");
    sst_ctx::Pending& pd = ctx->pend;
    pd = sst_ctx::Pending{};
    pd.mass = mass; pd.thr = thr; pd.max_mods = max_mods; pd.P = P; pd.ind = ind; pd.is_mod = is_mod;
    pd.precision = precision; pd.tolerance = tolerance; pd.with_memo = with_memo;
    const BlockLayout lay(P);
    if (!out_block || block_bytes < lay.recs) return fail(ctx, SST_ERR_BAD_ARG, "result block of %llu bytes: %llu are needed before the first record", (unsigned long long)block_bytes, (unsigned long long)lay.recs);
    pd.block = out_block;
    pd.status = out_block + lay.status; pd.off32 = (uint32_t*)(out_block + lay.off32); pd.recs = out_block + lay.recs; pd.recs_bytes = block_bytes - lay.recs;
    pd.active = true;
    ctx->have_result = false;
    // Everything is queued without a look at the batch, on the assumption that it is like the previous ones: no budget
    // binds (every peak FREE), compositions fit the record width of last time, the depth-first pass is the right one.
    // The staging kernel's summary (largest window end, MEMO / EXACT peaks, non-finite inputs, summed cost estimate)
    // stays on the device, where the pass checks it before it does anything; a batch that breaks an assumption comes
    // back untouched and is carried out synchronously in sst_explain_collect, which also decides whether the next
    // submission speculates again.  No pass over the host arrays.
    const bool fast = ctx->pass_choice != 1 && ctx->pass_choice != 3 && P > 0 && ctx->spec_ok;
    hp_mark(0);
    if (fast) {
        ctx->deepest = ctx->spec_rec_width;
        pd.rec_width = ctx->spec_rec_width;
    }
    hp_mark(1);
    if (!fast) return SST_OK;  // sst_explain_collect does the work
    // the device block: room for twice the records that are copied back blindly (a larger batch gets the rest in collect)
    uint64_t guess = (ctx->last_comps + ctx->last_comps * (uint64_t)ctx->spec_margin_pct / 100) * (uint64_t)pd.rec_width + 4096;
    if (guess > pd.recs_bytes) guess = pd.recs_bytes;
    // split records: 4 + hp bytes per composition in planes laid out for capN compositions, the same in both blocks
    pd.split = ctx->split_records && pd.rec_width == 8;
    uint64_t guessN = 0;
    if (pd.split) {
        pd.hp = ctx->spec_hi_planes < 0 ? 0 : (ctx->spec_hi_planes > 4 ? 4 : ctx->spec_hi_planes);
        pd.capN = (pd.recs_bytes / (uint64_t)(4 + pd.hp)) & ~15ULL;
        guessN = ctx->last_comps + ctx->last_comps * (uint64_t)ctx->spec_margin_pct / 100 + 512;
        if (guessN > pd.capN) guessN = pd.capN;
        if (!pd.capN) pd.split = false;
    }
    int rc;
    {
        uint64_t dev_recs = 2 * guess > ((uint64_t)16 << 20) ? 2 * guess : ((uint64_t)16 << 20);
        if (dev_recs < pd.recs_bytes && pd.recs_bytes <= ((uint64_t)1 << 30)) dev_recs = pd.recs_bytes;
        if (pd.split) dev_recs = pd.recs_bytes;  // (plane offsets are the caller's)
        if ((rc = reserve(ctx, ctx->d_block, (size_t)(lay.recs + dev_recs)))) {
            pd.active = false;
            return rc;
        }
    }
    const OutBlock ob{(uint8_t*)ctx->d_block.p, lay, (uint64_t)ctx->d_block.cap - lay.recs, pd.split ? pd.capN : 0, pd.hp};
    rc = stage_f64_enqueue(ctx, t, mass, thr, nullptr, max_mods, P, ind, is_mod, precision, tolerance, with_memo,
                           (unsigned long long*)(ob.base + kBlockStageSummary));
    if (rc) {
        pd.active = false;
        return rc;
    }
    ctx->n_memo = 0;
    ctx->has_exact = false;
    ctx->window_total = 0;  // not known without the summary; only the level-synchronous pass sizes its buffers from it
    MemoMap mp{};
    pd.direct = false;
    SpecGuard g{};
    g.summary = (const unsigned long long*)(ob.base + kBlockStageSummary);
    {
        const int64_t cap = t->C * 32 - 1;
        const int holds = pd.split ? 4 + pd.hp : pd.rec_width;  // nucleotides a record of this submission holds
        const bool table_bounds = t->w_min > 0 && cap / t->w_min <= (int64_t)holds;  // no window value of this table is deeper
        g.max_hi = (t->w_min > 0 && !table_bounds) ? (unsigned long long)((int64_t)(holds + 1) * t->w_min - 1) : ~0ULL;
        ctx->max_hi = g.max_hi < (unsigned long long)cap ? (int64_t)g.max_hi : cap;
        const double heavy = kHeavyCostPerPeak * (double)P;
        g.max_cost = heavy < 1.8e19 ? (unsigned long long)heavy : ~0ULL;
    }
    if ((rc = dfs_enqueue(ctx, t, pd.rec_width, mp, &g, &ob))) {
        pd.active = false;
        return rc;
    }
    // the result block on its way back in ONE copy: summaries, status, peak offsets and as many records as the previous
    // batch had (+ spec_margin_pct); a batch that turns out larger gets the rest in sst_explain_collect
    if (pd.split) {  // header, status, offsets and the lo plane in one copy, then the byte planes
        CK(cudaMemcpyAsync(out_block, ob.base, (size_t)(lay.recs + 4 * guessN), cudaMemcpyDeviceToHost, ctx->stream));
        for (int k = 0; k < pd.hp; k++) {
            const size_t at = (size_t)(lay.recs + 4 * pd.capN + (uint64_t)k * pd.capN);
            CK(cudaMemcpyAsync(out_block + at, ob.base + at, (size_t)guessN, cudaMemcpyDeviceToHost, ctx->stream));
        }
        pd.copiedN = guessN;
        guess = 0;
    } else {
        CK(cudaMemcpyAsync(out_block, ob.base, (size_t)(lay.recs + guess), cudaMemcpyDeviceToHost, ctx->stream));
    }
    hp_mark(11);
    hp_mark(12);
    trace_mark(ctx, 4);
    pd.copied = guess;
    pd.done = true;  // queued
    return SST_OK;
}

uint64_t sst_explain_d2h_bytes(const sst_ctx* ctx) { return ctx->last_d2h_bytes; }

int sst_set_record_split(sst_ctx* ctx, int enable) {
    ctx->split_records = enable != 0;
    return SST_OK;
}

int sst_explain_rec_layout(const sst_ctx* ctx, int* split, uint64_t* cap_n, int* hi_planes) {
    if (split) *split = ctx->last_split ? 1 : 0;
    if (cap_n) *cap_n = ctx->last_capN;
    if (hi_planes) *hi_planes = ctx->last_hp;
    return SST_OK;
}

int sst_explain_block_layout(int64_t P, uint64_t* status_off, uint64_t* off32_off, uint64_t* recs_off) {
    if (P < 0) return SST_ERR_BAD_ARG;
    const BlockLayout lay(P);
    if (status_off) *status_off = lay.status;
    if (off32_off) *off32_off = lay.off32;
    if (recs_off) *recs_off = lay.recs;
    return SST_OK;
}

int sst_explain_collect(sst_ctx* ctx, const sst_table* t, uint64_t* n_comps, int* rec_width) {
    CK(cudaSetDevice(ctx->device));
    sst_ctx::Pending& pd = ctx->pend;
    if (!pd.active) return fail(ctx, SST_ERR_STATE, "no submitted batch to collect");
    pd.active = false;
    int rc;
    bool redo = !pd.done;
    unsigned long long roots = 0, comps = 0;
    if (pd.done) {
        hp_begin();
        CK(cudaStreamSynchronize(ctx->stream));
        hp_mark(13);
        memcpy(ctx->h_run, pd.block, 352);  // where the evaluation below looks
        memcpy(ctx->h_misc, pd.block + kBlockStageSummary, 64);
        const unsigned long long* hs = (const unsigned long long*)ctx->h_misc;  // the staging kernel's summary
        if ((rc = nf_error(ctx, hs[4]))) return rc;
        MemoMap mp{};
        rc = pd.direct ? direct_evaluate(ctx, pd.rec_width, 0, &roots, &comps) : dfs_evaluate(ctx, pd.rec_width, mp, false, 0, &roots, &comps);
        if (rc == DFS_OK && comps < (1ULL << 32)) {
            {
                const int64_t cap = t->C * 32 - 1, max_hi = (int64_t)hs[1];
                ctx->max_hi = max_hi;
                ctx->window_total = (int64_t)hs[0];
                ctx->deepest = t->w_min > 0 ? (max_hi < cap ? max_hi : cap) / t->w_min : 0;
                if (ctx->deepest <= 8) ctx->spec_rec_width = 8;  // (a batch of short compositions after longer ones)
                ctx->spec_hi_planes = ctx->deepest <= 4 ? 0 : (ctx->deepest >= 8 ? 4 : (int)ctx->deepest - 4);
            }
            ctx->n_roots = roots;
            ctx->n_comps = comps;
            ctx->rec_width = pd.rec_width;
            ctx->last_pass = pd.direct ? 3 : 2;
            ctx->have_result = false;  // (delivered; nothing is left in the buffers sst_explain_fetch reads)
            ctx->last_comps = comps;
            const uint64_t need = comps * (uint64_t)pd.rec_width;
            ctx->last_split = pd.split;
            ctx->last_capN = pd.capN;
            ctx->last_hp = pd.hp;
            if (pd.split) {  // (more compositions than capN never get here: the pass reports that the records do not fit)
                const BlockLayout lay(pd.P);
                if (comps > pd.copiedN) {
                    const uint64_t n0 = pd.copiedN, dn = comps - pd.copiedN;
                    CK(cudaMemcpyAsync(pd.block + lay.recs + 4 * n0, (const char*)ctx->d_block.p + lay.recs + 4 * n0, (size_t)(4 * dn), cudaMemcpyDeviceToHost,
                                       ctx->stream));
                    for (int k = 0; k < pd.hp; k++) {
                        const size_t at = (size_t)(lay.recs + 4 * pd.capN + (uint64_t)k * pd.capN + n0);
                        CK(cudaMemcpyAsync(pd.block + at, (const char*)ctx->d_block.p + at, (size_t)dn, cudaMemcpyDeviceToHost, ctx->stream));
                    }
                    CK(cudaStreamSynchronize(ctx->stream));
                }
                ctx->last_d2h_bytes = lay.recs + (uint64_t)(4 + pd.hp) * (comps > pd.copiedN ? comps : pd.copiedN);
            } else if (need > pd.recs_bytes) {  // the caller's block is too small: the synchronous path leaves the result on the device for sst_explain_fetch
                redo = true;
            } else {
                ctx->last_d2h_bytes = BlockLayout(pd.P).recs + (need > pd.copied ? need : pd.copied);
            }
            if (!pd.split && need <= pd.recs_bytes && need > pd.copied) {
                const BlockLayout lay(pd.P);
                CK(cudaMemcpyAsync(pd.recs + pd.copied, (const char*)ctx->d_block.p + lay.recs + pd.copied, (size_t)(need - pd.copied),
                                   cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
            }
        } else if (rc == DFS_OK || rc == DFS_RETRY || rc == PASS_FALLBACK) {
            redo = true;  // larger buffers / the other pass / 64-bit offsets: the synchronous path sorts it out
        } else {
            return rc;
        }
    }
    if (redo) {
        rc = stage_f64(ctx, t, pd.mass, pd.thr, nullptr, pd.max_mods, pd.P, pd.ind, pd.is_mod, pd.precision, pd.tolerance, pd.with_memo);
        if (rc) return rc;
        uint64_t r64 = 0, c64 = 0;
        if ((rc = sst_explain_run(ctx, t, 0, 0, &r64, &c64))) return rc;
        comps = c64;
        ctx->last_comps = comps;
        // would the speculative path have carried this batch?  Then the next submission takes it.
        ctx->spec_ok = ctx->last_pass == 2 && !ctx->n_memo && !ctx->has_exact && ctx->rec_width <= 16;
        if (ctx->spec_ok) ctx->spec_rec_width = ctx->rec_width;
        ctx->spec_hi_planes = ctx->deepest <= 4 ? 0 : (ctx->deepest >= 8 ? 4 : (int)ctx->deepest - 4);
        ctx->last_split = false;  // (whole records of rec_width bytes)
        if (comps >= (1ULL << 32)) return fail(ctx, SST_ERR_NOMEM, "%llu compositions: more than the 32-bit offsets of the asynchronous entry hold", (unsigned long long)comps);
        const uint64_t need = comps * (uint64_t)ctx->rec_width;
        if (n_comps) *n_comps = comps;
        if (rec_width) *rec_width = ctx->rec_width;
        if (need > pd.recs_bytes)
            return fail(ctx, SST_ERR_NOMEM, "record buffer of %llu bytes is too small for %llu compositions", (unsigned long long)pd.recs_bytes,
                        (unsigned long long)comps);
        // 64-bit offsets -> 32-bit on the host (this path is the exception)
        std::vector<uint64_t> off((size_t)pd.P + 1);
        if ((rc = sst_explain_fetch(ctx, pd.status, off.data(), pd.recs))) return rc;
        for (int64_t i = 0; i <= pd.P; i++) pd.off32[i] = (uint32_t)off[(size_t)i];
        ctx->last_d2h_bytes = (uint64_t)pd.P + 8 * ((uint64_t)pd.P + 1) + need;  // what sst_explain_fetch copied
    }
    if (n_comps) *n_comps = ctx->n_comps;
    if (rec_width) *rec_width = ctx->rec_width;
    return SST_OK;
}

int sst_trace_ms(sst_ctx* ctx, int enable, float* out) {
    CK(cudaSetDevice(ctx->device));
    if (!g_trace_base) {
        CK(cudaEventCreate(&g_trace_base));
        CK(cudaEventRecord(g_trace_base, ctx->stream));
        CK(cudaEventSynchronize(g_trace_base));
    }
    if (out)
        for (int k = 0; k < 8; k++) {
            out[k] = -1.f;
            if (ctx->trace_ev[k] && cudaEventElapsedTime(&out[k], g_trace_base, ctx->trace_ev[k]) != cudaSuccess) {
                out[k] = -1.f;
                cudaGetLastError();
            }
        }
    if (enable && !ctx->trace_ev[0])
        for (auto& e : ctx->trace_ev) CK(cudaEventCreate(&e));
    ctx->trace_on = enable != 0;
    return SST_OK;
}

// Compositions per call without enumerating them (k_count_compositions): what the host-side partition of a workload
// over several GPUs balances by.
int sst_count_compositions_f64(sst_ctx* ctx, sst_table* t, const double* mass, const double* thr, int64_t P, double precision, double tolerance,
                               uint64_t* counts_out) {
    CK(cudaSetDevice(ctx->device));
    if (P < 0) return fail(ctx, SST_ERR_BAD_ARG, "negative call count");
    if (!P) return SST_OK;
    int rc;
    if ((rc = ensure_counts(ctx, t))) return rc;
    if ((rc = reserve(ctx, ctx->d_vmass, (size_t)P * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vthrf, (size_t)P * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_peakoff, (size_t)(P + 2) * 8))) return rc;
    ctx->have_result = false;  // (the staging arrays and the offsets of a staged batch are overwritten)
    ctx->R_staged = -1;
    CK(cudaMemcpyAsync(ctx->d_vmass.p, mass, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (thr) CK(cudaMemcpyAsync(ctx->d_vthrf.p, thr, (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_count_compositions<<<(unsigned)((P + 255) / 256), 256, 0, ctx->stream>>>(view_of(t), CountView{t->d_cnt2d, t->Mcnt}, (const double*)ctx->d_vmass.p,
                                                                               thr ? (const double*)ctx->d_vthrf.p : nullptr, P, precision, tolerance,
                                                                               (unsigned long long*)ctx->d_peakoff.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(counts_out, ctx->d_peakoff.p, (size_t)P * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SST_OK;
}

// ---------------- N3 / N4: ladder differences and alphabet reduction on a device-resident frame (sst_ladder.cuh) ----------------
int sst_ladder_stage(sst_ctx* ctx, const double* su, const double* obs, const uint8_t* flags, int64_t F) {
    CK(cudaSetDevice(ctx->device));
    if (F < 0 || F >= ((int64_t)1 << 31)) return fail(ctx, SST_ERR_BAD_ARG, "fragment count out of range");
    ctx->LF = -1;
    int rc;
    const size_t n = (size_t)(F ? F : 1);
    if ((rc = reserve(ctx, ctx->d_lsu, n * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_lobs, n * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_lflags, n))) return rc;
    if ((rc = reserve(ctx, ctx->d_lalive, n))) return rc;
    if ((rc = reserve(ctx, ctx->d_lidx, 3 * n * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_lreach, 2 * n * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_lfirst, 2 * n * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_lhdr, kLadderHdrWords * 8))) return rc;
    if (F) {
        for (int64_t i = 1; i < F; i++)  // the window walks sorted masses (prediction.py:68-72 sorts the frame first)
            if (su[i] < su[i - 1]) return fail(ctx, SST_ERR_BAD_ARG, "standard-unit masses must ascend (fragment %lld)", (long long)i);
        CK(cudaMemcpyAsync(ctx->d_lsu.p, su, (size_t)F * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_lobs.p, obs, (size_t)F * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_lflags.p, flags, (size_t)F, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_lalive.p, 1, (size_t)F, ctx->stream));
    }
    CK(cudaMemsetAsync(ctx->d_lhdr.p, 0, kLadderHdrWords * 8, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->LF = F;
    ctx->l_calls = 0;
    return SST_OK;
}

int sst_ladder_round(sst_ctx* ctx, const sst_table* t, double max_weight, double precision, double tolerance, int32_t max_mods,
                     const int32_t* ind, const uint8_t* is_mod, int with_memo, uint32_t* mask_out, uint64_t* n_calls, uint64_t* n_comps) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->LF < 0) return fail(ctx, SST_ERR_STATE, "no fragment frame staged (sst_ladder_stage)");
    const int64_t F = ctx->LF;
    int rc;
    ctx->l_calls = 0;
    ctx->have_result = false;
    if (mask_out) mask_out[0] = mask_out[1] = mask_out[2] = mask_out[3] = 0u;
    if (n_calls) *n_calls = 0;
    if (n_comps) *n_comps = 0;
    if (!F) return SST_OK;
    LadderFrame fr{(const double*)ctx->d_lsu.p, (const double*)ctx->d_lobs.p, (const uint8_t*)ctx->d_lflags.p, (uint8_t*)ctx->d_lalive.p, F};
    unsigned long long* hdr = (unsigned long long*)ctx->d_lhdr.p;
    uint32_t* idx = (uint32_t*)ctx->d_lidx.p;
    uint32_t* reach = (uint32_t*)ctx->d_lreach.p;
    unsigned long long* first = (unsigned long long*)ctx->d_lfirst.p;
    const unsigned gx = (unsigned)((F + 255) / 256);
    k_ladder_sides<<<1, kPassThreads, 0, ctx->stream>>>(fr, idx, hdr);
    k_ladder_reach<<<dim3(gx, 2), 256, 0, ctx->stream>>>(fr, idx, hdr, max_weight, reach);
    k_ladder_scan<<<1, kPassThreads, 0, ctx->stream>>>(fr, hdr, reach, first);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_misc, hdr, kLadderHdrWords * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const unsigned long long* h = (const unsigned long long*)ctx->h_misc;
    const uint64_t calls = h[7], pairs = h[5] + h[6];
    if (calls >= ((uint64_t)1 << 32) - 1) return fail(ctx, SST_ERR_NOMEM, "%llu ladder differences: more than one batch holds", (unsigned long long)calls);
    if (n_calls) *n_calls = calls;
    if (!calls) return SST_OK;
    if ((rc = reserve(ctx, ctx->d_vmass, (size_t)calls * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_vthrf, (size_t)calls * 8))) return rc;
    k_ladder_fill<<<dim3(gx, 3), 256, 0, ctx->stream>>>(fr, idx, hdr, reach, first, tolerance, (double*)ctx->d_vmass.p, (double*)ctx->d_vthrf.p);
    CK(cudaGetLastError());
    // the calls are a staged batch now: integerise, budget modes, enumeration — nothing has left the device
    if ((rc = stage_f64(ctx, t, (const double*)ctx->d_vmass.p, (const double*)ctx->d_vthrf.p, nullptr, max_mods, (int64_t)calls, ind, is_mod, precision,
                        tolerance, with_memo, true)))
        return rc;
    uint64_t roots = 0, comps = 0;
    for (uint64_t memo_cap = 0;;) {  // (MEMO mode: a first-visit map that turns out too small is enlarged, the staged batch run again)
        rc = sst_explain_run(ctx, t, 0, memo_cap, &roots, &comps);
        if (rc != SST_ERR_MEMO_FULL) break;
        memo_cap = (memo_cap ? memo_cap : ((uint64_t)1 << 20)) * 4;
        if (memo_cap > ((uint64_t)1 << 30)) break;
    }
    if (rc) return rc;
    if (n_comps) *n_comps = comps;
    // dedup by key + union of the rows the winners use
    uint64_t pow2 = 1024;
    while (pow2 < 2 * calls) pow2 <<= 1;
    if ((rc = reserve(ctx, ctx->d_lkeys, (size_t)pow2 * 8))) return rc;
    if ((rc = reserve(ctx, ctx->d_llast, (size_t)pow2 * 4))) return rc;
    if ((rc = reserve(ctx, ctx->d_lcall, (size_t)calls))) return rc;
    CK(cudaMemsetAsync(ctx->d_lkeys.p, 0, (size_t)pow2 * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_llast.p, 0, (size_t)pow2 * 4, ctx->stream));
    KeyTable kt{(unsigned long long*)ctx->d_lkeys.p, (unsigned int*)ctx->d_llast.p, (uint32_t)(pow2 - 1)};
    const unsigned gc = (unsigned)((calls + 255) / 256);
    k_ladder_enter<<<gc, 256, 0, ctx->stream>>>((const double*)ctx->d_vmass.p, (const unsigned long long*)ctx->d_peakoff.p,
                                                (const uint8_t*)ctx->d_status.p, calls, pairs, kt, (uint8_t*)ctx->d_lcall.p, hdr);
    k_ladder_union<<<gc, 256, 0, ctx->stream>>>((const double*)ctx->d_vmass.p, (const unsigned long long*)ctx->d_peakoff.p,
                                                (const unsigned long long*)ctx->d_recs.p, ctx->rec_width / 8, calls, kt, (uint8_t*)ctx->d_lcall.p, hdr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_misc, hdr, kLadderHdrWords * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (mask_out) memcpy(mask_out, h + 8, 16);
    ctx->l_calls = calls;
    if (h[14])  // explain_mass_with_table raises for such a call (mass_explanation.py:134-138), and with it the whole round upstream
        return fail(ctx, SST_ERR_OUT_OF_TABLE, "A value of the mass window is not in the DP table. Extend its size if you want to compute larger masses.");
    return SST_OK;
}

int sst_ladder_revalidate(sst_ctx* ctx, const sst_table* t, double precision, double tolerance, int64_t* n_alive) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->LF < 0) return fail(ctx, SST_ERR_STATE, "no fragment frame staged (sst_ladder_stage)");
    const int64_t F = ctx->LF;
    if (n_alive) *n_alive = 0;
    if (!F) return SST_OK;
    LadderFrame fr{(const double*)ctx->d_lsu.p, (const double*)ctx->d_lobs.p, (const uint8_t*)ctx->d_lflags.p, (uint8_t*)ctx->d_lalive.p, F};
    unsigned long long* hdr = (unsigned long long*)ctx->d_lhdr.p;
    CK(cudaMemsetAsync(hdr + 12, 0, 16, ctx->stream));
    k_ladder_revalidate<<<(unsigned)((F + 255) / 256), 256, 0, ctx->stream>>>(view_of(t), fr, precision, tolerance, hdr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ctx->h_misc, hdr, kLadderHdrWords * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const unsigned long long* h = (const unsigned long long*)ctx->h_misc;
    if (h[13]) return fail(ctx, SST_ERR_OUT_OF_TABLE, "A value of the mass window is not in the DP table. Extend its size if you want to compute larger masses.");
    if (n_alive) *n_alive = (int64_t)h[12];
    return SST_OK;
}

int sst_ladder_fetch(sst_ctx* ctx, uint8_t* alive_out, double* key_out, double* thr_out, uint8_t* call_flags_out) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->LF < 0) return fail(ctx, SST_ERR_STATE, "no fragment frame staged (sst_ladder_stage)");
    if (alive_out && ctx->LF) CK(cudaMemcpyAsync(alive_out, ctx->d_lalive.p, (size_t)ctx->LF, cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->l_calls) {
        if (key_out) CK(cudaMemcpyAsync(key_out, ctx->d_vmass.p, (size_t)ctx->l_calls * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (thr_out) CK(cudaMemcpyAsync(thr_out, ctx->d_vthrf.p, (size_t)ctx->l_calls * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (call_flags_out) CK(cudaMemcpyAsync(call_flags_out, ctx->d_lcall.p, (size_t)ctx->l_calls, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return SST_OK;
}

int sst_host_profile(int enable, uint64_t* ns_out, uint64_t* calls_out) {
    if (ns_out) memcpy(ns_out, g_hp.ns, sizeof(g_hp.ns));
    if (calls_out) memcpy(calls_out, g_hp.calls, sizeof(g_hp.calls));
    memset(g_hp.ns, 0, sizeof(g_hp.ns));
    memset(g_hp.calls, 0, sizeof(g_hp.calls));
    g_hp.on = enable != 0;
    return SST_OK;
}

int sst_explain_phase_ns(const sst_ctx* ctx, uint64_t* out) {
    for (int i = 0; i < 32; i++) out[i] = ctx->phase_ns[i];
    return SST_OK;
}

int sst_explain_rec_width(const sst_ctx* ctx) { return ctx->rec_width; }

int sst_explain(sst_ctx* ctx, const sst_table* t, const int64_t* target, const int64_t* thr, const int32_t* max_mods,
                const uint8_t* mode, int64_t P, const int32_t* ind, const uint8_t* is_mod, int rec_width,
                uint64_t memo_capacity, uint64_t* n_roots, uint64_t* n_comps) {
    int rc = sst_explain_stage(ctx, t, target, thr, max_mods, mode, P, ind, is_mod);
    if (rc) return rc;
    return sst_explain_run(ctx, t, rec_width, memo_capacity, n_roots, n_comps);
}

int sst_explain_fetch(sst_ctx* ctx, uint8_t* status, uint64_t* peak_off, uint8_t* recs) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->have_result) return fail(ctx, SST_ERR_STATE, "no enumeration result to fetch");
    if (status && ctx->P) CK(cudaMemcpyAsync(status, ctx->d_status.p, (size_t)ctx->P, cudaMemcpyDeviceToHost, ctx->stream));
    if (peak_off) CK(cudaMemcpyAsync(peak_off, ctx->d_peakoff.p, (size_t)(ctx->P + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (recs && ctx->n_comps)
        CK(cudaMemcpyAsync(recs, ctx->d_recs.p, (size_t)ctx->n_comps * ctx->rec_width, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return SST_OK;
}

}  // extern "C"
