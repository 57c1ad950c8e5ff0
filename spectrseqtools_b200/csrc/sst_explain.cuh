// K2a validity probe, K3 integer window + K2b enumerator (one cooperative pass, + first-visit budget replay),
// staging of float batches, N2 fragment classification, N1 sequence-length bounds.
//
// Replaces is_valid_mass (reference mass_explanation.py:45-89), explain_mass_with_table (:92-203, inner backtrack
// :118-188), the per-pair callbacks of classify_fragments (fragment_classification.py:39-82) and
// compute_sequence_length_bound (mass_table.py:343-487).  Kernels work on integer masses; the float -> integer
// conversions (k_stage_f64, k_is_valid_f64, k_classify) use the same IEEE operations as the reference's Python
// (division, round-half-even, ceil; no FMA contraction), so they match CPython bit for bit.
//
// Enumeration model.  The reference walks (mass m, row r): UP to (m, r-1) if bit0, LEFT to (m-w_r, r) if
// bit1.  Because bit0(i,m) = OR_{r<i} bit1(r,m) (plus mass 0), the children of "mass m, rows <= rmax" are
// exactly {(m - w_r, r) : r <= rmax, bit1(r, m)}: one 16-byte load of the mass-major mask H[m] replaces
// the ~100 dependent UP reads per nucleotide.  A composition is emitted when the remainder hits 0.
//
// Budget modes (per peak):
//   FREE   budgets cannot bind (closed-form test when the batch is staged): enabled edges = table bits.
//   EXACT  with_memo=False: every path carries its own budgets (global `all`, per-row `ind`).
//   MEMO   with_memo=True with binding budgets: the reference's memo is keyed (m, r) WITHOUT budgets, so
//          each node is expanded once with the budgets of its first arrival in DFS order (UP before
//          LEFT, window ascending).  Phase A replays that order sequentially per peak, one mass at a
//          time (rows visited at a mass always form a contiguous range [r0(m), top(m)]), and records per
//          mass the LEFT edges that were enabled AND lead to at least one solution.  The enumeration pass then
//          reads those masks from a hash map instead of H.
#pragma once
#include "sst_common.cuh"

namespace sst {

enum : int { MODE_FREE = 0, MODE_EXACT = 1, MODE_MEMO = 2 };
enum : int { ST_ZERO_IN_WINDOW = 1, ST_OUT_OF_TABLE = 2 };
constexpr int kBudgetInf = 1 << 30;

struct TableView {
    const uint64_t* tbl;  // R x C, row-major (reference layout)
    const uint4* H;       // C*32 row masks
    const uint32_t* any;  // last-row summary: bit (j & 31) of any[j >> 5] = word j of the last row is non-zero (may be null)
    const uint8_t* wbucket;  // wbucket[k] = first row whose weight is >= k * 1024, k <= w_top / 1024 (may be null)
    const int32_t* weights;
    int R;
    int64_t C;
};

struct RowMeta {  // per-call budget metadata, device arrays of length R
    const int32_t* ind;      // IND[r] = round(max_len * rate_r) (host, Python round)
    const uint8_t* is_mod;   // 1 if the row is a modification
};

struct PeakBatch {
    const int64_t* target;
    const int64_t* thr;
    const int32_t* max_mods;  // kBudgetInf for "unbounded"
    const uint8_t* mode;
    int64_t P;
};

// ---------------- first-visit memo map (MEMO mode) ----------------
struct MemoMap {
    unsigned long long* keys;  // 0 = empty, else ((peak+1) << 32) | mass
    uint4* alive;              // enabled-and-productive LEFT edges of the visited rows
    uint32_t* top;             // highest visited row at this mass
    uint32_t cap_mask;         // capacity - 1 (power of two)
    unsigned int* fill;        // occupied slots
    int* overflow;             // set when the map is too small
};

__device__ __forceinline__ uint64_t mix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ unsigned long long memo_key(int64_t peak, uint32_t m) {
    return ((unsigned long long)(peak + 1) << 32) | m;
}
__device__ inline int memo_find(const MemoMap& mp, unsigned long long key) {
    uint32_t h = (uint32_t)mix64(key) & mp.cap_mask;
    for (uint32_t probes = 0; probes <= mp.cap_mask; probes++) {
        unsigned long long k = mp.keys[h];
        if (k == key) return (int)h;
        if (k == 0ULL) return -1;
        h = (h + 1) & mp.cap_mask;
    }
    return -1;
}
// `inserted` counts the slots this thread has claimed; the caller adds it to mp.fill ONCE when it is done (one
// atomic per claimed slot on a single counter serialised the whole replay) and the host compares the total with
// the load-factor limit after the launch.
__device__ inline int memo_find_or_insert(const MemoMap& mp, unsigned long long key, unsigned int& inserted) {
    uint32_t h = (uint32_t)mix64(key) & mp.cap_mask;
    for (uint32_t probes = 0; probes < 4096; probes++) {
        unsigned long long k = mp.keys[h];
        if (k == key) return (int)h;
        if (k == 0ULL) {
            unsigned long long old = atomicCAS(mp.keys + h, 0ULL, key);
            if (old == 0ULL) {
                inserted++;
                return (int)h;  // fresh slot: alive = 0, top = 0 (buffers are zeroed by the host)
            }
            if (old == key) return (int)h;
        }
        h = (h + 1) & mp.cap_mask;
    }
    *mp.overflow = 1;
    return -1;
}

// ---------------- K2a: validity ----------------
// Float -> integer conversion on the device with the same IEEE operations the reference does on the host
// (mass_explanation.py:51-58): true division by `precision`, round-half-even, ceil.  No FMA contraction.
__device__ __forceinline__ void integerise(double mass, double thr, double precision, double tolerance, int64_t& target, int64_t& ithr) {
    target = (int64_t)rint(__ddiv_rn(mass, precision));
    const double t = isnan(thr) ? __dmul_rn(tolerance, mass) : thr;  // NaN = "threshold None" -> relative
    ithr = (int64_t)ceil(__ddiv_rn(t, precision));
}

// Non-finite inputs: the reference integerises with int(round(x)) / int(np.ceil(x)) (mass_explanation.py:51-58,107-114),
// which raises ValueError for NaN and OverflowError for an infinity.  The kernels that integerise report them instead of
// casting: a NaN mass, an infinite mass (its relative threshold is infinite too) or an infinite threshold.  A NaN
// threshold means "None" (relative) in the batched entries.
enum : int { NF_NAN = 1, NF_INF = 2 };
enum : int { VALID_CODE_NAN = 8, VALID_CODE_INF = 9 };  // k_is_valid_f64's codes for them (0 / 1 / 2 are answers; bit 3 is set by nothing else)
enum : int { CLASS_CODE_NAN = 8, CLASS_CODE_INF = 9 };  // k_classify's (bit 3 is set by nothing else)
__device__ __forceinline__ int non_finite(double mass, double thr) {
    return (isnan(mass) ? NF_NAN : 0) | ((isinf(mass) || isinf(thr)) ? NF_INF : 0);
}

// out[p] = 0 not valid, 1 valid, 2 out-of-table value met before any hit (-> NotImplementedError)
__device__ __forceinline__ uint8_t valid_code(const TableView& tv, int64_t target, int64_t thr) {
    const int64_t limit = tv.C * 32;
    const uint64_t* last = tv.tbl + (int64_t)(tv.R - 1) * tv.C;
    const int64_t lo = target - thr, hi = target + thr;
    const int64_t a = lo < 1 ? 1 : lo;
    const int64_t b = hi < limit - 1 ? hi : limit - 1;
    if (a <= b) {
        for (int64_t wd = a >> 5; wd <= (b >> 5); wd++) {
            uint64_t x = __ldg(last + wd);
            x = (x | (x >> 1)) & kBit0Mask;
            if (wd == (a >> 5)) x &= (1ULL << (2 * (31 - (int)(a & 31)) + 1)) - 1ULL;
            if (wd == (b >> 5)) x &= ~0ULL << (2 * (31 - (int)(b & 31)));
            if (x) return 1;
        }
    }
    return (hi >= limit && hi >= 1 && lo <= hi) ? 2 : 0;
}

__global__ void k_is_valid(TableView tv, const int64_t* __restrict__ target, const int64_t* __restrict__ thr,
                           int64_t P, uint8_t* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) out[p] = valid_code(tv, target[p], thr[p]);
}

__global__ void k_is_valid_f64(TableView tv, const double* __restrict__ mass, const double* __restrict__ thr,
                               double precision, double tolerance, int64_t P, uint8_t* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double m = mass[p], th = thr ? thr[p] : nan("");
    const int bad = non_finite(m, th);
    if (bad) {  // the reference's int(round(nan)) / int(round(inf)): the host turns the code into the exception
        out[p] = (uint8_t)(bad & NF_NAN ? VALID_CODE_NAN : VALID_CODE_INF);
        return;
    }
    int64_t t, h;
    integerise(m, th, precision, tolerance, t, h);
    out[p] = valid_code(tv, t, h);
}

// ---------------- scheduling cost of a peak ----------------
// The depth-first pass deals the batch to CTAs by ESTIMATED work, not by peak count (one 5-nt ladder gap has 10^4
// compositions, a 1-nt one has one).  The estimate needs no table access: lamq[k] is the number of compositions per
// unit of mass around mass k * width, from a coarse coin-change count over the row weights made by the host when the
// table is created (Q16, times an enrichment factor: the mass of a real fragment difference sits where compositions
// cluster).  Window values that are reachable ~ W * min(1, lam), compositions ~ W * lam.
struct CostModel {
    const uint32_t* lamq;
    uint32_t width;
    uint32_t K;
};
constexpr int kCostBlock = 256;  // peaks per entry of the block-cost array
__device__ __forceinline__ uint32_t peak_cost(const CostModel& cm, int64_t target, int64_t thr, int64_t limit) {
    const int64_t lo = target - thr, hi = target + thr;
    const int64_t a = lo < 1 ? 1 : lo, b = hi < limit - 1 ? hi : limit - 1;
    const unsigned long long W = b >= a ? (unsigned long long)(b - a + 1) : 0ULL;
    unsigned long long cost = 16ULL + (W >> 4);
    if (W && cm.lamq) {
        const int64_t mid = (a + b) >> 1;
        unsigned long long k = ((unsigned long long)mid + (cm.width >> 1)) / cm.width;
        if (k >= cm.K) k = cm.K - 1;
        const unsigned long long lam = __ldg(cm.lamq + k);
        const unsigned long long n_est = (W * (lam < 65536ULL ? lam : 65536ULL)) >> 16;
        unsigned long long c_est = (W * lam) >> 16;
        if (c_est < n_est) c_est = n_est;
        cost += 6ULL * n_est + 4ULL * c_est;
    }
    return cost < 0x7FFFFFFFULL ? (uint32_t)cost : 0x7FFFFFFFu;
}
// CTA-wide: per-peak costs -> cost[p], their sum -> blk[blockIdx.x]   (blockDim.x == kCostBlock)
__device__ __forceinline__ void store_costs(uint32_t c, int64_t p, int64_t P, uint32_t* __restrict__ cost, unsigned long long* __restrict__ blk) {
    __shared__ unsigned long long s_blk;
    if (threadIdx.x == 0) s_blk = 0ULL;
    __syncthreads();
    if (p < P) cost[p] = c;
    unsigned long long v = p < P ? c : 0u;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_blk, v);
    __syncthreads();
    if (threadIdx.x == 0) blk[blockIdx.x] = s_blk;
}
__global__ void __launch_bounds__(kCostBlock)
k_peak_costs(const int64_t* __restrict__ target, const int64_t* __restrict__ thr, int64_t P, int64_t limit, CostModel cm,
             uint32_t* __restrict__ cost, unsigned long long* __restrict__ blk) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    store_costs(p < P ? peak_cost(cm, target[p], thr[p], limit) : 0u, p, P, cost, blk);
}

// ---------------- staging of a float batch: integerise + budget mode + batch summary on the device ----------------
// Same float operations as mass_explanation.py:107-114 (integerise above).  Mode: FREE when no composition inside
// the window can exhaust a budget (max_mods >= hi / w_min_mod and hi < hi_limit, both from the host), else `slow`.
// summary: [0] summed window sizes (upper bound for level-0 items), [1] largest window end, [2] MEMO peaks (their
// indices are appended to memo_peaks), [3] EXACT peaks, [4] NF_* bits of non-finite inputs, [5] summed peak_cost().
__global__ void __launch_bounds__(256)
k_stage_f64(const double* __restrict__ mass, const double* __restrict__ thr, int32_t* __restrict__ max_mods, int32_t uniform_mods,
            int use_uniform, int64_t P, double precision, double tolerance, int64_t w_min_mod, int64_t hi_limit, int slow, int64_t limit,
            int64_t* __restrict__ target, int64_t* __restrict__ ithr, uint8_t* __restrict__ mode, uint32_t* __restrict__ memo_peaks,
            unsigned long long* __restrict__ summary, CostModel cm, uint32_t* __restrict__ cost, unsigned long long* __restrict__ blk) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long win = 0, hi_u = 0;
    uint32_t my_cost = 0;
    if (p < P) {
        int64_t t, h;
        const double m = mass[p], th = thr ? thr[p] : nan("");
        const int bad = non_finite(m, th);
        if (bad) {  // reported in the summary; the peak itself becomes an empty window
            atomicOr(summary + 4, (unsigned long long)bad);
            t = 0;
            h = -1;
        } else {
            integerise(m, th, precision, tolerance, t, h);
        }
        target[p] = t;
        ithr[p] = h;
        int64_t hi = t + h;
        const int64_t lo = t - h;
        if (h >= 0) {
            const int64_t a = lo < 1 ? 1 : lo, b = hi < limit - 1 ? hi : limit - 1;
            if (b >= a) win = (unsigned long long)(b - a + 1);
            if (hi > 0) hi_u = (unsigned long long)hi;
        }
        if (hi < 0) hi = 0;
        if (use_uniform) max_mods[p] = uniform_mods;  // one budget for the whole batch: filled here instead of copied in
        const int64_t mm = use_uniform ? (int64_t)uniform_mods : (int64_t)max_mods[p];
        const bool free_ok = !w_min_mod || (mm >= hi / w_min_mod && hi < hi_limit);
        const int md = free_ok ? MODE_FREE : slow;
        mode[p] = (uint8_t)md;
        if (md == MODE_MEMO) memo_peaks[atomicAdd(summary + 2, 1ULL)] = (uint32_t)p;
        if (md == MODE_EXACT) atomicAdd(summary + 3, 1ULL);
        my_cost = peak_cost(cm, t, h, limit);
    }
    store_costs(my_cost, p, P, cost, blk);
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        win += __shfl_xor_sync(0xFFFFFFFFu, win, o);
        const unsigned long long y = __shfl_xor_sync(0xFFFFFFFFu, hi_u, o);
        hi_u = y > hi_u ? y : hi_u;
    }
    unsigned long long csum = my_cost;
#pragma unroll
    for (int o = 16; o; o >>= 1) csum += __shfl_xor_sync(0xFFFFFFFFu, csum, o);
    if ((threadIdx.x & 31) == 0) {
        if (win) atomicAdd(summary, win);
        if (hi_u) atomicMax(summary + 1, hi_u);
        atomicAdd(summary + 5, csum);
    }
}

// ---------------- N2: fused fragment classification ----------------
// One thread per fragment f and group of kClassifyPerThread breakage offsets (blockIdx.y), output breakage-major
// like the reference's concat (fragment_classification.py:39-49): standard-unit mass = observed - offset_b (offset_b =
// breakage weight * precision, multiplied on the host exactly as the reference does), threshold = tolerance *
// observed (integerised once per fragment), then the validity probe (:52-67 -> is_valid_mass) and the singleton test
// (:104-119: some row weight, 0 included, inside the window).  The probes of a thread are independent: their
// last-row loads overlap.  out bits: 1 valid, 2 out-of-table value met before any hit (the reference raises), 4 singleton.
constexpr int kClassifyPerThread = 3;

// The first version (one thread per probe, window words scanned with early exits, a binary search over the row
// weights per probe) was bound by the integer pipe and by per-lane dependent loads (ncu: ALU 68 % of peak, 137 integer
// instructions per probe).  Now a thread whose three probes fit 32-bit arithmetic takes a staged path:
//   * the interior words of a window are answered by the last-row summary (one bit per table word, 86 KB for the
//     full table: L1 resident) - above a few nucleotides nearly every word holds a reachable mass, so most probes
//     never touch the table itself;
//   * only when the interior is empty are the two end words loaded and masked to the window;
//   * the singleton test (some row weight inside the window) starts from a bucket index of the weights
//     (first row at or above every multiple of 1024) instead of a binary search;
//   * every stage is done for the three probes together, so their loads are in flight at the same time.
// Anything else (negative or huge masses, NaN, tables without the side arrays) takes the general 64-bit path.
constexpr int kWeightBucketShift = 10;

// last-row summary for the probes above: one bit per table word
__global__ void __launch_bounds__(256)
k_last_row_summary(const uint64_t* __restrict__ last, int64_t C, uint32_t* __restrict__ any) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool nz = j < C && last[j] != 0ULL;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, nz);
    if ((threadIdx.x & 31) == 0 && j < C) any[j >> 5] = m;
}

__global__ void __launch_bounds__(256)
k_classify(TableView tv, const double* __restrict__ observed, int64_t F, const double* __restrict__ offsets, int B,
           double precision, double tolerance, uint8_t* __restrict__ out, int pack4) {
    constexpr int K = kClassifyPerThread;
    const int32_t* __restrict__ s_w = tv.weights;  // <= 512 B, read-only: L1 resident, no per-CTA staging and no barrier
    const int64_t f_real = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t f = f_real < F ? f_real : F - 1;  // lanes past the end repeat the last fragment (whole warps reach the packing shuffle)
    const double obs = observed[f];
    const int64_t h64 = (int64_t)ceil(__ddiv_rn(__dmul_rn(tolerance, obs), precision));  // the same for every offset
    const int b0 = blockIdx.y * K;
    const uint64_t* __restrict__ last = tv.tbl + (int64_t)(tv.R - 1) * tv.C;
    const int limit = (int)(tv.C * 32);  // < 2^31 (checked when the table is allocated)
    int64_t t64[K];
    bool lean = tv.any && tv.wbucket && h64 >= 0 && h64 < (1 << 24);
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int b = b0 + k < B ? b0 + k : B - 1;
        t64[k] = (int64_t)rint(__ddiv_rn(__dsub_rn(obs, __ldg(offsets + b)), precision));
        lean = lean && t64[k] > -(1LL << 30) && t64[k] < (1LL << 30);
    }
    uint8_t code[K];
    if (!isfinite(obs)) {
#pragma unroll
        for (int k = 0; k < K; k++) code[k] = (uint8_t)(isnan(obs) ? CLASS_CODE_NAN : CLASS_CODE_INF);
    } else if (lean) {
        const int h = (int)h64;
        const int w_top = __ldg(s_w + tv.R - 1);
        int lo[K], hi[K], a[K], e[K];
        uint32_t inner[K];  // summary bits of the window's interior words
        int x[K];           // singleton search position
#pragma unroll
        for (int k = 0; k < K; k++) {
            lo[k] = (int)t64[k] - h;
            hi[k] = (int)t64[k] + h;
            a[k] = lo[k] < 1 ? 1 : lo[k];
            e[k] = hi[k] < limit - 1 ? hi[k] : limit - 1;
            inner[k] = 0u;
            const int i0 = (a[k] >> 5) + 1, i1 = (e[k] >> 5) - 1;  // interior words (a <= e is implied by i0 <= i1)
            if (i0 <= i1) {
                const int q0 = i0 >> 5, q1 = i1 >> 5;
                uint32_t m0 = __ldg(tv.any + q0) & (~0u << (i0 & 31));
                if (q1 == q0) {
                    m0 &= ~0u >> (31 - (i1 & 31));
                } else {
                    m0 |= __ldg(tv.any + q1) & (~0u >> (31 - (i1 & 31)));
                    for (int q = q0 + 1; q < q1; q++) m0 |= __ldg(tv.any + q);  // windows wider than 1024 masses
                }
                inner[k] = m0;
            }
            // first row at or above the bucket of lo (rows ascend; row 0 has weight 0)
            const int lo_c = lo[k] < 0 ? 0 : lo[k];
            x[k] = (lo_c <= w_top) ? (int)__ldg(tv.wbucket + (lo_c >> kWeightBucketShift)) : tv.R;
        }
        uint64_t ends[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            ends[k] = 0ULL;
            if (!inner[k] && a[k] <= e[k]) {  // the two end words decide, masked to the window
                const int w0 = a[k] >> 5, w1 = e[k] >> 5;
                const uint64_t mf = ~0ULL >> (2 * (a[k] & 31));         // cells of masses >= a in word w0
                const uint64_t ml = ~0ULL << (2 * (31 - (e[k] & 31)));  // cells of masses <= e in word w1
                const uint64_t xf = __ldg(last + w0), xl = __ldg(last + w1);
                ends[k] = (w0 == w1) ? (xf & mf & ml) : ((xf & mf) | (xl & ml));
            }
        }
#pragma unroll
        for (int k = 0; k < K; k++) {
            const bool hit = inner[k] != 0u || ends[k] != 0ULL;
            code[k] = hit ? 1 : ((hi[k] >= limit && hi[k] >= 1 && lo[k] <= hi[k]) ? 2 : 0);
            // singleton: some row weight (0 included) inside [lo, hi]
            if (hi[k] >= 0 && lo[k] <= hi[k]) {
                int xx = x[k];
                while (xx < tv.R && __ldg(s_w + xx) < lo[k]) xx++;
                if (xx < tv.R && __ldg(s_w + xx) <= hi[k]) code[k] |= 4;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; k++) {
            code[k] = valid_code(tv, t64[k], h64);  // 0 / 1 / 2
            const int64_t lo = t64[k] - h64, hi = t64[k] + h64;
            int x = 0, z = tv.R;  // first index with w >= lo
            while (x < z) {
                const int mid = (x + z) >> 1;
                if ((int64_t)__ldg(s_w + mid) < lo) x = mid + 1;
                else z = mid;
            }
            if (x < tv.R && (int64_t)__ldg(s_w + x) <= hi) code[k] |= 4;
        }
    }
    if (!pack4) {
#pragma unroll
        for (int k = 0; k < K; k++)
            if (b0 + k < B && f_real < F) out[(int64_t)(b0 + k) * F + f] = code[k];
    } else {  // two flags per byte: fragment f in the low nibble of byte (b * Fp + f) / 2, Fp = F rounded up to even
        const int64_t Fp = (F + 1) & ~1LL;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const unsigned mine = code[k], next = __shfl_down_sync(0xFFFFFFFFu, mine, 1);
            if (b0 + k < B && f_real < F && !(f_real & 1)) out[((int64_t)(b0 + k) * Fp + f_real) >> 1] = (uint8_t)(mine | ((f_real + 1 < F ? next : 0u) << 4));
        }
    }
}

// ---------------- MEMO phase A: sequential first-visit replay, one thread per peak ----------------
// The explicit stack (one frame per nucleotide on the current path, `frames` of them) lives in DYNAMIC SHARED MEMORY,
// field-major and thread-minor, not in local memory: a per-thread stack decides how much local memory the driver
// reserves for EVERY resident thread of the device (4.8 KB x 303 104 threads = 1.5 GB for 96 frames) and the first
// launch of such a kernel in a process pays for resizing that pool (hundreds of ms).  The host sizes the CTA so that
// frames x threads x kReplayFrameBytes fits (replay_threads()).
constexpr int kReplayFrameBytes = 52;  // m, slot, all, ind (4 each), pend, fresh (16 each), rin, cur (1 each), padded
struct ReplayStack {
    uint32_t* m;
    int* slot;
    int* all;
    int* ind;
    uint32_t* pend;   // [frames][4][T]
    uint32_t* fresh;  // [frames][4][T]
    uint8_t* rin;
    uint8_t* cur;
    int T, tid;
    __device__ ReplayStack(unsigned char* base, int frames, int threads, int thread) : T(threads), tid(thread) {
        const size_t n = (size_t)frames * threads;
        m = reinterpret_cast<uint32_t*>(base);
        slot = reinterpret_cast<int*>(m + n);
        all = slot + n;
        ind = all + n;
        pend = reinterpret_cast<uint32_t*>(ind + n);
        fresh = pend + 4 * n;
        rin = reinterpret_cast<uint8_t*>(fresh + 4 * n);
        cur = rin + n;
    }
    __device__ __forceinline__ int at(int d) const { return d * T + tid; }
    __device__ __forceinline__ Mask128 get(const uint32_t* plane, int d) const {
        Mask128 r;
#pragma unroll
        for (int k = 0; k < 4; k++) r.w[k] = plane[(d * 4 + k) * T + tid];
        return r;
    }
    __device__ __forceinline__ void put(uint32_t* plane, int d, const Mask128& v) const {
#pragma unroll
        for (int k = 0; k < 4; k++) plane[(d * 4 + k) * T + tid] = v.w[k];
    }
    __device__ __forceinline__ void set_bit(uint32_t* plane, int d, int r) const { plane[(d * 4 + (r >> 5)) * T + tid] |= 1u << (r & 31); }
};
// threads per CTA of a replay kernel whose deepest walk needs `frames` frames (0: does not fit at all)
inline int replay_threads(int64_t frames, size_t smem_limit) {
    for (int t = 64; t >= 8; t >>= 1)
        if ((size_t)frames * t * kReplayFrameBytes + 2048 <= smem_limit) return t;
    return 0;
}

__global__ void __launch_bounds__(64)
k_memo_phase_a(TableView tv, RowMeta meta, PeakBatch pk, const uint32_t* __restrict__ memo_peaks, int n_memo,
               MemoMap mp, int frames) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ int32_t s_w[kMaxRows];
    __shared__ int32_t s_ind[kMaxRows];
    __shared__ uint8_t s_mod[kMaxRows];
    for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) {
        s_w[i] = i < tv.R ? tv.weights[i] : 0;
        s_ind[i] = i < tv.R ? meta.ind[i] : 0;
        s_mod[i] = i < tv.R ? meta.is_mod[i] : 0;
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_memo) return;
    const int64_t p = memo_peaks[i];
    const int64_t limit = tv.C * 32;
    const uint64_t* last = tv.tbl + (int64_t)(tv.R - 1) * tv.C;
    const int64_t lo = pk.target[p] - pk.thr[p], hi = pk.target[p] + pk.thr[p];
    const int64_t a = lo < 1 ? 1 : lo;
    const int64_t b = hi < limit - 1 ? hi : limit - 1;
    const int top_row = tv.R - 1;

    // explicit recursion stack (one frame per mass on the current path), in shared memory
    ReplayStack f(s_dyn, frames, blockDim.x, threadIdx.x);
    int sp = 0;

    // arrival at (m, r_in) with budgets; either answers from the map (returns false, sets alive) or
    // opens a frame for the rows (top(m), r_in] that this arrival visits for the first time
    unsigned int inserted = 0;
    auto arrive = [&](uint32_t m, int r_in, int all, int ind, bool& alive) -> bool {
        alive = false;
        const int slot = memo_find_or_insert(mp, memo_key(p, m), inserted);
        if (slot < 0 || sp >= frames) {
            *mp.overflow = 1;
            return false;
        }
        const int top = (int)mp.top[slot];
        if (top >= r_in) {
            Mask128 A = mk(mp.alive[slot]);
            mask_keep_le(A, r_in);
            alive = !mask_empty(A);
            return false;
        }
        Mask128 pend = mk(ld_nc_u4(tv.H + m));
        mask_keep_le(pend, r_in);
        mask_keep_gt(pend, top);
        {   // the walk below visits the children one after the other, each a dependent chain of a map probe and a row-mask
            // load: ask for all of them now (fire-and-forget, into L2), so that they are there when their turn comes
            Mask128 ahead = pend;
            while (!mask_empty(ahead)) {
                const int r = mask_pop_lowest(ahead);
                const uint32_t m2 = m - (uint32_t)s_w[r];
                if (m2 == 0u || m2 > m) continue;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(tv.H + m2));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(mp.keys + ((uint32_t)mix64(memo_key(p, m2)) & mp.cap_mask)));
            }
        }
        const int o = f.at(sp);
        f.m[o] = m; f.slot[o] = slot; f.rin[o] = (uint8_t)r_in; f.all[o] = all; f.ind[o] = ind;
        f.put(f.pend, sp, pend);
        f.put(f.fresh, sp, Mask128{{0u, 0u, 0u, 0u}});
        sp++;
        return true;
    };

    for (int64_t v = a; v <= b; v++) {
        if (cell_bits(__ldg(last + (v >> 5)), v) == 0u) continue;
        bool alive;
        if (!arrive((uint32_t)v, top_row, pk.max_mods[p], s_ind[top_row], alive)) continue;
        while (sp > 0) {
            const int d = sp - 1, o = f.at(d);
            Mask128 pend = f.get(f.pend, d);
            if (mask_empty(pend)) {  // all new rows of this mass expanded: publish and return
                const int slot = f.slot[o];
                const Mask128 fresh = f.get(f.fresh, d);
                uint4 A = mp.alive[slot];
                A.x |= fresh.w[0]; A.y |= fresh.w[1]; A.z |= fresh.w[2]; A.w |= fresh.w[3];
                mp.alive[slot] = A;
                mp.top[slot] = f.rin[o];
                Mask128 t = mk(A);
                mask_keep_le(t, f.rin[o]);
                const bool ok = !mask_empty(t);
                sp--;
                if (sp > 0 && ok) f.set_bit(f.fresh, sp - 1, f.cur[f.at(sp - 1)]);
                continue;
            }
            const int r = mask_pop_lowest(pend);  // LEFT edges fire in ascending row order (UP first)
            f.put(f.pend, d, pend);
            f.cur[o] = (uint8_t)r;
            const int ind_here = (r == f.rin[o]) ? f.ind[o] : s_ind[r];
            const int mod = s_mod[r];
            if (mod && !(f.all[o] > 0 && ind_here > 0)) continue;
            const uint32_t m2 = f.m[o] - (uint32_t)s_w[r];
            if (m2 == 0u) {
                f.set_bit(f.fresh, d, r);
                continue;
            }
            bool child_alive;
            if (!arrive(m2, r, f.all[o] - mod, ind_here - mod, child_alive)) {
                if (child_alive) f.set_bit(f.fresh, d, r);
            }
        }
    }
    if (inserted) atomicAdd(mp.fill, inserted);
}

// ---------------- N1: sequence-length bounds (compute_sequence_length_bound, reference mass_table.py:343-487) ----------------
// Smallest / largest number of nucleotides over all explanations of the window, with the reference's memo
// semantics: the memo is keyed (mass, row) WITHOUT the budgets, so every node keeps the value of its FIRST visit
// in DFS order (UP before LEFT, window ascending), and a dead end reached through LEFT contributes default + 1
// (for the upper bound that is 0, not -1).  First-visit order makes this inherently sequential: ONE thread walks
// the same mass-at-a-time replay as k_memo_phase_a — rows visited at a mass always form a prefix [1, top(m)], and
// node (m, r) is the running min / max over the enabled LEFT candidates of rows <= r — and computes both bounds in
// the same walk (the traversal does not depend on the direction).
struct BoundMap {
    uint32_t* keys;   // 0 = empty, else the mass
    uint8_t* top;     // highest visited row
    int8_t* lower;    // [cap][kMaxRows] value of node (m, r) for r <= top, "lower" direction
    int8_t* upper;    // the same for "upper"
    uint32_t cap_mask;
    int* overflow;
};

__device__ inline int bound_slot(const BoundMap& mp, uint32_t m, unsigned int* fill) {
    uint32_t h = (uint32_t)mix64(m) & mp.cap_mask;
    for (uint32_t probes = 0; probes < 4096; probes++) {
        const uint32_t k = mp.keys[h];
        if (k == m) return (int)h;
        if (k == 0u) {
            mp.keys[h] = m;  // single writer
            if (++(*fill) > (mp.cap_mask >> 1) + (mp.cap_mask >> 2)) *mp.overflow = 1;
            return (int)h;
        }
        h = (h + 1) & mp.cap_mask;
    }
    *mp.overflow = 1;
    return -1;
}

// out[0] = lower bound, out[1] = upper bound, out[2] = 1 if a window value lies beyond the table (the reference raises)
__global__ void __launch_bounds__(32)
k_length_bounds(TableView tv, RowMeta meta, int64_t target, int64_t thr, int max_mods, int max_len, BoundMap mp, int64_t* __restrict__ out) {
    __shared__ int32_t s_w[kMaxRows];
    __shared__ int32_t s_ind[kMaxRows];
    __shared__ uint8_t s_mod[kMaxRows];
    for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) {
        s_w[i] = i < tv.R ? tv.weights[i] : 0;
        s_ind[i] = i < tv.R ? meta.ind[i] : 0;
        s_mod[i] = i < tv.R ? meta.is_mod[i] : 0;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int64_t limit = tv.C * 32;
    const int top_row = tv.R - 1;
    const int dl = max_len + 1, du = -1;  // defaults (mass_table.py:365-372)
    const int64_t lo = target - thr, hi = target + thr;
    if (hi >= limit && hi >= 1 && lo <= hi) {
        out[2] = 1;
        return;
    }
    unsigned int fill = 0;

    // one frame per mass on the current path; the one working thread keeps them in shared memory (3.6 KB), not in a
    // per-thread stack (see k_memo_phase_a)
    constexpr int DEPTH = kMaxDepth;
    __shared__ uint32_t f_m[DEPTH + 2];
    __shared__ int f_slot[DEPTH + 2];
    __shared__ uint8_t f_rin[DEPTH + 2], f_cur[DEPTH + 2], f_fill[DEPTH + 2];
    __shared__ int f_all[DEPTH + 2], f_ind[DEPTH + 2];
    __shared__ int8_t f_lo[DEPTH + 2], f_up[DEPTH + 2];  // running min / max
    __shared__ Mask128 f_pend[DEPTH + 2];
    int sp = 0;
    int ret_lo = 0, ret_up = 0;  // value of the node an arrival / a finished frame stands for

    // arrival at node (m, r_in), m > 0: answered from the map (returns false, sets ret_*) or opens a frame
    auto arrive = [&](uint32_t m, int r_in, int all, int ind) -> bool {
        const int slot = bound_slot(mp, m, &fill);
        if (slot < 0 || sp > DEPTH) {
            *mp.overflow = 1;
            ret_lo = dl;
            ret_up = du;
            return false;
        }
        const int top = mp.top[slot];
        if (top >= r_in) {
            ret_lo = mp.lower[(size_t)slot * kMaxRows + r_in];
            ret_up = mp.upper[(size_t)slot * kMaxRows + r_in];
            return false;
        }
        Mask128 pend = mk(ld_nc_u4(tv.H + m));
        mask_keep_le(pend, r_in);
        mask_keep_gt(pend, top);
        f_m[sp] = m; f_slot[sp] = slot; f_rin[sp] = (uint8_t)r_in; f_all[sp] = all; f_ind[sp] = ind;
        f_pend[sp] = pend;
        f_fill[sp] = (uint8_t)(top + 1);
        f_lo[sp] = (int8_t)(top > 0 ? mp.lower[(size_t)slot * kMaxRows + top] : dl);
        f_up[sp] = (int8_t)(top > 0 ? mp.upper[(size_t)slot * kMaxRows + top] : du);
        sp++;
        return true;
    };

    int best_lo = 0, best_up = 0;
    bool first = true;
    for (int64_t v = lo; v <= hi; v++) {
        int b_lo, b_up;
        if (v < 0) {
            b_lo = dl; b_up = du;
        } else if (v == 0) {
            b_lo = 0; b_up = 0;
        } else {
            if (arrive((uint32_t)v, top_row, max_mods, s_ind[top_row])) {
                while (sp > 0) {
                    const int d = sp - 1;
                    const size_t base = (size_t)f_slot[d] * kMaxRows;
                    if (mask_empty(f_pend[d])) {  // every new row of this mass is done: publish, return to the parent
                        for (int r = f_fill[d]; r <= f_rin[d]; r++) {
                            mp.lower[base + r] = f_lo[d];
                            mp.upper[base + r] = f_up[d];
                        }
                        mp.top[f_slot[d]] = f_rin[d];
                        ret_lo = f_lo[d];
                        ret_up = f_up[d];
                        sp--;
                        if (sp > 0) {  // the parent's LEFT candidate through row f_cur: child + 1
                            const int pd = sp - 1;
                            const int cl = ret_lo + 1, cu = ret_up + 1;
                            if (cl < f_lo[pd]) f_lo[pd] = (int8_t)cl;
                            if (cu > f_up[pd]) f_up[pd] = (int8_t)cu;
                            const size_t pb = (size_t)f_slot[pd] * kMaxRows;
                            mp.lower[pb + f_cur[pd]] = f_lo[pd];
                            mp.upper[pb + f_cur[pd]] = f_up[pd];
                        }
                        continue;
                    }
                    const int r = mask_pop_lowest(f_pend[d]);  // LEFT edges fire in ascending row order (UP first)
                    for (int q = f_fill[d]; q < r; q++) {      // rows without an edge inherit the running value
                        mp.lower[base + q] = f_lo[d];
                        mp.upper[base + q] = f_up[d];
                    }
                    f_fill[d] = (uint8_t)(r + 1);
                    f_cur[d] = (uint8_t)r;
                    const int ind_here = (r == f_rin[d]) ? f_ind[d] : s_ind[r];
                    const int mod = s_mod[r];
                    if (mod && !(f_all[d] > 0 && ind_here > 0)) {  // budget-blocked: no candidate from this row
                        mp.lower[base + r] = f_lo[d];
                        mp.upper[base + r] = f_up[d];
                        continue;
                    }
                    const uint32_t m2 = f_m[d] - (uint32_t)s_w[r];
                    bool opened = false;
                    if (m2 == 0u) {
                        ret_lo = 0;
                        ret_up = 0;
                    } else {
                        opened = arrive(m2, r, f_all[d] - mod, ind_here - mod);
                    }
                    if (!opened) {
                        const int cl = ret_lo + 1, cu = ret_up + 1;
                        if (cl < f_lo[d]) f_lo[d] = (int8_t)cl;
                        if (cu > f_up[d]) f_up[d] = (int8_t)cu;
                        mp.lower[base + r] = f_lo[d];
                        mp.upper[base + r] = f_up[d];
                    }
                }
            }
            b_lo = ret_lo;
            b_up = ret_up;
            // a window value whose last-row cell is empty never enters the memo upstream and yields the default;
            // here its mask is empty, the frame closes at once with the defaults: the same value
        }
        if (first) {
            best_lo = b_lo; best_up = b_up; first = false;
        } else {
            if (b_lo < best_lo) best_lo = b_lo;
            if (b_up > best_up) best_up = b_up;
        }
    }
    if (best_lo == dl) best_lo = 1;       // mass_table.py:477-479
    if (best_up == du) best_up = max_len;  // :481-483
    out[0] = best_lo;
    out[1] = best_up;
}

// ---------------- K3 + K2b: the enumeration pass, ONE cooperative launch ----------------
// Level-synchronous expansion.  An ITEM is a partial composition: (remainder m, largest row still allowed rmax,
// rows chosen so far, peak).  Level 0 holds one item per reachable window value (K3, the integer window over the
// last row).  A level is expanded into the next by giving every OPEN item its children {(m - w_r, r) : r <= rmax,
// edge r enabled at m} — one 16-byte row-mask load per item, every thread does the same work — while items that
// need no more loads are carried over unchanged.  In FREE mode (budgets cannot bind) an item stops being open
// early, because the table bits decide everything near the leaves:
//     m == 0            DONE   the path is a composition
//     m <  2*w_min      LEAF   exactly one more nucleotide (the table bit guarantees it exists): 1 composition
//     m <  3*w_min      POPC   every child is DONE or LEAF: popc(children) compositions
// so 1-4 nt ladder differences need two or three levels.  When a level has no open item, records are written by
// ONE THREAD PER COMPOSITION (unranking: j-th child of the item), 8-byte stores that coalesce across the warp.
//
// Every level is count -> exclusive scan -> write over a flat list: each CTA owns a contiguous slice, publishes
// its slice total, all CTAs meet at a grid barrier (cooperative launch: all co-resident) and sum the totals of
// the CTAs before them.  Critical path = (2 barriers + ~3 dependent memory round trips) per level, independent
// of the batch size.  Items stay in depth-first order at every level, so the output is grouped by peak and
// deterministic: (peak, window value, rows descending) — no atomics on the data path except one counter add per
// item for the per-peak totals.
constexpr int kPassThreads = 512;
enum : int { KIND_DONE = 0, KIND_LEAF = 1, KIND_POPC = 2, KIND_OPEN = 3 };

__device__ __forceinline__ Mask128 child_mask(const TableView& tv, const MemoMap& mp, int mode, int64_t p, uint32_t m, int rmax) {
    Mask128 c;
    if (mode == MODE_MEMO) {
        const int slot = memo_find(mp, memo_key(p, m));
        if (slot >= 0) c = mk(mp.alive[slot]);
        else c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0u;
    } else {
        c = mk(ld_nc_u4(tv.H + m));
    }
    mask_keep_le(c, rmax);
    return c;
}
__device__ __forceinline__ int mask_popc(const Mask128& m) { return __popc(m.w[0]) + __popc(m.w[1]) + __popc(m.w[2]) + __popc(m.w[3]); }
// j-th (0-based) set row of a mask with more than j rows
__device__ __forceinline__ int mask_select(const Mask128& m, int j) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int c = __popc(m.w[k]);
        if (j < c) return k * 32 + (int)__fns(m.w[k], 0, j + 1);
        j -= c;
    }
    return 0;
}

// weight -> row in O(1): a collision-free multiplicative hash of the <= 127 row weights, multiplier found by the
// host (sst_cabi.cu, find_leaf_hash), table rebuilt in shared memory by every CTA.  mul == 0: no multiplier was
// found (does not happen for <= 127 keys in 4096 slots within the host's trial budget) -> binary search.
constexpr int kLeafSlots = 4096;
struct LeafHash {
    uint32_t mul;
};
__device__ __forceinline__ uint32_t leaf_slot(uint32_t m, uint32_t mul) { return (m * mul) >> 20; }
__device__ __forceinline__ void leaf_table_init(uint8_t* s_leaf, const int32_t* s_w, int R, LeafHash lh) {
    if (!lh.mul) return;
    for (int i = threadIdx.x; i < kLeafSlots; i += blockDim.x) s_leaf[i] = 0;
    __syncthreads();
    for (int r = 1 + threadIdx.x; r < R; r += blockDim.x) s_leaf[leaf_slot((uint32_t)s_w[r], lh.mul)] = (uint8_t)r;
}
// row q <= rmax with w_q == m, or 0
__device__ __forceinline__ int leaf_row(const uint8_t* s_leaf, const int32_t* s_w, LeafHash lh, uint32_t m, int rmax) {
    if (lh.mul) {
        const int q = s_leaf[leaf_slot(m, lh.mul)];
        return (q && q <= rmax && (uint32_t)s_w[q] == m) ? q : 0;
    }
    int lo = 1, hi = rmax;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)s_w[mid] < m) lo = mid + 1;
        else hi = mid;
    }
    return (uint32_t)s_w[lo] == m ? lo : 0;
}

// One level of items, structure of arrays with capacity `cap` each.  `path` is the record under construction:
// a little-endian number of nw 64-bit words (word k of item i at path[k * cap + i]); appending a row shifts it
// left by one byte, so byte 0 is always the smallest row and the finished number IS the record.
struct ItemBuf {
    uint32_t* m;
    uint32_t* peak;
    uint32_t* meta;              // rmax | depth << 8 | cached POPC count << 16 | budget mode << 24 (bit 31: scratch)
    int32_t* all;                // remaining global modification budget   (only when has_budget)
    int32_t* ind;                // remaining budget of row rmax            (only when has_budget)
    unsigned long long* path;
};
constexpr int kMaxPathWords = kMaxDepth / 8;

constexpr int kMaxLevels = kMaxDepth + 2;

struct PassArgs {
    TableView tv;
    RowMeta meta;
    PeakBatch pk;
    MemoMap mp;
    uint8_t* status;                 // [P]
    ItemBuf buf[2];                  // the frontier, ping-pong
    unsigned long long cap;          // nodes per frontier buffer
    unsigned long long item_limit;   // blow-up guard: more nodes than this in one level -> flags[0]
    uint32_t* cnt;                   // [max(P, cap)] per-entity counts between count and write (children | records << 16)
    uint32_t* chunk_k;               // [max(P, cap) / 32 + grid + 64] children (stage 1: roots) of every 32-entity chunk
    uint32_t* chunk_r;               // the same for the records
    uint4* node_mask;                // [cap] enabled rows of every node, kept from the count phase for the write phase
    int nw;                          // path / record words (W = 8 * nw bytes)
    int has_budget;                  // some peak is in EXACT mode
    unsigned long long* tmp_recs;    // [rec_capacity][nw] records in level order
    uint32_t* tmp_peak;              // [rec_capacity] their peaks
    uint8_t* recs;                   // [rec_capacity][W] records in peak order (the result)
    unsigned long long rec_capacity;
    uint32_t* lvl_cnt;               // [lvl_cap][P+1] records of every peak that were finished at every level
    unsigned long long* lvl_A;       // [lvl_cap][P+1] final position of a level-l record of peak p, minus its index in the level
    int lvl_cap;
    unsigned long long* cta_lvl;     // [lvl_cap][gridDim.x] per-level record totals of every CTA's slice of the peaks
    unsigned long long* peak_off;    // [P+1]
    unsigned long long* cta_tot;     // [2][gridDim.x] children / records of every CTA's slice of the nodes
    // run summary, kept in shared memory by thread 0 of CTA 0 (PassSummary below) and stored to host_out on the way out:
    //   totals [0] roots [1] widest level [2] compositions [3] levels, [8..39] timestamps
    //   flags  [0] item_limit / level limit hit, [1] records overflow (totals[2] = needed), [2] nodes overflow (totals[1] = needed)
    unsigned int* barrier;           // arrival counter of this launch (zero: the previous launch cleared it)
    unsigned int* barrier_next;      // the next launch's counter: cleared by this one (two counters alternate, no memset per run)
    unsigned long long* host_out;    // pinned host memory, [0,40) = totals, [40,42) = flags as ints: written once, by CTA 0, on the way out
    LeafHash leaf;
};

__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// All CTAs of the (cooperative) grid.  `gen` counts the barriers this CTA has passed in this launch.
// The wait polls with a RELAXED load and fences once at the end: an acquire load per poll costs a CCTL.IVALL
// each time, i.e. early CTAs would keep wiping the L1 of the CTAs on the same SM that are still working.
__device__ __forceinline__ void grid_barrier(const PassArgs& a, unsigned int& gen) {
    __syncthreads();
    gen++;
    if (threadIdx.x == 0) {
        __threadfence();  // the CTA's writes (ordered before this by the barrier above) become visible first
        atomicAdd(a.barrier, 1u);
        const unsigned int want = gen * gridDim.x;
        while ((int)(ld_relaxed_u32(a.barrier) - want) < 0) __nanosleep(20);
        __threadfence();  // acquire side: one L1 invalidation, after the wait
    }
    __syncthreads();
}

// timestamp k (diagnostics): CTA 0's clock when it gets here.  Right after a grid barrier that is everybody's
// clock; before one it is only CTA 0's own finish time.  (An atomicMax over all CTAs measured ~1.5 us per stamp.)
struct PassSummary {
    unsigned long long totals[40];
    int flags[4];
};
__device__ __forceinline__ void stamp(PassSummary& sum, int k) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && k < 32) sum.totals[8 + k] = globaltimer_ns();
}

// exclusive scan of one value per thread across the CTA (any whole number of warps up to 32); *total = CTA sum.  Warp
// scans, then warp 0 scans the per-warp totals (three barriers, ~50 instructions per warp).
__device__ __forceinline__ unsigned long long block_scan(unsigned long long x, unsigned long long* total) {
    __shared__ unsigned long long s_w64[32];
    __shared__ unsigned long long s_tot64;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_w64[w] = incl;
    __syncthreads();
    if (w == 0) {
        const unsigned long long v = lane < (int)(blockDim.x >> 5) ? s_w64[lane] : 0ULL;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane < (int)(blockDim.x >> 5)) s_w64[lane] = inc - v;
        if (lane == 31) s_tot64 = inc;
    }
    __syncthreads();
    const unsigned long long r = s_w64[w] + incl - x;
    *total = s_tot64;
    __syncthreads();
    return r;
}

// 32-bit flavour for the per-round scans (a round's counts always fit): ~35 instructions per warp, 3 barriers
__device__ __forceinline__ unsigned int block_scan32(unsigned int x, unsigned int* total) {
    __shared__ unsigned int s_w32[32];
    __shared__ unsigned int s_tot32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_w32[w] = incl;
    __syncthreads();
    if (w == 0) {
        const unsigned int v = lane < (int)(blockDim.x >> 5) ? s_w32[lane] : 0u;
        unsigned int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane < (int)(blockDim.x >> 5)) s_w32[lane] = inc - v;
        if (lane == 31) s_tot32 = inc;
    }
    __syncthreads();
    const unsigned int r = s_w32[w] + incl - x;
    *total = s_tot32;
    __syncthreads();
    return r;
}
template <int N>
__device__ __forceinline__ void block_sum_n(unsigned long long (&v)[N]);
__device__ __forceinline__ unsigned long long block_sum(unsigned long long x);

// four independent 32-bit scans at once (the per-level scans over the peaks share their barriers)
__device__ __forceinline__ void block_scan32x4(unsigned int (&x)[4], unsigned int (&tot)[4]) {
    __shared__ unsigned int s_w4[32][4];
    __shared__ unsigned int s_t4[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int incl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        incl[k] = x[k];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, incl[k], o);
            if (lane >= o) incl[k] += y;
        }
        if (lane == 31) s_w4[w][k] = incl[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned int v = lane < (int)(blockDim.x >> 5) ? s_w4[lane][k] : 0u;
            unsigned int inc = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane < (int)(blockDim.x >> 5)) s_w4[lane][k] = inc - v;
            if (lane == 31) s_t4[k] = inc;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; k++) {
        x[k] = s_w4[w][k] + incl[k] - x[k];  // exclusive prefix
        tot[k] = s_t4[k];
    }
    __syncthreads();
}
// CTA-wide sums of N values per thread: warp shuffles, then warp 0 folds the per-warp partials (three barriers, a
// handful of shared-memory accesses per thread — every thread re-adding all partials cost ~2.5 us per call)
template <int N>
__device__ __forceinline__ void block_sum_n(unsigned long long (&v)[N]) {
    __shared__ unsigned long long s_part[32][N];
    __shared__ unsigned long long s_res[N];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v[k] += __shfl_xor_sync(0xFFFFFFFFu, v[k], o);
        if (lane == 0) s_part[w][k] = v[k];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int k = 0; k < N; k++) {
            unsigned long long t = lane < (int)(blockDim.x >> 5) ? s_part[lane][k] : 0ULL;
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            if (lane == 0) s_res[k] = t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; k++) v[k] = s_res[k];
    __syncthreads();
}

__device__ __forceinline__ unsigned long long block_sum(unsigned long long x) {
    unsigned long long v[1] = {x};
    block_sum_n<1>(v);
    return v[0];
}

// after a grid barrier: sum of the slice totals of the CTAs before this one, and of all CTAs
__device__ __forceinline__ void slice_prefix(const unsigned long long* cta_tot, unsigned long long* base, unsigned long long* all) {
    unsigned long long v[2] = {0ULL, 0ULL};
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const unsigned long long x = __ldcg(cta_tot + b);
        v[1] += x;
        if (b < blockIdx.x) v[0] += x;
    }
    block_sum_n<2>(v);
    *base = v[0];
    *all = v[1];
}
// ---- output-balanced work split ----
// The count phases deal NODES evenly (one load each); the write phases cost per OUTPUT (a child or a record), and
// outputs cluster (a wide 3-nt window is thousands of consecutive open nodes).  So the write phases re-split the
// list by outputs: every 32-node chunk publishes its children and record totals in the count phase, and after
// the grid barrier CTA b takes the chunks whose first output falls in [T*b/G, T*(b+1)/G), T = all outputs.  Finding
// the two boundary chunks is a two-level search: slice totals (G values, scanned in shared memory), then the chunk
// totals of one slice.
struct ChunkRange {
    long long cb, ce;             // this CTA's chunks
    unsigned long long kbase;     // children before chunk cb
    unsigned long long rbase;     // records before chunk cb
    unsigned long long ktotal, rtotal;
};

// number of chunks whose exclusive output prefix is < t, and the children / records before that boundary.  One
// WARP does it (lane-parallel over the chunks of one slice), so both boundaries of a CTA are found at once.
__device__ __forceinline__ void chunk_boundary_warp(const unsigned long long* s_pk, const unsigned long long* s_pr,
                                                    const uint32_t* __restrict__ chunk_k, const uint32_t* __restrict__ chunk_r, long long cps,
                                                    long long nchunks, unsigned long long t, long long* F, unsigned long long* kb,
                                                    unsigned long long* rb) {
    const int lane = threadIdx.x & 31;
    if (t == 0ULL) {
        *F = 0;
        *kb = 0ULL;
        *rb = 0ULL;
        return;
    }
    int lo = 0, hi = (int)gridDim.x;  // last slice that starts before t (slice 0 starts at 0 < t)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_pk[mid] + s_pr[mid] < t) lo = mid;
        else hi = mid;
    }
    const long long s = lo;
    unsigned long long runk = s_pk[s], runr = s_pr[s];
    long long flagged = 0;
    for (long long i0 = 0; i0 < cps; i0 += 32) {
        const long long i = i0 + lane, c = s * cps + i;
        const bool in = i < cps && c < nchunks;
        const unsigned int xk = in ? __ldcg(chunk_k + c) : 0u, xr = in ? __ldcg(chunk_r + c) : 0u;
        unsigned int ik = xk, ir = xr;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int yk = __shfl_up_sync(0xFFFFFFFFu, ik, o), yr = __shfl_up_sync(0xFFFFFFFFu, ir, o);
            if (lane >= o) {
                ik += yk;
                ir += yr;
            }
        }
        const bool f = in && runk + runr + (ik - xk) + (ir - xr) < t;
        const int nf = __popc(__ballot_sync(0xFFFFFFFFu, f));  // a prefix of the lanes: the prefix sums are monotone
        flagged += nf;
        runk += nf ? __shfl_sync(0xFFFFFFFFu, ik, nf - 1) : 0u;
        runr += nf ? __shfl_sync(0xFFFFFFFFu, ir, nf - 1) : 0u;
        if (nf < 32) break;
    }
    *F = s * cps + flagged;
    *kb = runk;
    *rb = runr;
}

__device__ __forceinline__ ChunkRange balanced_range(const unsigned long long* __restrict__ cta_tot, const uint32_t* __restrict__ chunk_k,
                                                     const uint32_t* __restrict__ chunk_r, long long n, long long per,
                                                     unsigned long long* s_pk, unsigned long long* s_pr) {
    __shared__ long long s_F[2];
    __shared__ unsigned long long s_K[2], s_R[2];
    const int G = (int)gridDim.x;  // <= kPassThreads (host)
    unsigned long long ktot, rtot;
    const unsigned long long xk = (int)threadIdx.x < G ? __ldcg(cta_tot + threadIdx.x) : 0ULL;
    const unsigned long long xr = (int)threadIdx.x < G ? __ldcg(cta_tot + G + threadIdx.x) : 0ULL;
    const unsigned long long ek = block_scan(xk, &ktot), er = block_scan(xr, &rtot);
    if ((int)threadIdx.x < G) {
        s_pk[threadIdx.x] = ek;
        s_pr[threadIdx.x] = er;
    }
    if (threadIdx.x == 0) {
        s_pk[G] = ktot;
        s_pr[G] = rtot;
    }
    __syncthreads();
    const long long cps = per / 32, nchunks = (n + 31) / 32;
    const int w = threadIdx.x >> 5;
    if (w < 2) {  // warp 0: where this CTA starts, warp 1: where it ends
        long long F;
        unsigned long long kb, rb;
        chunk_boundary_warp(s_pk, s_pr, chunk_k, chunk_r, cps, nchunks, (ktot + rtot) * (blockIdx.x + w) / G, &F, &kb, &rb);
        if ((threadIdx.x & 31) == 0) {
            s_F[w] = F;
            s_K[w] = kb;
            s_R[w] = rb;
        }
    }
    __syncthreads();
    ChunkRange r;
    r.ktotal = ktot;
    r.rtotal = rtot;
    r.cb = s_F[0];
    r.kbase = s_K[0];
    r.rbase = s_R[0];
    r.ce = (int)blockIdx.x == G - 1 ? nchunks : s_F[1];  // trailing chunks without outputs go to the last CTA
    __syncthreads();
    return r;
}

// entities of a list of n are dealt to CTAs in contiguous slices of `per` (a multiple of 32)
__device__ __forceinline__ long long slice_size(long long n) {
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    return (per + 31) & ~31LL;
}

// reachable window values of one peak: calls f(word_index, reach_bits) for every word of the window, in
// ascending order; the loads go out four at a time (the words are independent, the loop would otherwise pay one
// DRAM round trip per word)
template <typename F>
__device__ __forceinline__ void for_window_words(const uint64_t* __restrict__ last, int64_t a, int64_t b, F f) {
    if (a > b) return;
    const int64_t w0 = a >> 5, w1 = b >> 5;
    for (int64_t wd = w0; wd <= w1; wd += 4) {
        uint64_t x[4];
#pragma unroll
        for (int q = 0; q < 4; q++) x[q] = (wd + q <= w1) ? __ldg(last + wd + q) : 0ULL;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (wd + q > w1) break;
            uint64_t y = (x[q] | (x[q] >> 1)) & kBit0Mask;
            if (wd + q == w0) y &= (1ULL << (2 * (31 - (int)(a & 31)) + 1)) - 1ULL;
            if (wd + q == w1) y &= ~0ULL << (2 * (31 - (int)(b & 31)));
            f(wd + q, y);
        }
    }
}

struct RowTables {  // per-CTA shared copies
    const int32_t* w;
    const int32_t* ind;
    const uint8_t* mod;
    const uint8_t* leaf;
    LeafHash lh;
    uint32_t wmin;
};

__device__ __forceinline__ int item_kind(int mode, uint32_t m, uint32_t wmin) {
    if (m == 0u) return KIND_DONE;
    if (mode != MODE_FREE) return KIND_OPEN;
    if (m < 2u * wmin) return KIND_LEAF;
    if (m < 3u * wmin) return KIND_POPC;
    return KIND_OPEN;
}

// enabled LEFT edges of an open item.  EXACT mode filters by the budgets the item carries (reference
// mass_explanation.py:165-172: a modification row needs all > 0 and ind > 0, where ind is the remaining budget
// of the row the level was entered by, or IND[r] for any other row).
__device__ __forceinline__ Mask128 open_children(const TableView& tv, const MemoMap& mp, const RowTables& rt, int mode, uint32_t p, uint32_t m, int rmax,
                                                 int all, int ind) {
    Mask128 c = child_mask(tv, mp, mode, p, m, rmax);
    if (mode == MODE_EXACT) {
        Mask128 scan = c;
        while (!mask_empty(scan)) {
            const int r = mask_pop_lowest(scan);
            if (!rt.mod[r]) continue;
            const int ind_here = (r == rmax) ? ind : rt.ind[r];
            if (!(all > 0 && ind_here > 0)) c.w[r >> 5] &= ~(1u << (r & 31));
        }
    }
    return c;
}

// append row r to a path held in registers
__device__ __forceinline__ void path_append(unsigned long long* pw, int nw, int r) {
    for (int k = nw - 1; k > 0; k--) pw[k] = (pw[k] << 8) | (pw[k - 1] >> 56);
    pw[0] = (pw[0] << 8) | (unsigned long long)r;
}

// NW = path words known at compile time (1 or 2: the path stays in registers), 0 = run-time a.nw (deep
// compositions of test alphabets; path arrays in local memory)
template <int NW>
__global__ void __launch_bounds__(kPassThreads, 2)
k_explain_pass(const PassArgs a) {
    constexpr int kPW = NW ? NW : kMaxPathWords;
    __shared__ int32_t s_w[kMaxRows];
    __shared__ int32_t s_ind[kMaxRows];
    __shared__ uint8_t s_mod[kMaxRows];
    __shared__ uint8_t s_leaf[kLeafSlots];
    __shared__ int s_nheavy;
    __shared__ unsigned long long s_pk[kPassThreads + 1], s_pr[kPassThreads + 1];
    __shared__ unsigned int s_cbase_k[kPassThreads], s_cbase_r[kPassThreads];
    __shared__ unsigned int s_ttot_k, s_ttot_r;
    __shared__ unsigned long long s_lbase[kMaxLevels + 1], s_lrun[kMaxLevels];
    __shared__ int s_heavy_p[kPassThreads];
    __shared__ unsigned long long s_heavy_off[kPassThreads];
    __shared__ uint32_t s_heavy_wa[kPassThreads], s_heavy_wb[kPassThreads];
    const TableView& tv = a.tv;
    for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) {
        s_w[i] = i < tv.R ? tv.weights[i] : 0;
        s_ind[i] = i < tv.R ? a.meta.ind[i] : 0;
        s_mod[i] = i < tv.R ? a.meta.is_mod[i] : 0;
    }
    __shared__ PassSummary s_sum;  // only thread 0 of CTA 0 touches it
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int k = 0; k < 40; k++) s_sum.totals[k] = 0ULL;
        for (int k = 0; k < 4; k++) s_sum.flags[k] = 0;
        for (int k = 0; k < 8; k++) a.barrier_next[k] = 0u;  // nobody uses the other set during this launch (the depth-first pass keeps a cursor and a flag there too)
        s_sum.totals[8] = globaltimer_ns();
    }
    // the run summary goes straight to pinned host memory when the pass ends (no copy operation after the launch)
    auto publish = [&]() {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            for (int k = 0; k < 40; k++) a.host_out[k] = s_sum.totals[k];
            int* hf = reinterpret_cast<int*>(a.host_out + 40);
            for (int k = 0; k < 4; k++) hf[k] = s_sum.flags[k];
        }
    };
    __syncthreads();
    leaf_table_init(s_leaf, s_w, tv.R, a.leaf);
    __syncthreads();
    const RowTables rt{s_w, s_ind, s_mod, s_leaf, a.leaf, tv.R > 1 ? (uint32_t)s_w[1] : 0u};
    const int lane = threadIdx.x & 31;
    const int nw = NW ? NW : a.nw;
    unsigned int gen = 0;
    int ts = 1;  // next timestamp slot
    const int64_t P = a.pk.P;
    const int64_t limit = tv.C * 32;
    const int top_row = tv.R - 1;
    const uint64_t* last = tv.tbl + (int64_t)top_row * tv.C;
    unsigned long long base, n_roots = 0, n_items = 0, n_comps = 0;
    (void)n_comps;
    int cur = 0, level = 0;

    // ---- stage 1 (K3): peaks -> level-0 items, one per reachable window value ----
    {
        const long long per = slice_size(P), first = (long long)blockIdx.x * per;
        unsigned long long mine = 0;
        for (long long li = threadIdx.x; li < per; li += blockDim.x) {
            const long long p = first + li;
            if (p >= P) break;
            const int64_t lo = a.pk.target[p] - a.pk.thr[p], hi = a.pk.target[p] + a.pk.thr[p];
            uint8_t st = 0;
            if (lo <= 0 && 0 <= hi) st |= ST_ZERO_IN_WINDOW;
            if (lo <= hi && hi >= limit) st |= ST_OUT_OF_TABLE;
            a.status[p] = st;
            unsigned int n = 0;
            for_window_words(last, lo < 1 ? 1 : lo, hi < limit - 1 ? hi : limit - 1, [&](int64_t, uint64_t x) { n += __popcll(x); });
            a.cnt[p] = n;
            mine += n;
        }
        const unsigned long long tot = block_sum(mine);
        if (threadIdx.x == 0) a.cta_tot[blockIdx.x] = tot;
        stamp(s_sum, ts++);
        grid_barrier(a, gen);
        slice_prefix(a.cta_tot, &base, &n_roots);
        n_items = n_roots;
        if (n_roots > a.cap) {  // the same answer in every CTA: the host grows the buffers and runs the pass again
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                s_sum.flags[2] = 1;
                s_sum.totals[0] = n_roots;
                s_sum.totals[1] = n_roots;
            }
            publish();
            return;
        }
        const ItemBuf& out = a.buf[0];
        const uint32_t meta0 = (uint32_t)top_row;
        auto put_root = [&](unsigned long long o, uint32_t v, uint32_t p) {
            out.m[o] = v;
            out.peak[o] = p;
            out.meta[o] = meta0 | ((uint32_t)a.pk.mode[p] << 24);
            for (int k = 0; k < nw; k++) out.path[(unsigned long long)k * a.cap + o] = 0ULL;
            if (a.has_budget) {
                out.all[o] = a.pk.max_mods[p];
                out.ind[o] = s_ind[top_row];
            }
        };
        for (long long l0 = 0; l0 < per; l0 += blockDim.x) {  // uniform trip count: block_scan has barriers
            const long long p = first + l0 + threadIdx.x;
            const bool ok = l0 + threadIdx.x < per && p < P;
            const unsigned int n = ok ? a.cnt[p] : 0u;
            unsigned int round_tot;
            unsigned long long off = base + block_scan32(n, &round_tot);
            base += round_tot;
            int64_t wa = 1, wb = 0;
            if (ok) {
                const int64_t lo = a.pk.target[p] - a.pk.thr[p], hi = a.pk.target[p] + a.pk.thr[p];
                wa = lo < 1 ? 1 : lo;
                wb = hi < limit - 1 ? hi : limit - 1;
            }
            // A peak with few roots writes them itself.  One with many (a wide 2-3 nt window) goes on a CTA-wide
            // list that all warps then drain together: lane l takes window word l, l+32, ..., a warp scan places the
            // roots.  (Heavy peaks cluster; left to their own warps they serialise the whole grid at the barrier.)
            const bool heavy = ok && n > 6;
            if (ok && n && !heavy) {
                for_window_words(last, wa, wb, [&](int64_t wd, uint64_t x) {
                    while (x) {  // ascending mass = descending bit position
                        const int pos = 63 - __clzll((long long)x);
                        x &= ~(1ULL << pos);
                        put_root(off++, (uint32_t)(wd * 32 + (31 - (pos >> 1))), (uint32_t)p);
                    }
                });
            }
            if (threadIdx.x == 0) s_nheavy = 0;
            __syncthreads();
            if (heavy) {
                const int q = atomicAdd(&s_nheavy, 1);
                s_heavy_p[q] = (int)(p - first);
                s_heavy_off[q] = off;
                s_heavy_wa[q] = (uint32_t)wa;  // window bounds (masses fit 32 bits: the table is < 2^31 masses wide)
                s_heavy_wb[q] = (uint32_t)wb;
            }
            __syncthreads();
            const int nheavy = s_nheavy;
            // four heavy peaks per warp at a time: an 8-lane group takes one peak, lane s of the group its window
            // words s, s+8, ..., an 8-wide scan places the roots
            const int grp = threadIdx.x >> 3, sub = lane & 7;
            for (int q0 = 0; q0 < nheavy; q0 += kPassThreads / 8) {  // uniform trip count
                const int q = q0 + grp;
                const bool live = q < nheavy;
                const uint32_t hp = live ? (uint32_t)(first + s_heavy_p[q]) : 0u;
                unsigned long long run = live ? s_heavy_off[q] : 0ULL;
                const int64_t sa = live ? (int64_t)s_heavy_wa[q] : 1, sb = live ? (int64_t)s_heavy_wb[q] : 0;
                const int64_t w0 = sa >> 5, w1 = sb >> 5;
                const int64_t w1_all = __reduce_max_sync(0xFFFFFFFFu, (int)(live ? (w1 - w0) : -1));  // longest window in the warp, in words
                for (int64_t k0 = 0; k0 <= w1_all; k0 += 8) {
                    const int64_t wd = w0 + k0 + sub;
                    uint64_t x = 0;
                    if (live && wd <= w1) {
                        x = __ldg(last + wd);
                        x = (x | (x >> 1)) & kBit0Mask;
                        if (wd == w0) x &= (1ULL << (2 * (31 - (int)(sa & 31)) + 1)) - 1ULL;
                        if (wd == w1) x &= ~0ULL << (2 * (31 - (int)(sb & 31)));
                    }
                    const unsigned c = __popcll(x);
                    unsigned incl = c;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o, 8);
                        if (sub >= o) incl += y;
                    }
                    unsigned long long o2 = run + (incl - c);
                    while (x) {
                        const int pos = 63 - __clzll((long long)x);
                        x &= ~(1ULL << pos);
                        put_root(o2++, (uint32_t)(wd * 32 + (31 - (pos >> 1))), hp);
                    }
                    run += __shfl_sync(0xFFFFFFFFu, incl, 7, 8);
                }
            }
            __syncthreads();
        }
        stamp(s_sum, ts++);
        grid_barrier(a, gen);
    }

    // ---- levels: every node of the frontier either FINISHES (its records go to the level-ordered buffer) or is
    //      EXPANDED (its children are the next frontier).  Nothing is carried along. ----
    unsigned long long rec_run = 0;  // records of the levels before this one
    bool dry = false;                // the record buffers are too small: keep counting, write no records
    unsigned long long widest = n_roots;
    for (;;) {
        const ItemBuf& in = a.buf[cur];
        const long long n = (long long)n_items, per = slice_size(n), first = (long long)blockIdx.x * per;
        if (level >= a.lvl_cap || level >= kMaxLevels) {  // cannot happen: the host sizes lvl_cap from the depth bound
            if (blockIdx.x == 0 && threadIdx.x == 0) s_sum.flags[0] = 1;
            publish();
            return;
        }
        if (threadIdx.x == 0) s_lbase[level] = rec_run;
        // this level's per-peak record counters start at zero (they are added to after the barrier)
        {
            uint32_t* lc = a.lvl_cnt + (size_t)level * (size_t)(P + 1);
            for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q <= P; q += (long long)gridDim.x * blockDim.x) lc[q] = 0u;
        }
        // count: one row-mask load per node that needs one
        unsigned long long mine_k = 0, mine_r = 0;
        for (long long li = threadIdx.x; li < per; li += blockDim.x) {  // whole warps enter together: per % 32 == 0
            const long long i = first + li;
            unsigned int k = 0, r = 0;
            if (i < n) {
                const uint32_t m = __ldcg(in.m + i);
                const uint32_t meta = __ldcg(in.meta + i);
                const int rmax = meta & 0xFF;
                const int mode = (meta >> 24) & 3;
                const uint32_t p = mode == MODE_MEMO ? __ldcg(in.peak + i) : 0u;  // only the memo key needs the peak
                const int kind = item_kind(mode, m, rt.wmin);
                Mask128 c;
                c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0u;
                if (kind == KIND_DONE || kind == KIND_LEAF) {
                    r = 1;
                } else if (kind == KIND_POPC) {
                    c = child_mask(tv, a.mp, mode, p, m, rmax);
                    r = (unsigned)mask_popc(c);
                } else {
                    const int all = a.has_budget ? __ldcg(in.all + i) : 0, ind = a.has_budget ? __ldcg(in.ind + i) : 0;
                    c = open_children(tv, a.mp, rt, mode, p, m, rmax, all, ind);
                    k = (unsigned)mask_popc(c);
                }
                a.cnt[i] = k | (r << 16);
                a.node_mask[i] = make_uint4(c.w[0], c.w[1], c.w[2], c.w[3]);  // the write phase reads it back coalesced
                mine_k += k;
                mine_r += r;
            }
            unsigned int wk = k, wr = r;  // the chunk's totals, for the output-balanced split of the write phase
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                wk += __shfl_xor_sync(0xFFFFFFFFu, wk, o);
                wr += __shfl_xor_sync(0xFFFFFFFFu, wr, o);
            }
            if (lane == 0) {
                a.chunk_k[i >> 5] = wk;
                a.chunk_r[i >> 5] = wr;
            }
        }
        unsigned long long tsum[2] = {mine_k, mine_r};
        block_sum_n<2>(tsum);
        if (threadIdx.x < 2) a.cta_tot[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = tsum[threadIdx.x];
        stamp(s_sum, ts++);
        grid_barrier(a, gen);
        stamp(s_sum, ts++);
        const ChunkRange cr = balanced_range(a.cta_tot, a.chunk_k, a.chunk_r, n, per, s_pk, s_pr);
        const unsigned long long n_next = cr.ktotal, n_rec = cr.rtotal;
        if (n_next > a.cap || n_next > a.item_limit) {  // the same answer in every CTA
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                s_sum.flags[n_next > a.item_limit ? 0 : 2] = 1;
                s_sum.totals[0] = n_roots;
                s_sum.totals[1] = n_next;
                s_sum.totals[3] = (unsigned long long)level;
            }
            publish();
            return;
        }
        if (rec_run + n_rec > a.rec_capacity) dry = true;
        if (n_next > widest) widest = n_next;
        const ItemBuf& out = a.buf[cur ^ 1];
        uint32_t* lc = a.lvl_cnt + (size_t)level * (size_t)(P + 1);
        unsigned long long tile_k = 0, tile_r = 0;
        for (long long tile0 = cr.cb; tile0 < cr.ce; tile0 += blockDim.x) {  // up to blockDim chunks per tile, one warp per chunk
            {
                const long long c = tile0 + threadIdx.x;
                const unsigned int xk = c < cr.ce ? __ldcg(a.chunk_k + c) : 0u, xr = c < cr.ce ? __ldcg(a.chunk_r + c) : 0u;
                unsigned int tk, tr;
                s_cbase_k[threadIdx.x] = block_scan32(xk, &tk);
                s_cbase_r[threadIdx.x] = block_scan32(xr, &tr);
                if (threadIdx.x == 0) {
                    s_ttot_k = tk;
                    s_ttot_r = tr;
                }
            }
            __syncthreads();
            for (long long wq = threadIdx.x >> 5; tile0 + wq < cr.ce && wq < (long long)blockDim.x; wq += kPassThreads / 32) {
                const long long i = (tile0 + wq) * 32 + lane;
                const bool ok = i < n;
                const unsigned int packed = ok ? __ldcg(a.cnt + i) : 0u;
                const unsigned int k = packed & 0xFFFFu, r = packed >> 16;
                unsigned int kin = k, rin = r;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int yk = __shfl_up_sync(0xFFFFFFFFu, kin, o), yr = __shfl_up_sync(0xFFFFFFFFu, rin, o);
                    if (lane >= o) {
                        kin += yk;
                        rin += yr;
                    }
                }
                const unsigned long long off = cr.kbase + tile_k + s_cbase_k[wq] + (kin - k);            // first child
                const unsigned long long roff = rec_run + cr.rbase + tile_r + s_cbase_r[wq] + (rin - r);  // first record
                uint32_t m = 0, p = 0xFFFFFFFFu, meta = 0;
                int all = 0, ind = 0, kind = KIND_DONE;
                unsigned long long pw[kPW];
#pragma unroll
                for (int q = 0; q < kPW; q++) pw[q] = 0ULL;
                Mask128 c;
                c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0u;
                if (ok) {  // everything a node needs comes in one round of coalesced loads (no dependent gather)
                    m = __ldcg(in.m + i);
                    p = __ldcg(in.peak + i);
                    meta = __ldcg(in.meta + i) & 0x7FFFFFFFu;
                    for (int q = 0; q < nw; q++) pw[q] = __ldcg(in.path + (unsigned long long)q * a.cap + i);
                    if (a.has_budget) {
                        all = __ldcg(in.all + i);
                        ind = __ldcg(in.ind + i);
                    }
                    c = mk(__ldcg(a.node_mask + i));
                    kind = item_kind((meta >> 24) & 3, m, rt.wmin);
                    if (!(k | r)) p = 0xFFFFFFFFu;  // a dead end: takes no part in the per-peak totals
                }

                // ---- records of the nodes that finish here ----
                {
                    // per-peak totals of this level: lanes with the same peak fold their counts first
                    const unsigned int grp = __match_any_sync(0xFFFFFFFFu, p);
                    const unsigned int sum = __reduce_add_sync(grp, r);
                    if (sum && lane == __ffs(grp) - 1) atomicAdd(lc + p, sum);
                }
                if (!dry) {
                    // the j-th record of a node: its path, plus (LEAF) the one row that closes it, or (POPC) the j-th
                    // enabled row r2 and, if something is left, the row that closes that
                    auto put = [&](unsigned long long at, uint32_t im, int rmax, int knd, int r2, const unsigned long long* path, uint32_t pk) {
                        unsigned long long w[kPW];
                        for (int q = 0; q < nw; q++) w[q] = path[q];
                        if (knd == KIND_LEAF) {
                            path_append(w, nw, leaf_row(s_leaf, s_w, a.leaf, im, rmax));
                        } else if (knd == KIND_POPC) {
                            const uint32_t m3 = im - (uint32_t)s_w[r2];
                            path_append(w, nw, r2);
                            if (m3) path_append(w, nw, leaf_row(s_leaf, s_w, a.leaf, m3, r2));
                        }
                        for (int q = 0; q < nw; q++) a.tmp_recs[at * (unsigned long long)nw + q] = w[q];
                        a.tmp_peak[at] = pk;
                    };
                    const bool heavy = r > 4;
                    if (r && !heavy) {  // the owner writes its few records itself
                        Mask128 left = c;
                        for (unsigned int j = 0; j < r; j++) put(roff + j, m, meta & 0xFF, kind, kind == KIND_POPC ? mask_pop_lowest(left) : 0, pw, p);
                    }
                    for (unsigned hm = __ballot_sync(0xFFFFFFFFu, heavy); hm; hm &= hm - 1) {  // many records: the warp shares them
                        const int src = __ffs(hm) - 1;
                        const uint32_t sm = __shfl_sync(0xFFFFFFFFu, m, src), smeta = __shfl_sync(0xFFFFFFFFu, meta, src);
                        const uint32_t sp = __shfl_sync(0xFFFFFFFFu, p, src);
                        const unsigned int sr = __shfl_sync(0xFFFFFFFFu, r, src);
                        const unsigned long long so = __shfl_sync(0xFFFFFFFFu, roff, src);
                        Mask128 scm;
#pragma unroll
                        for (int q = 0; q < 4; q++) scm.w[q] = __shfl_sync(0xFFFFFFFFu, c.w[q], src);
                        unsigned long long spw[kPW];
                        for (int q = 0; q < nw; q++) spw[q] = __shfl_sync(0xFFFFFFFFu, pw[q], src);
                        for (unsigned int j = lane; j < sr; j += 32) put(so + j, sm, smeta & 0xFF, KIND_POPC, mask_select(scm, (int)j), spw, sp);
                    }
                }

                // ---- children of the nodes that are expanded ----
                // Output-centric placement: the warp's open nodes produce T consecutive children; lane l writes children
                // l, l+32, ... — finds the parent lane by a shuffle search over the per-lane child offsets, pulls the
                // parent through shuffles and picks its j-th enabled row.  Work is per CHILD and every store coalesces.
                const unsigned long long off0 = __shfl_sync(0xFFFFFFFFu, off, 0);
                const unsigned wp = (unsigned)(off - off0);
                const unsigned T = __shfl_sync(0xFFFFFFFFu, wp + k, 31);
                for (unsigned o0 = 0; o0 < T; o0 += 32) {
                    const unsigned o = o0 + lane;
                    int src = 0;
#pragma unroll
                    for (int step = 16; step; step >>= 1) {
                        const int cand = src + step;
                        const unsigned v = __shfl_sync(0xFFFFFFFFu, wp, cand & 31);
                        if (cand < 32 && v <= o) src = cand;
                    }
                    const int j = (int)(o - __shfl_sync(0xFFFFFFFFu, wp, src));
                    const uint32_t sm = __shfl_sync(0xFFFFFFFFu, m, src), sp = __shfl_sync(0xFFFFFFFFu, p, src);
                    const uint32_t smeta = __shfl_sync(0xFFFFFFFFu, meta, src);
                    Mask128 sc;
#pragma unroll
                    for (int q = 0; q < 4; q++) sc.w[q] = __shfl_sync(0xFFFFFFFFu, c.w[q], src);
                    unsigned long long cw[kPW];
                    for (int q = 0; q < nw; q++) cw[q] = __shfl_sync(0xFFFFFFFFu, pw[q], src);
                    int sall = 0, sind = 0;
                    if (a.has_budget) {
                        sall = __shfl_sync(0xFFFFFFFFu, all, src);
                        sind = __shfl_sync(0xFFFFFFFFu, ind, src);
                    }
                    if (o < T) {
                        const unsigned long long o2 = off0 + o;
                        const int rr = mask_select(sc, j);
                        const int srmax = smeta & 0xFF, sdepth = (smeta >> 8) & 0xFF;
                        path_append(cw, nw, rr);
                        out.m[o2] = sm - (uint32_t)s_w[rr];
                        out.peak[o2] = sp;
                        out.meta[o2] = (uint32_t)rr | ((uint32_t)(sdepth + 1) << 8) | (smeta & 0x03000000u);
                        for (int q = 0; q < nw; q++) out.path[(unsigned long long)q * a.cap + o2] = cw[q];
                        if (a.has_budget) {
                            const int mod = s_mod[rr];
                            out.all[o2] = sall - mod;
                            out.ind[o2] = ((rr == srmax) ? sind : s_ind[rr]) - mod;
                        }
                    }
                }
            }
            tile_k += s_ttot_k;
            tile_r += s_ttot_r;
            __syncthreads();
        }
        rec_run += n_rec;
        level++;
        stamp(s_sum, ts++);
        grid_barrier(a, gen);
        cur ^= 1;
        n_items = n_next;
        if (n_next == 0) break;
    }
    n_comps = rec_run;
    if (threadIdx.x == 0) s_lbase[level] = rec_run;
    const int n_levels = level;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (dry) s_sum.flags[1] = 1;
        s_sum.totals[0] = n_roots;
        s_sum.totals[1] = widest;
        s_sum.totals[2] = n_comps;
        s_sum.totals[3] = (unsigned long long)n_levels;
    }
    if (dry) {  // the host enlarges the record buffers and runs the pass again
        publish();
        return;
    }

    // ---- where the records go: peak_off[p] = records of the peaks before p; a level-l record of peak p that is
    //      the x-th record of its level goes to A[l][p] + x, A[l][p] = peak_off[p] + (records of p finished at
    //      earlier levels) - (records of level l that belong to peaks before p) ----
    const long long pper = slice_size(P + 1), pfirst = (long long)blockIdx.x * pper;
    const size_t pstride = (size_t)(P + 1);
    for (int l4 = 0; l4 < n_levels; l4 += 4) {  // four levels share one reduction
        unsigned long long mine[4] = {0ULL, 0ULL, 0ULL, 0ULL};
        for (long long li = threadIdx.x; li < pper; li += blockDim.x) {
            const long long q = pfirst + li;
            if (q <= P) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (l4 + k < n_levels) mine[k] += __ldcg(a.lvl_cnt + (size_t)(l4 + k) * pstride + q);
            }
        }
        block_sum_n<4>(mine);
        if (threadIdx.x < 4 && l4 + (int)threadIdx.x < n_levels) a.cta_lvl[(size_t)(l4 + threadIdx.x) * gridDim.x + blockIdx.x] = mine[threadIdx.x];
    }
    stamp(s_sum, ts++);
    grid_barrier(a, gen);
    for (int l4 = 0; l4 < n_levels; l4 += 4) {  // records of every level in the slices before this CTA's
        unsigned long long mine[4] = {0ULL, 0ULL, 0ULL, 0ULL};
        for (unsigned b2 = threadIdx.x; b2 < blockIdx.x; b2 += blockDim.x) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (l4 + k < n_levels) mine[k] += __ldcg(a.cta_lvl + (size_t)(l4 + k) * gridDim.x + b2);
        }
        block_sum_n<4>(mine);
        if (threadIdx.x < 4 && l4 + (int)threadIdx.x < n_levels) s_lrun[l4 + threadIdx.x] = mine[threadIdx.x];
    }
    __syncthreads();
    for (long long l0 = 0; l0 < pper; l0 += blockDim.x) {
        const long long q = pfirst + l0 + threadIdx.x;
        const bool ok = l0 + threadIdx.x < pper && q <= P;
        unsigned long long poff = 0;
        for (int l4 = 0; l4 < n_levels; l4 += 4) {  // first pass: start of peak q inside every level
            unsigned int x[4], tot[4];
            unsigned long long before[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                x[k] = (ok && l4 + k < n_levels) ? __ldcg(a.lvl_cnt + (size_t)(l4 + k) * pstride + q) : 0u;
                before[k] = l4 + k < n_levels ? s_lrun[l4 + k] : 0ULL;  // read before the scan's barriers: advanced after them
            }
            block_scan32x4(x, tot);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (l4 + k < n_levels) {
                    const unsigned long long start = before[k] + x[k];
                    if (ok) a.lvl_A[(size_t)(l4 + k) * pstride + q] = start;
                    poff += start;
                }
            }
            if (threadIdx.x < 4 && l4 + (int)threadIdx.x < n_levels) s_lrun[l4 + threadIdx.x] += tot[threadIdx.x];
            __syncthreads();
        }
        if (ok) {
            a.peak_off[q] = poff;
            unsigned long long cum = 0;
            for (int l = 0; l < n_levels; l++) {  // second pass: the placement table
                const size_t at = (size_t)l * pstride + q;
                const unsigned long long start = a.lvl_A[at];
                a.lvl_A[at] = poff + cum - start;
                cum += __ldcg(a.lvl_cnt + at);
            }
        }
    }
    if (n_levels == 0)
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q <= P; q += (long long)gridDim.x * blockDim.x) a.peak_off[q] = 0ULL;
    stamp(s_sum, ts++);
    grid_barrier(a, gen);

    // ---- permute: level order -> peak order (four records per thread in flight) ----
    {
        unsigned long long* recs64 = reinterpret_cast<unsigned long long*>(a.recs);
        const unsigned long long j0 = n_comps * blockIdx.x / gridDim.x, j1 = n_comps * (blockIdx.x + 1) / gridDim.x;
        for (unsigned long long jb = j0; jb < j1; jb += 4ULL * blockDim.x) {
            unsigned long long j[4], dst[4];
            uint32_t pk[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                j[u] = jb + (unsigned long long)u * blockDim.x + threadIdx.x;
                pk[u] = j[u] < j1 ? __ldcg(a.tmp_peak + j[u]) : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                dst[u] = 0;
                if (j[u] < j1) {
                    int l = 0;
                    while (l + 1 < n_levels && j[u] >= s_lbase[l + 1]) l++;
                    dst[u] = __ldcg(a.lvl_A + (size_t)l * pstride + pk[u]) + (j[u] - s_lbase[l]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (j[u] < j1)
                    for (int q = 0; q < nw; q++) recs64[dst[u] * (unsigned long long)nw + q] = __ldcg(a.tmp_recs + j[u] * (unsigned long long)nw + q);
        }
    }
    stamp(s_sum, ts++);
    publish();
}

}  // namespace sst
