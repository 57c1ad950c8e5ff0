// K3 + K2b, count -> scan -> fill over partial compositions in PEAK ORDER: one cooperative launch, ONE grid barrier.
//
// Replaces explain_mass_with_table (reference mass_explanation.py:92-203; the window loop :192-201 and the inner
// backtrack :118-188) for batches whose compositions are at most kDfsDepth nucleotides long — ladder differences and
// singletons, i.e. everything prediction.py / skeleton_building.py ask for.  Deeper batches (test alphabets with tiny
// weights) and batches with single subtrees too large for one CTA take the level-synchronous pass (k_explain_pass,
// sst_explain.cuh), which spreads one huge subtree over the whole machine.
//
// An ITEM is a partial composition (remainder m, largest row still allowed, rows so far, peak); a window ROOT — a
// reachable value of a peak's integer window, last-row bit set — is the item with no rows yet.  The compositions below
// an item are the paths to remainder 0 of the DAG {(m, rows <= rmax) -> (m - w_r, rows <= r) : r <= rmax, edge r
// enabled at m}; its edges come from one 16-byte load of the mass-major row mask H[m].  In FREE mode the table bits
// decide everything near the leaves (item_kind: DONE / LEAF / POPC are closed forms with at most that one load).
//
// Every CTA owns a contiguous stretch of the batch — contiguous in (peak, window value) order and equal in ESTIMATED
// work (peak_cost(): compositions expected in the window, from a coarse count made by the host; no table access), so a
// 5-nt ladder gap with 10^4 compositions is shared by several CTAs, each taking a sub-range of its window, while a
// CTA of 1-nt differences takes a few hundred whole peaks.  A CTA works through its stretch entirely on its own (no
// grid-wide step until the counts are complete), every step over ALL its peaks / items at once, so the number of
// dependent steps does not grow with the batch:
//   count phase
//     roots:   thread per PEAK: window words -> number of roots; scan; the roots are written to a stretch of the
//              item pool in peak order (a peak with many roots is emitted by its whole warp).
//     split:   up to kMaxSplit rounds.  Every OPEN item is replaced by its children, closed items are carried over;
//              thread per item, so every row-mask load of a round is in flight at once, and the list stays in
//              peak order (depth-first order inside a peak).  Offsets come from warp scans over 32-item chunks plus
//              one scan over the chunk totals: two CTA barriers per round however long the list is.
//     count:   compositions below every item of the final list: closed forms, or — for what is left of deeper
//              subtrees — a depth-first walk with an explicit stack of at most kDfsDepth frames.
//   grid barrier: CTA totals -> record base of every CTA's slice (the only grid-wide dependency).
//   fill phase: item i of the CTA's final list starts at (records of the CTAs before) + (exclusive scan of the
//     counts), because items are in peak order; peak_off[p] is the start of the peak's first item.  Thread per item,
//     8-byte record stores at the final positions; the masks are read back from the pool, not loaded again.
// Inside a thread the items it handles are processed kU at a time, loads first, so that a CTA with a long list (one
// peak with thousands of compositions) is bound by throughput, not by kU-times as many dependent round trips.
//
// The level-synchronous pass needs 2 grid barriers per level + placement + permute (12 for a batch of 1-3 nt
// differences); here the only data ever written are items, counts and records.
#pragma once
#include <type_traits>

#include "sst_explain.cuh"

namespace sst {

constexpr int kDfsThreads = 512;             // threads per CTA, two CTAs per SM: the throughput shape (large batches)
constexpr int kDfsThreadsWide = 1024;        // one CTA per SM: half as many stretches, each with twice the threads — a heavy
                                             // stretch needs half as many steps and the estimate errors average out over
                                             // twice as many peaks; measured 82 against 88 us on the C4 batch, 0.54 against
                                             // 0.53 ms on the batch tiled 16x, so the host picks by batch size
constexpr int64_t kDfsWideMaxPeaks = 400000; // batches up to this many peaks take the wide shape
constexpr int kDfsDepth = 16;               // frames of the per-thread stack = longest composition this pass accepts
constexpr int kMaxSplit = 2;                // split rounds before the remaining subtrees are walked depth-first (measured on C4 /
                                            // the batch tiled 16x: six rounds 82.2 us / 0.532 ms, two 79.5 / 0.508, one 88 us —
                                            // a round costs every CTA ~5 us, a thread walks what two rounds leave of a subtree)
constexpr int kU = 2;                       // items a thread has in flight
constexpr unsigned kDfsNodeCap = 1u << 15;  // edge expansions per item before the pass hands the batch to the level-synchronous one

struct ItemPool {
    uint32_t* m;
    uint32_t* peak;
    uint32_t* meta;               // rmax | rows so far << 8 | budget mode << 24
    unsigned long long* path;     // [cap][nw] the record under construction (byte 0 = smallest row so far)
    int32_t* all;                 // remaining budgets (only when some peak is in EXACT mode)
    int32_t* ind;
    uint4* mask;                  // enabled children of an item (kept from a count pass for the pass that writes)
    uint32_t* cnt;                // compositions below an item of a final list
    uint32_t* off;                // outputs (children / records) of the items before this one in its 32-item chunk
    uint32_t* cpre;               // [cap / 32] outputs of the chunks before this one in its list
    unsigned long long cap;       // items (a multiple of 32)
};

// Speculative launch (sst_explain_submit_f64): the batch was staged but its summary has not been looked at by the host.
// The pass reads it first and leaves the batch alone when a peak is not FREE, a window reaches further than the record
// width holds, an input was not finite, or the batch is too heavy for this pass (summary layout: k_stage_f64).
struct SpecGuard {
    const unsigned long long* summary;  // null: no speculation
    unsigned long long max_hi;          // largest window end the launch can take
    unsigned long long max_cost;        // largest summed peak_cost()
};

struct DfsArgs {
    TableView tv;
    RowMeta meta;
    PeakBatch pk;
    MemoMap mp;
    uint8_t* status;                  // [P]
    uint32_t* peak_first;             // [P] first item of every peak in its CTA's list (kept up to date through the split rounds)
    uint32_t* peak_cpre;              // [P / 32 + gridDim.x + 1] roots of the 32-peak chunks before this one in its CTA's stretch
    ItemPool pool;
    const uint32_t* peak_cost;        // [P] estimated work of every peak (peak_cost(), written when the batch is staged)
    const unsigned long long* blk_cost;  // [ceil(P / kCostBlock)] their sums
    unsigned long long* cta_info;     // [gridDim.x][3] start, length and compositions of every CTA's final list
    unsigned long long* recs;         // [rec_capacity][nw] records, final order
    unsigned long long rec_capacity;
    // SPLIT records (pipelined submissions, 8-byte records only): the first four nucleotides of record i as rec_lo[i], the
    // k-th further one as rec_hi[k * rec_capacity + i], k < rec_hi_planes — a batch of <= 5-nt differences crosses the bus
    // with 5 bytes per composition instead of 8.  rec_lo == null: whole 8-byte words go to recs.
    uint32_t* rec_lo;
    uint8_t* rec_hi;
    int rec_hi_planes;
    unsigned long long* peak_off;     // [P+1]
    uint32_t* peak_off32;             // [P+1] the same as uint32 (may be null; only meaningful below 2^32 compositions)
    unsigned long long* cta_tot;      // [2][gridDim.x] compositions / roots of every CTA's slice
    unsigned int* sync;               // this launch's words: [0] barrier arrivals, [2..3] pool cursor (u64), [4] fallback flag
    unsigned int* sync_next;          // the next launch's words: cleared by this one
    unsigned long long* host_out;     // pinned + mapped run summary (layout of PassSummary)
    LeafHash leaf;
    unsigned long long* cta_ns;       // diagnostics (may be null): [gridDim.x][8] %globaltimer of every CTA at its phase boundaries
    SpecGuard guard;
};

__device__ __forceinline__ void cta_stamp(const DfsArgs& a, int k) {
    if (a.cta_ns && threadIdx.x == 0) a.cta_ns[(size_t)blockIdx.x * 8 + k] = globaltimer_ns();
}

__device__ __forceinline__ void dfs_barrier(const DfsArgs& a, unsigned int& gen) {
    __syncthreads();
    gen++;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.sync, 1u);
        const unsigned int want = gen * gridDim.x;
        while ((int)(ld_relaxed_u32(a.sync) - want) < 0) __nanosleep(20);
        __threadfence();
    }
    __syncthreads();
}

// enabled children of an OPEN item, given the raw row mask of its remainder (FREE / EXACT; MEMO looks the mass up in
// the first-visit map instead).  EXACT filters by the budgets the item carries (mass_explanation.py:165-172).
template <bool BUDGET>
__device__ __forceinline__ Mask128 children_of(const DfsArgs& a, const RowTables& rt, int mode, uint32_t pk, uint32_t m, int rmax, int all,
                                               int ind, uint4 raw) {
    Mask128 c;
    if (mode == MODE_MEMO) {
        const int slot = memo_find(a.mp, memo_key(pk, m));
        if (slot >= 0) c = mk(a.mp.alive[slot]);
        else c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0u;
    } else {
        c = mk(raw);
    }
    mask_keep_le(c, rmax);
    if (BUDGET && mode == MODE_EXACT) {
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

template <int NW>
__device__ __forceinline__ void store_record(const DfsArgs& a, unsigned long long at, const unsigned long long* w) {
    if (NW == 1 && a.rec_lo) {
        a.rec_lo[at] = (uint32_t)w[0];
        uint32_t hi = (uint32_t)(w[0] >> 32);
        for (int k = 0; k < a.rec_hi_planes; k++, hi >>= 8) a.rec_hi[(unsigned long long)k * a.rec_capacity + at] = (uint8_t)hi;
    } else {
#pragma unroll
        for (int q = 0; q < NW; q++) a.recs[at * NW + q] = w[q];
    }
}

// One item, depth-first: what is left of a subtree after the split rounds.  EMIT = false: returns the number of
// compositions below it.  EMIT = true: also stores them at recs[at...] (children in ascending row order, like the
// reference's UP-before-LEFT walk).  `capped` is set when the walk gives up.
template <int NW, bool BUDGET, bool EMIT>
__device__ __noinline__ unsigned int dfs_item(const DfsArgs& a, const RowTables& rt, uint32_t m, int rmax, int all, int ind,
                                              unsigned long long* path, uint32_t p, int mode, unsigned long long at, bool* capped) {
    uint32_t f_m[kDfsDepth];
    Mask128 f_pend[kDfsDepth];
    uint8_t f_rmax[kDfsDepth];
    int f_all[BUDGET ? kDfsDepth : 1], f_ind[BUDGET ? kDfsDepth : 1];
    unsigned long long f_path[EMIT ? kDfsDepth : 1][NW];
    int sp = 0;
    unsigned int total = 0, work = 0;

    auto emit = [&](const unsigned long long* w) {
        if (EMIT) {
            store_record<NW>(a, at, w);
            at++;
        }
        total++;
    };

    for (;;) {
        // ---- enter node (m, rows <= rmax) with prefix `path`
        const int kind = item_kind(mode, m, rt.wmin);
        if (kind == KIND_DONE) {
            emit(path);
        } else if (kind == KIND_LEAF) {
            if (EMIT) path_append(path, NW, leaf_row(rt.leaf, rt.w, rt.lh, m, rmax));
            emit(path);
        } else if (kind == KIND_POPC) {
            Mask128 c = child_mask(a.tv, a.mp, mode, p, m, rmax);
            if (EMIT) {
                while (!mask_empty(c)) {
                    const int r2 = mask_pop_lowest(c);
                    unsigned long long w[NW];
#pragma unroll
                    for (int q = 0; q < NW; q++) w[q] = path[q];
                    const uint32_t m3 = m - (uint32_t)rt.w[r2];
                    path_append(w, NW, r2);
                    if (m3) path_append(w, NW, leaf_row(rt.leaf, rt.w, rt.lh, m3, r2));
                    emit(w);
                }
            } else {
                total += (unsigned)mask_popc(c);
            }
        } else {
            const Mask128 c = open_children(a.tv, a.mp, rt, mode, p, m, rmax, all, ind);
            if (!mask_empty(c) && sp < kDfsDepth) {
                f_m[sp] = m;
                f_pend[sp] = c;
                f_rmax[sp] = (uint8_t)rmax;
                if (BUDGET) {
                    f_all[sp] = all;
                    f_ind[sp] = ind;
                }
                if (EMIT) {
#pragma unroll
                    for (int q = 0; q < NW; q++) f_path[sp][q] = path[q];
                }
                sp++;
            }
        }
        // ---- next edge of the deepest open frame
        while (sp > 0 && mask_empty(f_pend[sp - 1])) sp--;
        if (sp == 0) break;
        if (++work > kDfsNodeCap) {
            *capped = true;
            break;
        }
        const int d = sp - 1;
        const int r = mask_pop_lowest(f_pend[d]);
        m = f_m[d] - (uint32_t)rt.w[r];
        if (BUDGET) {
            const int mod = rt.mod[r];
            all = f_all[d] - mod;
            ind = ((r == f_rmax[d]) ? f_ind[d] : rt.ind[r]) - mod;
        }
        rmax = r;
        if (EMIT) {
#pragma unroll
            for (int q = 0; q < NW; q++) path[q] = f_path[d][q];
            path_append(path, NW, r);
        }
    }
    return total;
}

template <int NW, bool BUDGET, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_explain_dfs(const DfsArgs a) {
    constexpr int kDfsThreads = THREADS;  // (shadows the namespace constant: everything below is per instance)
    __shared__ int32_t s_w[kMaxRows];
    __shared__ int32_t s_ind[kMaxRows];
    __shared__ uint8_t s_mod[kMaxRows];
    __shared__ uint8_t s_leaf[kLeafSlots];
    __shared__ unsigned long long s_base;
    __shared__ PassSummary s_sum;  // only thread 0 of CTA 0 touches it
    const TableView& tv = a.tv;
    const ItemPool& pool = a.pool;
    for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) {
        s_w[i] = i < tv.R ? tv.weights[i] : 0;
        s_ind[i] = i < tv.R ? a.meta.ind[i] : 0;
        s_mod[i] = i < tv.R ? a.meta.is_mod[i] : 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (int k = 0; k < 40; k++) s_sum.totals[k] = 0ULL;
        for (int k = 0; k < 4; k++) s_sum.flags[k] = 0;
        for (int k = 0; k < 8; k++) a.sync_next[k] = 0u;  // nobody uses the other set during this launch
        s_sum.totals[8] = globaltimer_ns();
    }
    auto publish = [&]() {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            for (int k = 0; k < 40; k++) a.host_out[k] = s_sum.totals[k];
            int* hf = reinterpret_cast<int*>(a.host_out + 40);
            for (int k = 0; k < 4; k++) hf[k] = s_sum.flags[k];
        }
    };
    __syncthreads();
    if (a.guard.summary) {  // the same answer in every CTA: nobody reaches the grid barrier
        const unsigned long long* g = a.guard.summary;
        if (__ldcg(g + 1) > a.guard.max_hi || __ldcg(g + 2) || __ldcg(g + 3) || __ldcg(g + 4) || __ldcg(g + 5) > a.guard.max_cost) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                s_sum.flags[3] = 1;
                a.host_out[0] = 0ULL;
                a.host_out[2] = 0ULL;
            }
            publish();
            return;
        }
    }
    leaf_table_init(s_leaf, s_w, tv.R, a.leaf);
    __syncthreads();
    const RowTables rt{s_w, s_ind, s_mod, s_leaf, a.leaf, tv.R > 1 ? (uint32_t)s_w[1] : 0u};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kDfsThreads / 32;
    const int64_t P = a.pk.P;
    const int64_t limit = tv.C * 32;
    const int top_row = tv.R - 1;
    const uint64_t* last = tv.tbl + (int64_t)top_row * tv.C;
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(a.sync + 2);
    unsigned int* fallback = a.sync + 4;
    unsigned int gen = 0;
    int ts = 1;

    // ---- this CTA's stretch of the batch: costs [E*b/G, E*(b+1)/G) of the running sum over (peak, window value)
    __shared__ unsigned long long s_S[kDfsThreads + 1];
    __shared__ long long s_loc_p[2];
    __shared__ unsigned long long s_loc_r[2], s_loc_e[2];
    {
        const long long n_blk = (P + kCostBlock - 1) / kCostBlock, q = (n_blk + kDfsThreads - 1) / kDfsThreads;
        unsigned long long mine = 0;
        for (long long j = (long long)threadIdx.x * q; j < ((long long)threadIdx.x + 1) * q && j < n_blk; j++) mine += __ldg(a.blk_cost + j);
        unsigned long long E;
        s_S[threadIdx.x] = block_scan(mine, &E);
        if (threadIdx.x == 0) s_S[kDfsThreads] = E;
        __syncthreads();
        if (warp < 2) {  // warp 0: where this CTA starts, warp 1: where it ends
            const unsigned long long X = (unsigned long long)(((unsigned __int128)E * (blockIdx.x + warp)) / gridDim.x);
            long long lp = P;
            unsigned long long lr = 0, le = 1;
            if (X < E) {
                int t0 = 0, t1 = kDfsThreads;  // last run of blocks that starts at or before X
                while (t1 - t0 > 1) {
                    const int mid = (t0 + t1) >> 1;
                    if (s_S[mid] <= X) t0 = mid;
                    else t1 = mid;
                }
                // first element of arr[j0, j1) whose inclusive running sum (from `run`) exceeds X
                auto find = [&](auto ld, long long j0, long long j1, unsigned long long& run) -> long long {
                    for (long long g = j0; g < j1; g += 32) {
                        const unsigned long long v = g + lane < j1 ? ld(g + lane) : 0ULL;
                        unsigned long long incl = v;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                            if (lane >= o) incl += y;
                        }
                        const unsigned hit = __ballot_sync(0xFFFFFFFFu, g + lane < j1 && run + incl > X);
                        if (hit) {
                            const int l = __ffs(hit) - 1;
                            run += __shfl_sync(0xFFFFFFFFu, incl - v, l);
                            return g + l;
                        }
                        run += __shfl_sync(0xFFFFFFFFu, incl, 31);
                    }
                    return j1;
                };
                unsigned long long run = s_S[t0];
                const long long hi_blk = ((long long)t0 + 1) * q < n_blk ? ((long long)t0 + 1) * q : n_blk;
                const long long blk = find([&](long long j) { return __ldg(a.blk_cost + j); }, (long long)t0 * q, hi_blk, run);
                const long long p1 = (blk + 1) * kCostBlock < P ? (blk + 1) * kCostBlock : P;
                lp = find([&](long long j) { return (unsigned long long)__ldg(a.peak_cost + j); }, blk * kCostBlock, p1, run);
                if (lp < P) {
                    lr = X - run;
                    le = __ldg(a.peak_cost + lp);
                } else {
                    lp = P;  // (rounding at the very end of the batch)
                }
            }
            if (lane == 0) {
                s_loc_p[warp] = lp;
                s_loc_r[warp] = lr;
                s_loc_e[warp] = le;
            }
        }
        __syncthreads();
    }
    // peaks pb .. p_last; the first one only from fraction rb / eb of its window on, the last one (if it is pe) only up
    // to fraction re / ee
    const long long pb = s_loc_p[0], pe = s_loc_p[1];
    const unsigned long long rb = s_loc_r[0], eb = s_loc_e[0], re = s_loc_r[1], ee = s_loc_e[1];
    const long long p_last = re > 0 ? pe : pe - 1;
    const long long n_mine = p_last >= pb ? p_last - pb + 1 : 0;  // peaks of this CTA (the first and the last possibly in part)
    uint32_t* const my_cpre = a.peak_cpre + (pb >> 5) + blockIdx.x;  // chunk j = peaks pb + 32 j ..

    // A stretch of the pool for a list of n items (whole 32-item chunks, so that chunk numbers are global).  The first
    // half of the pool is divided among the CTAs up front — a bump pointer every thread keeps for itself, no atomic and
    // no barrier — the second half is handed out by one atomic add per list on a global cursor once a CTA's share is
    // used up.  The order of the stretches in the pool is irrelevant.
    const unsigned long long share = ((pool.cap / 2) / gridDim.x) & ~31ULL;
    unsigned long long bump = share * blockIdx.x;
    const unsigned long long bump_end = bump + share;
    auto reserve = [&](unsigned long long n) -> unsigned long long {
        const unsigned long long need = (n + 31ULL) & ~31ULL;
        if (bump + need <= bump_end) {
            const unsigned long long at = bump;
            bump += need;
            return at;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = need ? (pool.cap / 2) + atomicAdd(cursor, need) : 0ULL;
        __syncthreads();
        return s_base;
    };
    auto put_item = [&](unsigned long long o, uint32_t m, uint32_t p, uint32_t meta, const unsigned long long* path, int all, int ind) {
        pool.m[o] = m;
        pool.peak[o] = p;
        pool.meta[o] = meta;
#pragma unroll
        for (int q = 0; q < NW; q++) pool.path[o * NW + q] = path[q];
        if (BUDGET) {
            pool.all[o] = all;
            pool.ind[o] = ind;
        }
    };
    // first output (child / record) of item i of the list at A with n items and `total` outputs
    auto offset_of = [&](unsigned long long A, unsigned int i, unsigned int n, unsigned int total) -> unsigned int {
        return i < n ? pool.cpre[(A + i) >> 5] + pool.off[A + i] : total;
    };

    // Outputs of every item of the list at A, and their offsets.  FINAL = false (a split round): an OPEN item has one
    // output per enabled child, any other item one (itself).  FINAL = true: outputs are compositions.  A warp takes kU
    // 32-item chunks at a time — loads first — scans each and stores the chunk totals; one CTA-wide scan over the chunk
    // totals follows.  Returns the number of outputs of the list.
    bool capped = false;
    auto count_pass = [&](auto final_tag, unsigned long long A, unsigned int NA, bool& any_open, unsigned long long& comps) -> unsigned int {
        constexpr bool FINAL = decltype(final_tag)::value;
        const unsigned int n_chunks = (NA + 31) >> 5;
        for (unsigned int c0 = warp * kU; c0 < n_chunks; c0 += kWarps * kU) {
            uint32_t m[kU], meta[kU];
            uint4 raw[kU];
            bool valid[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const unsigned int i = ((c0 + u) << 5) + lane;
                valid[u] = i < NA;
                m[u] = valid[u] ? pool.m[A + i] : 0u;
                meta[u] = valid[u] ? pool.meta[A + i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const int mode = (meta[u] >> 24) & 3;
                const int kind = item_kind(mode, m[u], rt.wmin);
                const bool need = valid[u] && mode != MODE_MEMO && (kind == KIND_OPEN || (FINAL && kind == KIND_POPC));
                raw[u] = need ? ld_nc_u4(tv.H + m[u]) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const unsigned int c = c0 + u, i = (c << 5) + lane;
                const int mode = (meta[u] >> 24) & 3, rmax = meta[u] & 0xFF;
                const int kind = item_kind(mode, m[u], rt.wmin);
                unsigned int k = 0;
                if (valid[u]) {
                    if (kind == KIND_OPEN && !FINAL) {
                        const uint32_t pk = mode == MODE_MEMO ? pool.peak[A + i] : 0u;
                        const Mask128 ch = children_of<BUDGET>(a, rt, mode, pk, m[u], rmax, BUDGET ? pool.all[A + i] : 0, BUDGET ? pool.ind[A + i] : 0, raw[u]);
                        pool.mask[A + i] = make_uint4(ch.w[0], ch.w[1], ch.w[2], ch.w[3]);
                        k = (unsigned)mask_popc(ch);
                        any_open = true;
                    } else if (kind == KIND_OPEN) {  // what the split rounds left of a deep subtree
                        unsigned long long nopath[NW] = {};
                        k = dfs_item<NW, BUDGET, false>(a, rt, m[u], rmax, BUDGET ? pool.all[A + i] : 0, BUDGET ? pool.ind[A + i] : 0, nopath,
                                                        pool.peak[A + i], mode, 0ULL, &capped);
                    } else if (kind == KIND_POPC && FINAL) {
                        Mask128 ch = mk(raw[u]);
                        mask_keep_le(ch, rmax);
                        pool.mask[A + i] = make_uint4(ch.w[0], ch.w[1], ch.w[2], ch.w[3]);
                        k = (unsigned)mask_popc(ch);
                    } else {
                        k = 1;
                    }
                    if (FINAL) {
                        pool.cnt[A + i] = k;
                        comps += k;
                    }
                }
                unsigned int incl = k;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += y;
                }
                if (valid[u]) pool.off[A + i] = incl - k;
                if (lane == 31 && c < n_chunks) pool.cpre[(A >> 5) + c] = incl;
            }
        }
        __syncthreads();
        unsigned int run = 0;  // exclusive scan of the chunk totals, kDfsThreads at a time
        for (unsigned int c0 = 0; c0 < n_chunks; c0 += kDfsThreads) {
            const unsigned int c = c0 + threadIdx.x;
            const unsigned int x = c < n_chunks ? pool.cpre[(A >> 5) + c] : 0u;
            unsigned int tot;
            const unsigned int ex = block_scan32(x, &tot);
            if (c < n_chunks) pool.cpre[(A >> 5) + c] = run + ex;
            run += tot;
        }
        __syncthreads();
        return run;
    };

    // ---------------- count phase ----------------
    cta_stamp(a, 0);
    unsigned long long my_comps = 0, my_roots = 0;
    bool overflow = false;
    // this CTA's part of the (clamped) window of peak p: [wa, wb], empty when wa > wb
    auto window_of = [&](long long p, uint32_t& wa, uint32_t& wb, bool owner) {
        wa = 1;
        wb = 0;
        const int64_t tg = a.pk.target[p], th = a.pk.thr[p];
        const int64_t lo = tg - th, hi = tg + th;
        if (owner) {  // the peak starts in this CTA's stretch: it writes status and peak_off
            uint8_t st = 0;
            if (lo <= 0 && 0 <= hi) st |= ST_ZERO_IN_WINDOW;
            if (lo <= hi && hi >= limit) st |= ST_OUT_OF_TABLE;
            a.status[p] = st;
        }
        int64_t ca = lo < 1 ? 1 : lo, cb = hi < limit - 1 ? hi : limit - 1;
        if (ca <= cb) {  // a peak shared with the neighbouring CTAs
            const unsigned long long W = (unsigned long long)(cb - ca + 1);
            const int64_t full_a = ca;
            if (p == pb && rb > 0) ca = full_a + (int64_t)((rb * W) / eb);
            if (p == pe) cb = full_a + (int64_t)((re * W) / ee) - 1;
        }
        if (ca <= cb) {
            wa = (uint32_t)ca;
            wb = (uint32_t)cb;
        }
    };
    // ---- roots of every peak; a warp handles 32 consecutive peaks at a time and scans their counts
    const long long n_pchunks = (n_mine + 31) >> 5;
    for (long long c = warp; c < n_pchunks; c += kWarps) {
        const long long j = (c << 5) + lane, p = pb + j;
        unsigned int n = 0;
        if (j < n_mine) {
            uint32_t wa, wb;
            window_of(p, wa, wb, p > pb || rb == 0);
            if (wa <= wb) for_window_words(last, (int64_t)wa, (int64_t)wb, [&](int64_t, uint64_t x) { n += __popcll(x); });
        }
        my_roots += n;
        unsigned int incl = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += y;
        }
        // (a first peak that started in the previous CTA's stretch keeps no entry: its part here starts at item 0, and
        //  peak_first[p] belongs to the CTA the peak starts in)
        if (j < n_mine && (p > pb || rb == 0)) a.peak_first[p] = incl - n;
        if (lane == 31) my_cpre[c] = incl;
    }
    __syncthreads();
    unsigned int NA = 0;
    for (long long c0 = 0; c0 < n_pchunks; c0 += kDfsThreads) {  // exclusive scan of the chunk totals
        const long long c = c0 + threadIdx.x;
        const unsigned int x = c < n_pchunks ? my_cpre[c] : 0u;
        unsigned int tot;
        const unsigned int ex = block_scan32(x, &tot);
        if (c < n_pchunks) my_cpre[c] = NA + ex;
        NA += tot;
    }
    __syncthreads();
    cta_stamp(a, 1);
    unsigned long long A = reserve(NA);
    if (A + NA > pool.cap) overflow = true;
    // ---- window roots in peak order; a peak with many roots is emitted by its whole warp (lane l takes window word l, l+32, ...)
    if (!overflow) {
        const unsigned long long zero[NW] = {};
        const int ind0 = s_ind[top_row];
        for (long long c = warp; c < n_pchunks; c += kWarps) {
            const long long j = (c << 5) + lane, p = pb + j;
            const bool ok = j < n_mine;
            uint32_t wa = 1, wb = 0, pf = 0, n = 0, meta0 = (uint32_t)top_row;
            int all0 = 0;
            if (ok) {
                window_of(p, wa, wb, false);
                if (p > pb || rb == 0) {
                    pf = my_cpre[c] + a.peak_first[p];
                    a.peak_first[p] = pf;
                }
                // roots of the peak = first item of the next peak - first item of this one (the last lane: chunk total)
                meta0 |= (uint32_t)a.pk.mode[p] << 24;
                if (BUDGET) all0 = a.pk.max_mods[p];
            }
            {
                const uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, pf, 1);
                const uint32_t chunk_end = c + 1 < n_pchunks ? my_cpre[c + 1] : NA;
                n = ok ? ((lane == 31 || j + 1 >= n_mine) ? chunk_end : nxt) - pf : 0u;
            }
            const bool heavy = n > 8;
            if (n && !heavy) {
                unsigned long long o = A + pf;
                for_window_words(last, (int64_t)wa, (int64_t)wb, [&](int64_t wd, uint64_t x) {
                    while (x) {  // ascending mass = descending bit position
                        const int pos = 63 - __clzll((long long)x);
                        x &= ~(1ULL << pos);
                        put_item(o++, (uint32_t)(wd * 32 + (31 - (pos >> 1))), (uint32_t)p, meta0, zero, all0, ind0);
                    }
                });
            }
            for (unsigned hm = __ballot_sync(0xFFFFFFFFu, heavy); hm; hm &= hm - 1) {
                const int src = __ffs(hm) - 1;
                const uint32_t sa = __shfl_sync(0xFFFFFFFFu, wa, src), sb = __shfl_sync(0xFFFFFFFFu, wb, src);
                const uint32_t sp = __shfl_sync(0xFFFFFFFFu, (uint32_t)p, src), smeta = __shfl_sync(0xFFFFFFFFu, meta0, src);
                const int sall = __shfl_sync(0xFFFFFFFFu, all0, src);
                unsigned long long run = A + __shfl_sync(0xFFFFFFFFu, pf, src);
                const uint32_t w0 = sa >> 5, w1 = sb >> 5;
                for (uint32_t k0 = w0; k0 <= w1; k0 += 32) {
                    const uint32_t wd = k0 + lane;
                    uint64_t x = 0;
                    if (wd <= w1) {
                        x = __ldg(last + wd);
                        x = (x | (x >> 1)) & kBit0Mask;
                        if (wd == w0) x &= (1ULL << (2 * (31 - (int)(sa & 31)) + 1)) - 1ULL;
                        if (wd == w1) x &= ~0ULL << (2 * (31 - (int)(sb & 31)));
                    }
                    const unsigned cc = __popcll(x);
                    unsigned incl = cc;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    unsigned long long o2 = run + (incl - cc);
                    while (x) {
                        const int pos = 63 - __clzll((long long)x);
                        x &= ~(1ULL << pos);
                        put_item(o2++, wd * 32 + (31 - (pos >> 1)), sp, smeta, zero, sall, ind0);
                    }
                    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
            }
        }
    }
    __syncthreads();  // the roots are visible to the whole CTA
    cta_stamp(a, 2);

    // ---- split rounds
    for (int round = 0; round < kMaxSplit && !overflow; round++) {
        bool any_open = false;
        unsigned long long unused = 0;
        const unsigned int NB = count_pass(std::false_type{}, A, NA, any_open, unused);
        if (!__syncthreads_or(any_open)) break;
        const unsigned long long B = reserve(NB);
        if (B + NB > pool.cap) {
            overflow = true;
            break;
        }
        for (unsigned int i0 = threadIdx.x; i0 < NA; i0 += kDfsThreads * kU) {
            uint32_t m[kU], meta[kU], pk[kU], o[kU];
            unsigned long long path[kU][NW];
            int all[kU], ind[kU];
            uint4 ch[kU];
            bool valid[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {  // everything an item needs comes in one round of independent loads
                const unsigned int i = i0 + u * kDfsThreads;
                valid[u] = i < NA;
                const unsigned long long at = A + (valid[u] ? i : 0u);
                m[u] = pool.m[at];
                meta[u] = pool.meta[at];
                pk[u] = pool.peak[at];
                o[u] = pool.cpre[at >> 5] + pool.off[at];
                ch[u] = pool.mask[at];
#pragma unroll
                for (int q = 0; q < NW; q++) path[u][q] = pool.path[at * NW + q];
                all[u] = BUDGET ? pool.all[at] : 0;
                ind[u] = BUDGET ? pool.ind[at] : 0;
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                if (!valid[u]) continue;
                unsigned long long dst = B + o[u];
                if (item_kind((meta[u] >> 24) & 3, m[u], rt.wmin) == KIND_OPEN) {
                    const int rmax = meta[u] & 0xFF;
                    Mask128 c = mk(ch[u]);
                    while (!mask_empty(c)) {
                        const int r = mask_pop_lowest(c);
                        unsigned long long w[NW];
#pragma unroll
                        for (int q = 0; q < NW; q++) w[q] = path[u][q];
                        path_append(w, NW, r);
                        const int mod = BUDGET ? s_mod[r] : 0;
                        put_item(dst++, m[u] - (uint32_t)s_w[r], pk[u], (uint32_t)r | ((meta[u] & 0xFFFFFF00u) + 0x100u), w, all[u] - mod,
                                 ((r == rmax) ? ind[u] : s_ind[r]) - mod);
                    }
                } else {
                    put_item(dst, m[u], pk[u], meta[u], path[u], all[u], ind[u]);
                }
            }
        }
        for (long long j = (rb > 0 ? 1 : 0) + threadIdx.x; j < n_mine; j += kDfsThreads)  // (not a first peak that the previous CTA owns)
            a.peak_first[pb + j] = offset_of(A, a.peak_first[pb + j], NA, NB);
        A = B;
        NA = NB;
        __syncthreads();
    }
    cta_stamp(a, 6);

    // ---- compositions below every item of the final list, and where its records start in the CTA's stretch
    unsigned int my_recs = 0;
    if (!overflow) {
        bool dummy = false;
        my_recs = count_pass(std::true_type{}, A, NA, dummy, my_comps);
    }
    if (threadIdx.x == 0) {
        unsigned long long* ci = a.cta_info + (size_t)blockIdx.x * 3;
        ci[0] = A;
        ci[1] = NA;
        ci[2] = my_recs;
        if (a.cta_ns) a.cta_ns[(size_t)blockIdx.x * 8 + 7] = ((unsigned long long)n_mine << 44) | ((unsigned long long)NA << 22) | my_recs;  // diagnostics
    }
    if (capped || overflow) atomicExch(fallback, overflow ? 2u : 1u);
    {
        unsigned long long v[2] = {my_comps, my_roots};
        block_sum_n<2>(v);
        if (threadIdx.x < 2) a.cta_tot[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = v[threadIdx.x];
    }
    stamp(s_sum, ts++);
    cta_stamp(a, 3);
    dfs_barrier(a, gen);
    cta_stamp(a, 4);
    stamp(s_sum, ts++);

    // ---------------- record base of this CTA's slice ----------------
    unsigned long long rec_base, n_comps, n_roots;
    {
        unsigned long long v[3] = {0ULL, 0ULL, 0ULL};
        for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
            const unsigned long long x = __ldcg(a.cta_tot + b);
            v[1] += x;
            if (b < blockIdx.x) v[0] += x;
            v[2] += __ldcg(a.cta_tot + gridDim.x + b);
        }
        block_sum_n<3>(v);
        rec_base = v[0];
        n_comps = v[1];
        n_roots = v[2];
    }
    const unsigned int fb = __ldcg(fallback);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        s_sum.totals[0] = n_roots;
        s_sum.totals[1] = __ldcg(cursor);                      // pool items taken from the shared half
        s_sum.totals[2] = n_comps;
        s_sum.totals[3] = 1ULL;
        if (fb == 1u) s_sum.flags[3] = 1;                      // a subtree too large for one thread: level-synchronous pass
        if (fb == 2u) s_sum.flags[2] = 1;                      // pool too small: the host grows it, runs again
        if (n_comps > a.rec_capacity) s_sum.flags[1] = 1;      // records do not fit: the host grows the buffer, runs again
    }
    if (fb || n_comps > a.rec_capacity) {
        publish();
        return;
    }

    // ---------------- fill phase ----------------
    {
        const unsigned long long run = rec_base;
        const unsigned long long* ci = a.cta_info + (size_t)blockIdx.x * 3;
        const unsigned long long A = ci[0];
        const unsigned int NA = (unsigned int)ci[1], my_recs = (unsigned int)ci[2];
        // peak role: a peak starts where its first item starts (items and records are both in peak order)
        for (long long j = threadIdx.x; j < n_mine; j += kDfsThreads) {
            const long long p = pb + j;
            if (p > pb || rb == 0) {
                const unsigned long long at = run + offset_of(A, a.peak_first[p], NA, my_recs);
                a.peak_off[p] = at;
                if (a.peak_off32) a.peak_off32[p] = (uint32_t)at;
            }
        }
        // item role
        for (unsigned int i0 = threadIdx.x; i0 < NA; i0 += kDfsThreads * kU) {
            uint32_t m[kU], meta[kU], cnt[kU], o[kU];
            unsigned long long path[kU][NW];
            uint4 ch[kU];
#pragma unroll
            for (int u = 0; u < kU; u++) {
                const unsigned int i = i0 + u * kDfsThreads;
                const bool valid = i < NA;
                const unsigned long long at = A + (valid ? i : 0u);
                cnt[u] = valid ? pool.cnt[at] : 0u;
                m[u] = pool.m[at];
                meta[u] = pool.meta[at];
                o[u] = pool.cpre[at >> 5] + pool.off[at];
                ch[u] = pool.mask[at];
#pragma unroll
                for (int q = 0; q < NW; q++) path[u][q] = pool.path[at * NW + q];
            }
#pragma unroll
            for (int u = 0; u < kU; u++) {
                if (!cnt[u]) continue;
                const int mode = (meta[u] >> 24) & 3, rmax = meta[u] & 0xFF;
                const int kind = item_kind(mode, m[u], rt.wmin);
                unsigned long long dst = run + o[u];
                if (kind == KIND_OPEN) {
                    const unsigned long long at = A + i0 + u * kDfsThreads;
                    bool dummy = false;
                    dfs_item<NW, BUDGET, true>(a, rt, m[u], rmax, BUDGET ? pool.all[at] : 0, BUDGET ? pool.ind[at] : 0, path[u], pool.peak[at], mode, dst,
                                               &dummy);
                } else if (kind == KIND_POPC) {
                    Mask128 c = mk(ch[u]);
                    while (!mask_empty(c)) {
                        const int r2 = mask_pop_lowest(c);
                        unsigned long long w[NW];
#pragma unroll
                        for (int q = 0; q < NW; q++) w[q] = path[u][q];
                        const uint32_t m3 = m[u] - (uint32_t)s_w[r2];
                        path_append(w, NW, r2);
                        if (m3) path_append(w, NW, leaf_row(s_leaf, s_w, a.leaf, m3, r2));
                        store_record<NW>(a, dst, w);
                        dst++;
                    }
                } else {
                    if (kind == KIND_LEAF) path_append(path[u], NW, leaf_row(s_leaf, s_w, a.leaf, m[u], rmax));
                    store_record<NW>(a, dst, path[u]);
                }
            }
        }
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        a.peak_off[P] = n_comps;
        if (a.peak_off32) a.peak_off32[P] = (uint32_t)n_comps;
    }
    __syncthreads();
    cta_stamp(a, 5);
    stamp(s_sum, ts++);
    publish();
}

}  // namespace sst
