// K3 + K2b for batches whose budgets cannot bind (every peak FREE — the production case): count -> scan -> fill with
// the COUNTS LOOKED UP instead of walked.
//
// Replaces explain_mass_with_table (reference mass_explanation.py:92-203; the window loop :192-201 and the inner
// backtrack :118-188).  In FREE mode the number of compositions below a partial composition (remainder m, largest row
// still allowed r) depends on (m, r) alone:
//     cnt(r, 0) = 1,   cnt(0, m > 0) = 0,   cnt(r, m) = cnt(r - 1, m) + [bit1(r, m)] * cnt(r, m - w_r)
// (UP then LEFT, the reference's own recursion :153-182 with the table bits as the guards; bit1(r, m) is bit r of the
// mass-major row mask H[m]).  k_count_row builds that table once per alphabet, one launch per row, for the masses
// below kCountMasses (u32, saturating): the same strided recurrence as the reachability table, on integers instead of
// bits.  With it
//   count phase   a peak's compositions = sum of cnt(top, m) over the reachable values m of its integer window: a
//                 gather from ONE L2-resident row (16 MB for 4.2 M masses) — no row-mask loads, no tree walk, no rounds;
//   scan          CTA totals -> grid barrier -> peak_off[p] (every CTA scans its own contiguous block of peaks);
//   fill phase    (after a second grid barrier) CTA b writes records [T*b/G, T*(b+1)/G) of the T records of the batch,
//                 whatever peaks they belong to: exact balance by OUTPUT, known before anything is written.  A peak
//                 with at most kSmall compositions that lies inside the range is written by one thread (closed forms
//                 for 1-2 nucleotides, an explicit-stack walk below).  Anything larger, or cut by the range, is split
//                 with the count table — peak -> window values -> children (m - w_r, r), each child's records start
//                 where the counts of its elder siblings end — until the pieces are small; pieces outside the range
//                 are dropped, pieces on its edge are walked whole and stored clipped.  The pieces live in two
//                 CTA-private bags in global memory (a LIFO of pieces still to be split, a bag of pieces to write);
//                 their order is irrelevant because every piece carries its absolute record offset.
// Records come out grouped by peak, window values ascending, rows ascending inside (depth-first order): the same
// order as the item pass (sst_enum.cuh).  Every writer compares what it wrote with the count it was promised; a
// mismatch (a table whose bits are not a consistent knapsack table), a saturated count or a bag overflow hands the
// batch to the item pass, which walks instead of looking up.
#pragma once
#include "sst_explain.cuh"

namespace sst {

constexpr int kDirThreads = 512;
constexpr int kDirWarps = kDirThreads / 32;
constexpr int kDirDepth = 16;              // longest composition this pass accepts (nucleotides; 8- or 16-byte records)
constexpr int kCntKeep = 8192;             // per-peak offsets a CTA keeps in shared memory between count and scan
constexpr unsigned kBagNarrow = 32768;     // LIFO of OPEN pieces a thread splits (few children)
constexpr unsigned kBagWide = 20480;       // LIFO of OPEN pieces a warp splits (many children)
constexpr unsigned kBagItems = kBagNarrow + kBagWide;
constexpr int kWidePerRound = 64;          // wide pieces split per round (a warp takes several)
constexpr unsigned kSmall = 32;            // a piece with at most this many records is written / split by one thread, a larger one by a warp
constexpr unsigned long long kSubRange = 24576;  // records a CTA fills between two fresh starts of its LIFOs
constexpr long long kPrefetchRoots = 1024; // roots whose row masks are asked for together before they are written
constexpr long long kRootChunk = 8192;     // window roots a CTA looks at between two looks at the LIFOs
constexpr int kLightRoots = 32;            // a peak with at most this many window roots lists them itself in the count phase
constexpr unsigned long long kRootGranule = 4096;  // root slots a CTA takes from the pool at a time
constexpr uint32_t kCntSat = 0xFFFFFFFFu;
constexpr int64_t kCountMasses = (int64_t)1 << 22;  // the count table covers integer masses below this (and below the table's width)

struct CountView {
    const uint32_t* c2d;  // [R][M] row-major
    int64_t M;
};

// fire-and-forget: bring the line into L2 (the row masks of a chunk of window values are asked for together, so that
// the threads that need them later do not pay one DRAM round trip after the other)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ uint32_t sat_add(uint32_t a, uint32_t b) {
    const uint32_t s = a + b;
    return s < a ? kCntSat : s;
}

// ---------------- compositions per call, looked up (no enumeration) ----------------
// out[p] = number of compositions explain_mass_with_table returns for call p when no modification budget binds: the sum
// of cnt(top, m) over the reachable values m of its integer window — one gather per reachable value from the top row of
// the count table.  ~0 (all ones) when the window reaches beyond the count table or a count saturated.  Used to cut a
// workload into blocks of equal OUTPUT before anything is enumerated (the host-side partition of SURVEY §8e).
__global__ void __launch_bounds__(256)
k_count_compositions(TableView tv, CountView cv, const double* __restrict__ mass, const double* __restrict__ thr, int64_t P, double precision,
                     double tolerance, unsigned long long* __restrict__ out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double m = mass[p], th = thr ? thr[p] : nan("");
    if (non_finite(m, th)) {
        out[p] = ~0ULL;
        return;
    }
    int64_t t, h;
    integerise(m, th, precision, tolerance, t, h);
    const int64_t limit = tv.C * 32;
    const uint64_t* last = tv.tbl + (int64_t)(tv.R - 1) * tv.C;
    const uint32_t* top = cv.c2d + (int64_t)(tv.R - 1) * cv.M;
    const int64_t lo = t - h, hi = t + h;
    const int64_t a = lo < 1 ? 1 : lo, b = hi < limit - 1 ? hi : limit - 1;
    unsigned long long sum = 0ULL;
    bool unknown = b >= cv.M && a <= b;
    if (!unknown)
        for_window_words(last, a, b, [&](int64_t wd, uint64_t x) {
            while (x) {
                const int pos = 63 - __clzll((long long)x);
                x &= ~(1ULL << pos);
                const uint32_t c = __ldg(top + (wd * 32 + (31 - (pos >> 1))));
                if (c == kCntSat) unknown = true;
                sum += c;
            }
        });
    out[p] = unknown ? ~0ULL : sum;
}

// ---------------- count table, one row per launch ----------------
__global__ void __launch_bounds__(256)
k_count_row0(uint32_t* __restrict__ row0, int64_t M) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) row0[m] = m == 0 ? 1u : 0u;
}

// thread j owns the residue class j mod w: cur[m] = prev[m] + (bit1(r, m) ? cur[m - w] : 0) along m = j, j + w, ...
// (adjacent threads touch adjacent masses: coalesced; the chain is at most M / w long)
__global__ void __launch_bounds__(256)
k_count_row(const uint64_t* __restrict__ tbl_row, const uint32_t* __restrict__ prev, uint32_t* __restrict__ cur, int64_t w, int64_t M) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= w || j >= M) return;
    uint32_t carry = 0u;
    for (int64_t m = j; m < M; m += w) {
        const uint32_t up = prev[m];
        const uint32_t bit = (uint32_t)(__ldg(tbl_row + (m >> 5)) >> (2 * (31 - (int)(m & 31)) + 1)) & 1u;
        const uint32_t v = (bit && m >= w) ? sat_add(up, carry) : up;
        cur[m] = v;
        carry = v;
    }
}

struct DirArgs {
    TableView tv;
    CountView cv;
    PeakBatch pk;
    uint8_t* status;                  // [P]
    uint32_t* rel;                    // [P] records of a peak's tile before the peak (only read back by the CTA that wrote them)
    unsigned long long* recs;         // [rec_capacity][nw] records, final order
    unsigned long long rec_capacity;
    unsigned long long* peak_off;     // [P+1]
    uint32_t* peak_off32;             // [P+1] the same as uint32 (may be null; only meaningful below 2^32 compositions)
    unsigned long long* cta_tot;      // [2][gridDim.x] compositions / window roots of every CTA's block of tiles
    // window roots of the batch, listed by the count phase tile by tile, peaks and window values ascending
    uint32_t* root_m;                 // mass
    uint32_t* root_pre;               // records of the tile before this root
    uint32_t* root_n;                 // records below it
    unsigned long long root_cap;
    int tile_size;                    // peaks per tile (a multiple of 32, at most kDirThreads; the host picks it so that every CTA has a tile)
    unsigned long long* tile_start;   // [n_tiles] where the roots of a tile (tile_size consecutive peaks) start in the root arrays
    uint32_t* tile_nroots;            // [n_tiles]
    unsigned long long* tile_base;    // [n_tiles + 1] records of the tiles before this one (count phase: records of the tile)
    uint32_t* bag_m;                  // [gridDim.x][kBagItems] the CTA-private LIFOs (first kBagNarrow: narrow, rest: wide)
    uint32_t* bag_meta;               // rmax | rows so far << 8
    uint32_t* bag_cnt;
    unsigned long long* bag_off;      // absolute index of the piece's first record
    unsigned long long* bag_path;     // [..][nw]
    unsigned int* sync;               // this launch's words: [0] barrier arrivals, [1] finished CTAs, [2..3] root cursor (u64), [4] fallback flag
    unsigned int* sync_next;          // the next launch's words: cleared by this one
    unsigned long long* host_out;     // pinned + mapped run summary (layout of PassSummary)
    LeafHash leaf;
    unsigned long long* cta_ns;       // diagnostics (may be null): [gridDim.x][8] %globaltimer of every CTA at its phase boundaries
};

__device__ __forceinline__ void dir_barrier(const DirArgs& a, unsigned int& gen) {
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

// first index i in [lo, hi) for which below(i) is false, hi if there is none (below is monotone: true ... true false
// ... false): the whole CTA probes kDirThreads evenly spaced entries per round (two rounds for 2.6 * 10^5 entries)
template <typename F>
__device__ __forceinline__ long long cta_search(long long lo, long long hi, F below) {
    while (lo < hi) {
        const long long step = (hi - lo + kDirThreads - 1) / kDirThreads;
        const long long i = lo + (long long)threadIdx.x * step;
        const bool is_below = i < hi && below(i);
        const int n_valid = (int)((hi - lo + step - 1) / step);
        const int n_below = __syncthreads_count(is_below);
        if (n_below < n_valid) {  // probe n_below is the first that is not below
            hi = lo + (long long)n_below * step;
            lo = n_below > 0 ? hi - step + 1 : hi;
        } else {
            lo = lo + (long long)(n_valid - 1) * step + 1;
        }
    }
    return hi;
}

// record idx of the batch, when it is one of this CTA's
template <int NW>
__device__ __forceinline__ void store_rec(unsigned long long* __restrict__ recs, unsigned long long idx, unsigned long long R0, unsigned long long R1,
                                          const unsigned long long* w) {
    if (idx >= R0 && idx < R1) {
#pragma unroll
        for (int q = 0; q < NW; q++) recs[idx * NW + q] = w[q];
    }
}

// Up to 32 closed forms, one per lane, written by the whole warp: a RECORD per lane.  A closed form is DONE / LEAF
// (m < 2 w_min: one record, the row of weight m appended) or POPC (m < 3 w_min: one record per enabled row r2 <= rmax
// of its row mask H[m], every child DONE or LEAF).  `has`: this lane holds a piece; `path` already holds the rows
// above, `off` is the absolute index of its first record.  The lanes' record counts are scanned, then lane l writes
// records l, l + 32, ... of the concatenation: it finds the piece that owns its record by a binary search over the
// scanned counts (shuffles), fetches the piece from its lane and selects the j-th enabled row — uniform work, stores
// that coalesce when the pieces are neighbours in the output.  Returns this lane's record count.
template <int NW>
__device__ __forceinline__ unsigned int emit_pieces32(const uint4* __restrict__ H, const RowTables& rt, bool has, uint32_t m, int rmax,
                                                      const unsigned long long* path, unsigned long long off, unsigned long long* __restrict__ recs,
                                                      unsigned long long R0, unsigned long long R1) {
    constexpr unsigned FULL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const bool popc_kind = has && m >= 2u * rt.wmin;
    Mask128 c;
    c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0u;
    if (popc_kind) {
        c = mk(ld_nc_u4(H + m));
        mask_keep_le(c, rmax);
    }
    const unsigned int n = !has ? 0u : (popc_kind ? (unsigned)mask_popc(c) : 1u);
    unsigned int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    const unsigned int pre = incl - n, S = __shfl_sync(FULL, incl, 31);
    for (unsigned int k0 = 0; k0 < S; k0 += 32) {
        const unsigned int k = k0 + lane;
        int o = 0;  // the last lane whose records start at or before k (it has one: k < S)
#pragma unroll
        for (int step = 16; step; step >>= 1) {
            const unsigned int pq = __shfl_sync(FULL, pre, (o + step) & 31);
            if (o + step < 32 && pq <= k) o += step;
        }
        const unsigned int j = k - __shfl_sync(FULL, pre, o);
        const uint32_t om = __shfl_sync(FULL, m, o);
        const int ormax = __shfl_sync(FULL, rmax, o);
        const unsigned long long ooff = __shfl_sync(FULL, off, o);
        Mask128 oc;
#pragma unroll
        for (int q = 0; q < 4; q++) oc.w[q] = __shfl_sync(FULL, c.w[q], o);
        unsigned long long w[NW];
#pragma unroll
        for (int q = 0; q < NW; q++) w[q] = __shfl_sync(FULL, path[q], o);
        if (k < S) {
            if (om < 2u * rt.wmin) {
                if (om) path_append(w, NW, leaf_row(rt.leaf, rt.w, rt.lh, om, ormax));
            } else {
                const int r2 = mask_select(oc, (int)j);
                const uint32_t m3 = om - (uint32_t)rt.w[r2];
                path_append(w, NW, r2);
                if (m3) path_append(w, NW, leaf_row(rt.leaf, rt.w, rt.lh, m3, r2));
            }
            store_rec<NW>(recs, ooff + j, R0, R1, w);
        }
    }
    return n;
}

template <int NW>
__global__ void __launch_bounds__(kDirThreads, 2)
k_explain_direct(const DirArgs a) {
    __shared__ int32_t s_w[kMaxRows];
    __shared__ uint8_t s_leaf[kLeafSlots];
    __shared__ uint32_t s_rel[kCntKeep];          // per-peak record offsets inside their tile
    __shared__ uint32_t s_defer[kDirThreads];
    __shared__ unsigned int s_pos[kDirThreads + 1];
    __shared__ unsigned int s_nnarrow, s_nwide, s_flag, s_ndefer, s_next;
    __shared__ unsigned long long s_region;
    __shared__ uint32_t s_it_m[kWidePerRound], s_it_meta[kWidePerRound], s_it_cnt[kWidePerRound];
    __shared__ unsigned long long s_it_off[kWidePerRound], s_it_path[kWidePerRound][NW];
    __shared__ PassSummary s_sum;  // only thread 0 of CTA 0 touches it
    const TableView& tv = a.tv;
    for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) s_w[i] = i < tv.R ? tv.weights[i] : 0;
    if (threadIdx.x == 0) s_nnarrow = s_nwide = s_flag = s_ndefer = 0u;
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
    auto cta_stamp = [&](int k) {
        if (a.cta_ns && threadIdx.x == 0) a.cta_ns[(size_t)blockIdx.x * 8 + k] = globaltimer_ns();
    };
    __syncthreads();
    leaf_table_init(s_leaf, s_w, tv.R, a.leaf);
    __syncthreads();
    const RowTables rt{s_w, nullptr, nullptr, s_leaf, a.leaf, tv.R > 1 ? (uint32_t)s_w[1] : 0u};
    const uint32_t wmin = rt.wmin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long P = a.pk.P;
    const int64_t limit = tv.C * 32;
    const int top_row = tv.R - 1;
    const uint64_t* __restrict__ last = tv.tbl + (int64_t)top_row * tv.C;
    const int64_t M = a.cv.M;
    const uint32_t* __restrict__ N = a.cv.c2d + (int64_t)top_row * M;  // compositions of every mass, all rows allowed
    unsigned long long* cursor = reinterpret_cast<unsigned long long*>(a.sync + 2);
    unsigned int* fallback = a.sync + 4;
    unsigned int gen = 0;
    int ts = 1;
    const unsigned G = gridDim.x, b = blockIdx.x;
    const int TS = a.tile_size;

    // the clamped integer window of peak p: [ca, cb], empty when ca > cb
    auto window_of = [&](long long p, int64_t& ca, int64_t& cb, uint8_t* st) {
        const int64_t tg = a.pk.target[p], th = a.pk.thr[p];
        const int64_t lo = tg - th, hi = tg + th;
        if (st) {
            uint8_t s = 0;
            if (lo <= 0 && 0 <= hi) s |= ST_ZERO_IN_WINDOW;
            if (lo <= hi && hi >= limit) s |= ST_OUT_OF_TABLE;
            *st = s;
        }
        ca = lo < 1 ? 1 : lo;
        cb = hi < limit - 1 ? hi : limit - 1;
    };
    auto masked_word = [&](int64_t wd, int64_t ca, int64_t cb) -> uint64_t {
        uint64_t x = __ldg(last + wd);
        x = (x | (x >> 1)) & kBit0Mask;
        if (wd == (ca >> 5)) x &= (1ULL << (2 * (31 - (int)(ca & 31)) + 1)) - 1ULL;
        if (wd == (cb >> 5)) x &= ~0ULL << (2 * (31 - (int)(cb & 31)));
        return x;
    };

    // ---------------- count phase: this CTA's contiguous block of tiles (a tile = TS consecutive peaks) ----------------
    // Per tile: (i) a thread per peak counts its window roots, a scan gives every peak its place in the tile's root
    // list; (ii) the masses of the roots are listed (a thread per peak, a warp for a peak with many); (iii) a thread
    // per ROOT looks its count up and a scan over the list gives every root the records of the tile before it — no
    // thread ever walks a list of loads; (iv) a peak's offset inside the tile is the prefix of its first root.
    cta_stamp(0);
    const long long n_tiles = (P + TS - 1) / TS;
    const long long tb = (long long)(((unsigned __int128)n_tiles * b) / G), te = (long long)(((unsigned __int128)n_tiles * (b + 1)) / G);
    const long long pb = tb * TS, pe = te * TS < P ? te * TS : P;
    const long long n_mine = pe > pb ? pe - pb : 0;
    {
        unsigned long long my_comps = 0, my_roots = 0;
        unsigned long long have_at = 0, have_n = 0;  // what is left of the last granule (every thread keeps the same copy)
        bool bad = false;
        for (long long t = tb; t < te; t++) {
            const long long j = (t - tb) * TS + threadIdx.x, p = pb + j;
            const bool mine = (int)threadIdx.x < TS && p < P;
            int64_t ca = 1, cb = 0;
            unsigned int nr = 0;
            if (mine) {
                uint8_t st;
                window_of(p, ca, cb, &st);
                a.status[p] = st;
                if (ca <= cb && cb >= M) {
                    bad = true;  // (the host checks the largest window end before it picks this pass)
                    cb = ca - 1;
                }
                for_window_words(last, ca, cb, [&](int64_t, uint64_t x) { nr += __popcll(x); });
            }
            unsigned int tile_roots;
            const unsigned int pos = block_scan32(nr, &tile_roots);
            s_pos[threadIdx.x] = pos;
            if (threadIdx.x == 0) {
                s_pos[kDirThreads] = tile_roots;
                s_ndefer = 0u;
                my_roots += tile_roots;
            }
            if (tile_roots > have_n) {  // uniform: a new granule from the pool (what was left of the old one stays unused)
                if (threadIdx.x == 0) s_region = atomicAdd(cursor, tile_roots > kRootGranule ? (unsigned long long)tile_roots : kRootGranule);
                __syncthreads();
                have_at = s_region;
                have_n = tile_roots > kRootGranule ? (unsigned long long)tile_roots : kRootGranule;
            }
            const unsigned long long region = have_at;
            have_at += tile_roots;
            have_n -= tile_roots;
            const bool fits = region + tile_roots <= a.root_cap;
            bad |= !fits;
            if (threadIdx.x == 0) {
                a.tile_start[t] = region;
                a.tile_nroots[t] = fits ? tile_roots : 0u;
            }
            __syncthreads();
            // (ii) the masses
            if (nr > 0u && fits) {
                if (nr <= (unsigned)kLightRoots) {
                    unsigned long long o = region + pos;
                    for_window_words(last, ca, cb, [&](int64_t wd, uint64_t x) {
                        while (x) {  // ascending mass = descending bit position
                            const int pos2 = 63 - __clzll((long long)x);
                            x &= ~(1ULL << pos2);
                            a.root_m[o++] = (uint32_t)(wd * 32 + (31 - (pos2 >> 1)));
                        }
                    });
                } else {
                    s_defer[atomicAdd(&s_ndefer, 1u)] = threadIdx.x;
                }
            }
            __syncthreads();
            {
                const unsigned int nd = s_ndefer;  // peaks with many roots: a warp each, a window value per lane
                for (unsigned int h = warp; h < nd; h += kDirWarps) {
                    const unsigned int owner = s_defer[h];
                    int64_t wa, wb;
                    window_of(pb + (t - tb) * TS + owner, wa, wb, nullptr);
                    unsigned long long at = region + s_pos[owner];
                    for (int64_t wd = wa >> 5; wd <= (wb >> 5); wd++) {
                        const uint64_t x = masked_word(wd, wa, wb);
                        const bool on = (x >> (2 * (31 - lane))) & 1ULL;
                        const unsigned int onmask = __ballot_sync(0xFFFFFFFFu, on);
                        if (on) a.root_m[at + __popc(onmask & ((1u << lane) - 1u))] = (uint32_t)(wd * 32 + lane);
                        at += __popc(onmask);
                    }
                }
                if (nd) __syncthreads();  // (uniform)
            }
            // (iii) counts and their running sum over the tile's list
            unsigned long long run = 0;
            if (fits) {
                for (unsigned int i0 = 0; i0 < tile_roots; i0 += kDirThreads) {
                    const unsigned int i = i0 + threadIdx.x;
                    uint32_t n = 0;
                    if (i < tile_roots) {
                        n = __ldg(N + __ldcg(a.root_m + region + i));
                        bad |= n == kCntSat;
                    }
                    unsigned long long tot;
                    const unsigned long long ex = block_scan((unsigned long long)n, &tot);
                    if (i < tile_roots) {
                        a.root_pre[region + i] = (uint32_t)(run + ex);
                        a.root_n[region + i] = n;
                    }
                    run += tot;
                }
                bad |= run >= 0xFFFFFFFFULL;  // (the offsets inside a tile are 32-bit)
            }
            __syncthreads();
            // (iv) where every peak starts inside the tile
            if (mine) {
                const uint32_t r0 = (fits && pos < tile_roots) ? __ldcg(a.root_pre + region + pos) : (uint32_t)run;
                if (j < kCntKeep) s_rel[j] = r0;
                else a.rel[p] = r0;
            }
            if (threadIdx.x == 0) {
                a.tile_base[t] = run;  // the tile's records; turned into the records before it after the grid barrier
                my_comps += run;
            }
        }
        if (bad) atomicExch(fallback, 1u);
        if (threadIdx.x == 0) {
            a.cta_tot[b] = my_comps;
            a.cta_tot[G + b] = my_roots;
        }
    }
    stamp(s_sum, ts++);
    cta_stamp(1);
    dir_barrier(a, gen);
    cta_stamp(2);
    stamp(s_sum, ts++);

    // ---------------- scan: record base of this CTA's tiles, peak offsets ----------------
    unsigned long long T, n_roots;
    {
        unsigned long long v[3] = {0ULL, 0ULL, 0ULL};
        for (unsigned k = threadIdx.x; k < G; k += kDirThreads) {
            const unsigned long long x = __ldcg(a.cta_tot + k);
            v[1] += x;
            if (k < b) v[0] += x;
            v[2] += __ldcg(a.cta_tot + G + k);
        }
        block_sum_n<3>(v);
        T = v[1];
        n_roots = v[2];
        const unsigned int fb = __ldcg(fallback);
        const unsigned long long asked = __ldcg(cursor);
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            s_sum.totals[0] = n_roots;
            s_sum.totals[1] = asked;                         // root slots asked for (granules included)
            s_sum.totals[2] = T;
            s_sum.totals[3] = 1ULL;
            if (asked > a.root_cap) s_sum.flags[2] = 1;      // root pool too small: the host grows it, runs again
            else if (fb) s_sum.flags[3] = 1;                 // a count out of range: the item pass walks instead
            if (T > a.rec_capacity) s_sum.flags[1] = 1;      // records do not fit: the host grows the buffer, runs again
        }
        if (fb || T > a.rec_capacity) {
            publish();
            return;
        }
        unsigned long long run = v[0];
        for (long long t = tb; t < te; t++) {
            const unsigned long long tile_recs = __ldcg(a.tile_base + t);  // written by this CTA's thread 0 before the barrier
            const long long j = (t - tb) * TS + threadIdx.x;
            if ((int)threadIdx.x < TS && j < n_mine) {
                const unsigned long long o = run + (j < kCntKeep ? s_rel[j] : a.rel[pb + j]);
                a.peak_off[pb + j] = o;
                if (a.peak_off32) a.peak_off32[pb + j] = (uint32_t)o;
            }
            __syncthreads();  // everybody has read the tile's records before they are replaced by its base
            if (threadIdx.x == 0) a.tile_base[t] = run;
            run += tile_recs;
        }
        if (b == G - 1 && threadIdx.x == 0) {
            a.peak_off[P] = T;
            if (a.peak_off32) a.peak_off32[P] = (uint32_t)T;
            a.tile_base[n_tiles] = T;
        }
    }
    cta_stamp(3);
    dir_barrier(a, gen);
    cta_stamp(4);
    stamp(s_sum, ts++);

    // ---------------- fill phase: records [Q0, Q1), at most kSubRange at a time ----------------
    // (a sub-range bounds the LIFOs: their pieces are disjoint and each holds a record of the sub-range)
    const unsigned long long Q0 = (unsigned long long)(((unsigned __int128)T * b) / G), Q1 = (unsigned long long)(((unsigned __int128)T * (b + 1)) / G);
    const uint4* __restrict__ H = tv.H;
    uint32_t* const bag_m = a.bag_m + (size_t)b * kBagItems;
    uint32_t* const bag_meta = a.bag_meta + (size_t)b * kBagItems;
    uint32_t* const bag_cnt = a.bag_cnt + (size_t)b * kBagItems;
    unsigned long long* const bag_off = a.bag_off + (size_t)b * kBagItems;
    unsigned long long* const bag_path = a.bag_path + (size_t)b * kBagItems * NW;
    bool wrong = false;
    unsigned long long R0 = Q0, R1 = Q0;  // the sub-range in hand

    // an OPEN piece (node m, rows <= rmax, `rows` rows so far in `path`, n records from `off` on) onto one of the LIFOs:
    // the narrow one (at most kSmall records: a lane splits it) or the wide one (a warp does); nothing when it lies
    // outside [R0, R1)
    auto push_piece = [&](uint32_t m, int rmax, int rows, const unsigned long long* path, unsigned long long off, uint32_t n) {
        if (n == 0u || off + n <= R0 || off >= R1) return;
        unsigned int at;
        if (n <= kSmall) {
            at = atomicAdd(&s_nnarrow, 1u);
            if (at >= kBagNarrow) {
                s_flag = 1u;
                return;
            }
        } else {
            at = atomicAdd(&s_nwide, 1u);
            if (at >= kBagWide) {
                s_flag = 1u;
                return;
            }
            at += kBagNarrow;
        }
        bag_m[at] = m;
        bag_meta[at] = (uint32_t)rmax | ((uint32_t)rows << 8);
        bag_cnt[at] = n;
        bag_off[at] = off;
#pragma unroll
        for (int q = 0; q < NW; q++) bag_path[(size_t)at * NW + q] = path[q];
    };

    // Up to 32 OPEN nodes with at most kSmall records each, one per lane (`has`), split by the whole warp in rounds:
    // in round q every lane takes its q-th child (ascending rows).  A child that is a closed form is a piece for
    // emit_pieces32 — the lanes' children of one round are written together, a record per lane; an OPEN child goes onto
    // a LIFO with its count from the table.  Either way the next sibling starts where this child's records end.
    auto split_lanes = [&](bool has, uint32_t m, int rmax, int rows, const unsigned long long* path, unsigned long long off, uint32_t n) {
        has = has && off + n > R0 && off < R1;
        Mask128 c;
        c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0u;
        if (has) {
            c = mk(ld_nc_u4(H + m));
            mask_keep_le(c, rmax);
        }
        unsigned long long at = off;
        while (__any_sync(0xFFFFFFFFu, !mask_empty(c))) {
            const bool go = !mask_empty(c);
            const int r = go ? mask_pop_lowest(c) : 0;
            const uint32_t m2 = go ? m - (uint32_t)s_w[r] : 0u;
            unsigned long long w[NW];
#pragma unroll
            for (int q = 0; q < NW; q++) w[q] = path[q];
            if (go) path_append(w, NW, r);
            const bool closed = go && m2 < 3u * wmin;
            uint32_t cn = 0;
            if (go && !closed) {
                cn = __ldg(a.cv.c2d + (int64_t)r * M + m2);
                push_piece(m2, r, rows + 1, w, at, cn);
            }
            const unsigned int ce = emit_pieces32<NW>(H, rt, closed, m2, r, w, at, a.recs, R0, R1);
            at += closed ? ce : cn;
        }
        if (has) wrong |= at - off != n;
    };

    // Both LIFOs until they are empty.  Every round: the youngest narrow pieces are split, 32 per warp (split_lanes);
    // then the youngest kWidePerRound wide pieces, a warp each (several per warp): lane l takes child rows l, l + 32,
    // ... — the closed forms among them are written together by emit_pieces32, the OPEN ones become pieces; a warp scan
    // over the children's counts gives every child its offset.
    auto drain = [&]() {
        for (;;) {
            __syncthreads();
            const unsigned int nn = s_nnarrow < kBagNarrow ? s_nnarrow : kBagNarrow;
            const unsigned int nw_ = s_nwide < kBagWide ? s_nwide : kBagWide;
            if ((nn == 0u && nw_ == 0u) || s_flag) break;
            const unsigned int take_n = nn < (unsigned)kDirThreads ? nn : (unsigned)kDirThreads;
            const unsigned int take_w = nw_ < (unsigned)kWidePerRound ? nw_ : (unsigned)kWidePerRound;
            uint32_t im = 0, imeta = 0, icnt = 0;
            unsigned long long ioff = 0, ipath[NW] = {};
            if (threadIdx.x < take_n) {
                const unsigned int at = nn - 1u - threadIdx.x;
                im = bag_m[at];
                imeta = bag_meta[at];
                icnt = bag_cnt[at];
                ioff = bag_off[at];
#pragma unroll
                for (int q = 0; q < NW; q++) ipath[q] = bag_path[(size_t)at * NW + q];
            }
            if (threadIdx.x < take_w) {  // staged in shared memory: their slots are free again
                const unsigned int at = kBagNarrow + nw_ - 1u - threadIdx.x;
                s_it_m[threadIdx.x] = bag_m[at];
                s_it_meta[threadIdx.x] = bag_meta[at];
                s_it_cnt[threadIdx.x] = bag_cnt[at];
                s_it_off[threadIdx.x] = bag_off[at];
#pragma unroll
                for (int q = 0; q < NW; q++) s_it_path[threadIdx.x][q] = bag_path[(size_t)at * NW + q];
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                s_nnarrow = nn - take_n;
                s_nwide = nw_ - take_w;
            }
            __syncthreads();
            if ((unsigned)(warp * 32) < take_n) split_lanes(threadIdx.x < take_n, im, (int)(imeta & 0xFFu), (int)(imeta >> 8), ipath, ioff, icnt);
            for (unsigned int it = warp; it < take_w; it += kDirWarps) {
                const uint32_t wm = s_it_m[it], wmeta = s_it_meta[it];
                const int rmax = (int)(wmeta & 0xFFu), rows = (int)(wmeta >> 8);
                const uint4 raw = ld_nc_u4(H + wm);
                const unsigned long long woff = s_it_off[it];
                const uint32_t mw[4] = {raw.x, raw.y, raw.z, raw.w};
                unsigned long long run = woff;
                for (int k = 0; k < 4 && k * 32 <= rmax; k++) {
                    const int r = k * 32 + lane;
                    const bool on = r <= rmax && ((mw[k] >> lane) & 1u);
                    const uint32_t m2 = on ? wm - (uint32_t)s_w[r] : 0u;
                    const bool closed = on && m2 < 3u * wmin;
                    Mask128 c2;
                    c2.w[0] = c2.w[1] = c2.w[2] = c2.w[3] = 0u;
                    uint32_t cn = 0;
                    if (closed) {
                        cn = 1u;
                        if (m2 >= 2u * wmin) {  // (emit_pieces32 loads it again: an L1 hit)
                            c2 = mk(ld_nc_u4(H + m2));
                            mask_keep_le(c2, r);
                            cn = (uint32_t)mask_popc(c2);
                        }
                    } else if (on) {
                        cn = __ldg(a.cv.c2d + (int64_t)r * M + m2);
                    }
                    unsigned long long incl = cn;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                        if (lane >= o) incl += y;
                    }
                    const unsigned long long at = run + incl - cn;
                    unsigned long long w[NW];
#pragma unroll
                    for (int q = 0; q < NW; q++) w[q] = s_it_path[it][q];
                    if (on) path_append(w, NW, r);
                    if (on && !closed) push_piece(m2, r, rows + 1, w, at, cn);
                    emit_pieces32<NW>(H, rt, closed, m2, r, w, at, a.recs, R0, R1);
                    run += __shfl_sync(0xFFFFFFFFu, incl, 31);
                }
                wrong |= run - woff != s_it_cnt[it];
            }
        }
    };

    const unsigned long long zero[NW] = {};
    cta_stamp(5);
    for (R0 = Q0; R0 < Q1 && !s_flag; R0 = R1) {
        R1 = R0 + kSubRange < Q1 ? R0 + kSubRange : Q1;
        // tiles that own a record of [R0, R1): the last one that starts at or before R0 .. the last that starts before R1
        const long long t_lo = cta_search(0, n_tiles + 1, [&](long long i) { return __ldcg(a.tile_base + i) <= R0; }) - 1;
        const long long t_hi = cta_search(t_lo, n_tiles + 1, [&](long long i) { return __ldcg(a.tile_base + i) < R1; });
        for (long long t = t_lo; t < t_hi && !s_flag; t++) {
            const unsigned long long rs = __ldcg(a.tile_start + t);
            const uint32_t* __restrict__ rm = a.root_m + rs;
            const uint32_t* __restrict__ rp = a.root_pre + rs;
            const uint32_t* __restrict__ rn = a.root_n + rs;
            const long long nr = (long long)__ldcg(a.tile_nroots + t);
            const unsigned long long tbase = __ldcg(a.tile_base + t), tnext = __ldcg(a.tile_base + t + 1);
            long long lo = 0, hi = nr;
            if (tbase < R0)  // first root with a record at or after R0
                lo = cta_search(0, nr, [&](long long i) { return tbase + __ldcg(rp + i) + __ldcg(rn + i) <= R0; });
            if (tnext > R1)  // first root that starts at or after R1
                hi = cta_search(lo, nr, [&](long long i) { return tbase + __ldcg(rp + i) < R1; });
            for (long long i0 = lo; i0 < hi && !s_flag; i0 += kRootChunk) {
                const long long i1 = i0 + kRootChunk < hi ? i0 + kRootChunk : hi;
                for (long long j0 = i0; j0 < i1; j0 += kPrefetchRoots) {
                    const long long j1 = j0 + kPrefetchRoots < i1 ? j0 + kPrefetchRoots : i1;
                    // the row masks this stretch of roots will need, asked for up front: first those of the roots, then
                    // (they are on their way or there) those of the children of the OPEN roots that are split here
                    for (long long i = j0 + threadIdx.x; i < j1; i += kDirThreads) {
                        const uint32_t m = __ldcg(rm + i);
                        if (m >= 2u * wmin) prefetch_l2(H + m);
                    }
                    for (long long i = j0 + threadIdx.x; i < j1; i += kDirThreads) {
                        const uint32_t m = __ldcg(rm + i);
                        if (m >= 3u * wmin && __ldcg(rn + i) <= kSmall) {
                            Mask128 c = mk(ld_nc_u4(H + m));
                            mask_keep_le(c, top_row);
                            while (!mask_empty(c)) {
                                const uint32_t m2 = m - (uint32_t)s_w[mask_pop_lowest(c)];
                                if (m2 >= 2u * wmin && m2 < 3u * wmin) prefetch_l2(H + m2);
                            }
                        }
                    }
                    // warps take 32 roots at a time from a shared counter: a warp that meets roots with many records
                    // takes fewer, and nobody waits at a barrier before the stretch is finished
                    if (threadIdx.x == 0) s_next = 0u;
                    __syncthreads();
                    for (;;) {
                        unsigned int c0 = 0;
                        if (lane == 0) c0 = atomicAdd(&s_next, 32u);
                        c0 = __shfl_sync(0xFFFFFFFFu, c0, 0);
                        if (j0 + c0 >= j1) break;
                        const long long i = j0 + c0 + lane;
                        const bool valid = i < j1;
                        const uint32_t m = valid ? __ldcg(rm + i) : 0u, n = valid ? __ldcg(rn + i) : 0u;
                        const unsigned long long off = tbase + (valid ? __ldcg(rp + i) : 0u);
                        const bool closed = valid && m < 3u * wmin;
                        // the closed forms among the 32 roots: written together, a record per lane
                        const unsigned int ce = emit_pieces32<NW>(H, rt, closed, m, top_row, zero, off, a.recs, R0, R1);
                        if (closed) wrong |= ce != n;
                        // the OPEN ones: few records -> split here (a child per lane and round), many -> a warp each, later
                        if (valid && !closed && n > kSmall) push_piece(m, top_row, 0, zero, off, n);
                        if (__any_sync(0xFFFFFFFFu, valid && !closed && n <= kSmall)) split_lanes(valid && !closed && n <= kSmall, m, top_row, 0, zero, off, n);
                    }
                    __syncthreads();  // (s_next is reset by thread 0 for the next stretch)
                }
                __syncthreads();
                if (s_nnarrow > 8192u || s_nwide + (unsigned)kRootChunk > kBagWide) drain();  // (uniform: read after the barrier)
                __syncthreads();
            }
        }
        __syncthreads();
        if (s_nnarrow || s_nwide) drain();
        __syncthreads();
    }
    cta_stamp(6);
    if (wrong) s_flag = 2u;
    __syncthreads();
    if (s_flag && threadIdx.x == 0) atomicExch(fallback, 1u + s_flag);
    cta_stamp(7);
    stamp(s_sum, ts++);
    // the fill phase's verdict reaches the host through CTA 0, which waits for the others (they do not wait for it)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.sync + 1, 1u);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            while (ld_relaxed_u32(a.sync + 1) < G) __nanosleep(20);
            __threadfence();
            if (__ldcg(fallback)) s_sum.flags[3] = 1;
            s_sum.totals[8 + ts] = globaltimer_ns();
        }
        __syncthreads();
        publish();
    }
}

}  // namespace sst
