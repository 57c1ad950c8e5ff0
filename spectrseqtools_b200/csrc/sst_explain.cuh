// K2a validity probe, K2b enumerator (+ first-visit budget pass), K3 integer window filter.
//
// Replaces is_valid_mass (reference mass_explanation.py:45-89) and explain_mass_with_table
// (:92-203, inner backtrack :118-188).  Everything works on integer masses; the float -> integer
// conversion stays on the host so it matches CPython bit for bit.
//
// Enumeration model.  The reference walks (mass m, row r): UP to (m, r-1) if bit0, LEFT to (m-w_r, r) if
// bit1.  Because bit0(i,m) = OR_{r<i} bit1(r,m) (plus mass 0), the children of "mass m, rows <= rmax" are
// exactly {(m - w_r, r) : r <= rmax, bit1(r, m)}: one 16-byte load of the mass-major mask H[m] replaces
// the ~100 dependent UP reads per nucleotide.  A composition is emitted when the remainder hits 0.
//
// Budget modes (per peak):
//   FREE   budgets cannot bind (host-checked): enabled edges = table bits.
//   EXACT  with_memo=False: every path carries its own budgets (global `all`, per-row `ind`).
//   MEMO   with_memo=True with binding budgets: the reference's memo is keyed (m, r) WITHOUT budgets, so
//          each node is expanded once with the budgets of its first arrival in DFS order (UP before
//          LEFT, window ascending).  Phase A replays that order sequentially per peak, one mass at a
//          time (rows visited at a mass always form a contiguous range [r0(m), top(m)]), and records per
//          mass the LEFT edges that were enabled AND lead to at least one solution.  Phase B is the same
//          parallel path enumeration as FREE, reading those masks from a hash map instead of H.
#pragma once
#include "sst_common.cuh"

namespace sst {

enum : int { MODE_FREE = 0, MODE_EXACT = 1, MODE_MEMO = 2 };
enum : int { ST_ZERO_IN_WINDOW = 1, ST_OUT_OF_TABLE = 2 };
constexpr int kBudgetInf = 1 << 30;

struct TableView {
    const uint64_t* tbl;  // R x C, row-major (reference layout)
    const uint4* H;       // C*32 row masks
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
__device__ inline int memo_find_or_insert(const MemoMap& mp, unsigned long long key) {
    uint32_t h = (uint32_t)mix64(key) & mp.cap_mask;
    for (uint32_t probes = 0; probes < 4096; probes++) {
        unsigned long long k = mp.keys[h];
        if (k == key) return (int)h;
        if (k == 0ULL) {
            unsigned long long old = atomicCAS(mp.keys + h, 0ULL, key);
            if (old == 0ULL) {
                if (atomicAdd(mp.fill, 1u) > (mp.cap_mask >> 1) + (mp.cap_mask >> 2)) *mp.overflow = 1;
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
    int64_t t, h;
    integerise(mass[p], thr ? thr[p] : nan(""), precision, tolerance, t, h);
    out[p] = valid_code(tv, t, h);
}

// ---------------- single-pass chained scan (decoupled look-back) ----------------
// Every stage below is count -> exclusive scan -> fill in ONE launch: a CTA counts its tile, publishes the
// tile aggregate, looks back over earlier tiles for its exclusive prefix and then writes.  Tile ids come
// from a ticket so that a tile's predecessors have always started.  Flag and value share one 64-bit word
// (one atomic store), so no fence is needed.
constexpr unsigned long long kScanFlagA = 1ULL << 62;  // tile aggregate available
constexpr unsigned long long kScanFlagP = 1ULL << 63;  // inclusive prefix available
constexpr unsigned long long kScanValue = (1ULL << 62) - 1ULL;
constexpr int kTile = 128;

struct ScanState {
    unsigned long long* state;  // one word per tile, zeroed by the host before the launch
    unsigned int* ticket;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// exclusive scan of one value per thread across a 128-thread CTA; *total = CTA sum
__device__ __forceinline__ unsigned long long block_scan128(unsigned long long x, unsigned long long* total) {
    __shared__ unsigned long long s_warp[kTile / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += y;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    unsigned long long before = 0, sum = 0;
#pragma unroll
    for (int i = 0; i < kTile / 32; i++) {
        const unsigned long long v = s_warp[i];
        if (i < w) before += v;
        sum += v;
    }
    *total = sum;
    __syncthreads();
    return before + incl - x;
}

// called by warp 0 of the CTA; returns the sum of the aggregates of tiles [0, tile)
__device__ __forceinline__ unsigned long long tile_exclusive_prefix(const ScanState& ss, int tile, unsigned long long aggregate) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(ss.state, kScanFlagP | aggregate);
        return 0ULL;
    }
    if (lane == 0) st_volatile_u64(ss.state + tile, kScanFlagA | aggregate);
    unsigned long long excl = 0;
    for (int base = tile - 1;; base -= 32) {
        const int idx = base - lane;
        unsigned long long v;
        do {
            v = idx >= 0 ? ld_volatile_u64(ss.state + idx) : kScanFlagP;
        } while (__any_sync(0xFFFFFFFFu, (v & (kScanFlagA | kScanFlagP)) == 0ULL));
        const unsigned closed = __ballot_sync(0xFFFFFFFFu, (v & kScanFlagP) != 0ULL);
        const int first = closed ? __ffs(closed) - 1 : 32;  // nearest predecessor with a full prefix
        unsigned long long part = lane <= first ? (v & kScanValue) : 0ULL;
#pragma unroll
        for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        excl += part;
        if (closed) break;
    }
    if (lane == 0) st_volatile_u64(ss.state + tile, kScanFlagP | (excl + aggregate));
    return excl;
}

// CTA-wide: exclusive offset of this thread's `count` among all tiles; one ticket per call
__device__ __forceinline__ unsigned long long chained_offset(const ScanState& ss, int tile, unsigned long long count,
                                                             unsigned long long* tile_total) {
    __shared__ unsigned long long s_base;
    unsigned long long total;
    const unsigned long long excl = block_scan128(count, &total);
    if (threadIdx.x < 32) {
        const unsigned long long base = tile_exclusive_prefix(ss, tile, total);
        if (threadIdx.x == 0) s_base = base;
    }
    __syncthreads();
    const unsigned long long off = s_base + excl;
    *tile_total = total;
    __syncthreads();
    return off;
}

__device__ __forceinline__ int next_tile(const ScanState& ss) {
    __shared__ int s_tile;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ss.ticket, 1u);
    __syncthreads();
    const int t = s_tile;
    __syncthreads();
    return t;
}

// ---------------- K3 + root discovery: integer window over the last row ----------------
// One thread per peak: count the reachable window values (roots), get the offset, write them.
// totals[0] = number of roots.
__global__ void __launch_bounds__(kTile)
k_window_roots(TableView tv, PeakBatch pk, uint8_t* __restrict__ status, unsigned long long* __restrict__ root_off,
               uint32_t* __restrict__ root_v, uint32_t* __restrict__ root_peak, ScanState ss,
               unsigned long long* __restrict__ totals) {
    const int tile = next_tile(ss);
    const int64_t p = (int64_t)tile * kTile + threadIdx.x;
    const int64_t limit = tv.C * 32;
    const uint64_t* last = tv.tbl + (int64_t)(tv.R - 1) * tv.C;
    int64_t a = 1, b = 0;
    unsigned long long n = 0;
    if (p < pk.P) {
        const int64_t lo = pk.target[p] - pk.thr[p], hi = pk.target[p] + pk.thr[p];
        uint8_t st = 0;
        if (lo <= 0 && 0 <= hi) st |= ST_ZERO_IN_WINDOW;
        if (lo <= hi && hi >= limit) st |= ST_OUT_OF_TABLE;
        status[p] = st;
        a = lo < 1 ? 1 : lo;
        b = hi < limit - 1 ? hi : limit - 1;
        for (int64_t wd = a >> 5; a <= b && wd <= (b >> 5); wd++) {
            uint64_t x = __ldg(last + wd);
            x = (x | (x >> 1)) & kBit0Mask;
            if (wd == (a >> 5)) x &= (1ULL << (2 * (31 - (int)(a & 31)) + 1)) - 1ULL;
            if (wd == (b >> 5)) x &= ~0ULL << (2 * (31 - (int)(b & 31)));
            n += __popcll(x);
        }
    }
    unsigned long long tile_total;
    unsigned long long off = chained_offset(ss, tile, n, &tile_total);
    if (p < pk.P) {
        root_off[p] = off;
        if (p == pk.P - 1) {
            root_off[pk.P] = off + n;
            totals[0] = off + n;
        }
        for (int64_t wd = a >> 5; a <= b && wd <= (b >> 5); wd++) {
            uint64_t x = __ldg(last + wd);
            x = (x | (x >> 1)) & kBit0Mask;
            if (wd == (a >> 5)) x &= (1ULL << (2 * (31 - (int)(a & 31)) + 1)) - 1ULL;
            if (wd == (b >> 5)) x &= ~0ULL << (2 * (31 - (int)(b & 31)));
            while (x) {  // ascending mass = descending bit position
                const int pos = 63 - __clzll((long long)x);
                x &= ~(1ULL << pos);
                root_v[off] = (uint32_t)(wd * 32 + (31 - (pos >> 1)));
                root_peak[off] = (uint32_t)p;
                off++;
            }
        }
    }
}

// ---------------- MEMO phase A: sequential first-visit replay, one thread per peak ----------------
__global__ void __launch_bounds__(64)
k_memo_phase_a(TableView tv, RowMeta meta, PeakBatch pk, const uint32_t* __restrict__ memo_peaks, int n_memo,
               MemoMap mp) {
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

    // explicit recursion stack (one frame per mass on the current path)
    uint32_t f_m[kMaxDepth + 1];
    int f_slot[kMaxDepth + 1];
    uint8_t f_rin[kMaxDepth + 1], f_cur[kMaxDepth + 1];
    int f_all[kMaxDepth + 1], f_ind[kMaxDepth + 1];
    Mask128 f_pend[kMaxDepth + 1], f_new[kMaxDepth + 1];
    int sp = 0;

    // arrival at (m, r_in) with budgets; either answers from the map (returns false, sets alive) or
    // opens a frame for the rows (top(m), r_in] that this arrival visits for the first time
    auto arrive = [&](uint32_t m, int r_in, int all, int ind, bool& alive) -> bool {
        alive = false;
        const int slot = memo_find_or_insert(mp, memo_key(p, m));
        if (slot < 0 || sp > kMaxDepth) {
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
        f_m[sp] = m; f_slot[sp] = slot; f_rin[sp] = (uint8_t)r_in; f_all[sp] = all; f_ind[sp] = ind;
        f_pend[sp] = pend;
        f_new[sp].w[0] = f_new[sp].w[1] = f_new[sp].w[2] = f_new[sp].w[3] = 0u;
        sp++;
        return true;
    };

    for (int64_t v = a; v <= b; v++) {
        if (cell_bits(__ldg(last + (v >> 5)), v) == 0u) continue;
        bool alive;
        if (!arrive((uint32_t)v, top_row, pk.max_mods[p], s_ind[top_row], alive)) continue;
        while (sp > 0) {
            const int d = sp - 1;
            if (mask_empty(f_pend[d])) {  // all new rows of this mass expanded: publish and return
                const int slot = f_slot[d];
                uint4 A = mp.alive[slot];
                A.x |= f_new[d].w[0]; A.y |= f_new[d].w[1]; A.z |= f_new[d].w[2]; A.w |= f_new[d].w[3];
                mp.alive[slot] = A;
                mp.top[slot] = f_rin[d];
                Mask128 t = mk(A);
                mask_keep_le(t, f_rin[d]);
                const bool ok = !mask_empty(t);
                sp--;
                if (sp > 0 && ok) mask_set(f_new[sp - 1], f_cur[sp - 1]);
                continue;
            }
            const int r = mask_pop_lowest(f_pend[d]);  // LEFT edges fire in ascending row order (UP first)
            f_cur[d] = (uint8_t)r;
            const int ind_here = (r == f_rin[d]) ? f_ind[d] : s_ind[r];
            const int mod = s_mod[r];
            if (mod && !(f_all[d] > 0 && ind_here > 0)) continue;
            const uint32_t m2 = f_m[d] - (uint32_t)s_w[r];
            if (m2 == 0u) {
                mask_set(f_new[d], r);
                continue;
            }
            bool child_alive;
            if (!arrive(m2, r, f_all[d] - mod, ind_here - mod, child_alive)) {
                if (child_alive) mask_set(f_new[d], r);
            }
        }
    }
}

// ---------------- K2b: path enumeration ----------------
// Work items.  A root is one window value v with a non-empty last-row cell; an ITEM is (root, first row
// r1): the subtree of compositions whose largest row is r1.  Splitting at the first level turns the long
// serial chain of a heavy root (every mask load of a DFS depends on the previous pop) into many short
// chains, which is what the latency-bound 1-3 nt production calls need.  Output order = (peak, window
// value, first row, DFS order): grouped by peak, deterministic, no atomics on the data path.

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

struct ItemList {  // structure of arrays, capacity `cap`
    uint32_t* v;     // window value of the item's root
    uint32_t* peak;
    uint8_t* r;      // first (largest) row
    unsigned long long cap;
};

// roots -> items.  flags[2] is set when the items do not fit.  root_item_off[root] = first item of the
// root, root_item_off[n_roots] = totals[1] = number of items.
__global__ void __launch_bounds__(kTile)
k_root_items(TableView tv, PeakBatch pk, const uint32_t* __restrict__ root_v, const uint32_t* __restrict__ root_peak,
             const unsigned long long* __restrict__ totals_in, unsigned long long* __restrict__ root_item_off, ItemList items,
             MemoMap mp, ScanState ss, unsigned long long* __restrict__ totals, int* __restrict__ flags) {
    const int64_t n_roots = (int64_t)totals_in[0];
    if (n_roots == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            root_item_off[0] = 0ULL;
            totals[1] = 0ULL;
        }
        return;
    }
    const int n_tiles = (int)((n_roots + kTile - 1) / kTile);
    for (;;) {
        const int tile = next_tile(ss);
        if (tile >= n_tiles) break;
        const int64_t root = (int64_t)tile * kTile + threadIdx.x;
        Mask128 c;
        c.w[0] = c.w[1] = c.w[2] = c.w[3] = 0u;
        uint32_t v = 0, p = 0;
        if (root < n_roots) {
            v = root_v[root];
            p = root_peak[root];
            c = child_mask(tv, mp, pk.mode[p], p, v, tv.R - 1);
        }
        const unsigned long long n = (unsigned long long)(__popc(c.w[0]) + __popc(c.w[1]) + __popc(c.w[2]) + __popc(c.w[3]));
        unsigned long long tile_total;
        unsigned long long off = chained_offset(ss, tile, n, &tile_total);
        if (root < n_roots) {
            root_item_off[root] = off;
            if (root == n_roots - 1) {
                root_item_off[n_roots] = off + n;
                totals[1] = off + n;
            }
            if (off + n > items.cap) {
                flags[2] = 1;
            } else {
                while (!mask_empty(c)) {
                    items.v[off] = v;
                    items.peak[off] = p;
                    items.r[off] = (uint8_t)mask_pop_lowest(c);
                    off++;
                }
            }
        }
    }
}

// Per-thread DFS under one item.  FILL=false counts, FILL=true writes W-byte records from record index `out`.
template <bool FILL>
__device__ __forceinline__ unsigned long long enumerate_item(const TableView& tv, const MemoMap& mp, const int32_t* s_w,
                                                             const int32_t* s_ind, const uint8_t* s_mod, int mode, int64_t p,
                                                             uint32_t v, int r1, int max_mods, uint8_t* __restrict__ recs, int W,
                                                             unsigned long long out, unsigned long long per_item_cap,
                                                             int* __restrict__ flags) {
    uint32_t l_m[kMaxDepth];
    Mask128 l_mask[kMaxDepth];
    uint8_t l_path[kMaxDepth];
    int l_all[kMaxDepth], l_ind[kMaxDepth];
    const int top_row = tv.R - 1;
    const uint32_t two_wmin = tv.R > 1 ? 2u * (uint32_t)s_w[1] : 0u;  // below this a remainder is ONE nucleotide
    unsigned long long count = 0;
    uint64_t packed = 0;  // W == 8 fast path: rows so far, ascending from byte 0

    // write the composition l_path[0..n-1] (descending rows) as an ascending, 0-padded record
    auto emit = [&](int n, uint64_t packed_rec) {
        if (FILL) {
            uint8_t* rec = recs + out * (unsigned long long)W;
            if (W == 8) {
                *reinterpret_cast<uint64_t*>(rec) = packed_rec;
            } else {
                for (int q = 0; q < W; q += 8) {
                    uint64_t word = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const int idx = q + i;
                        if (idx < n) word |= (uint64_t)l_path[n - 1 - idx] << (8 * i);
                    }
                    *reinterpret_cast<uint64_t*>(rec + q) = word;
                }
            }
            out++;
        }
        count++;
    };

    // level 0 is the root restricted to this item's first row: no mask load needed
    int d = 0;
    l_m[0] = v;
    l_mask[0].w[0] = l_mask[0].w[1] = l_mask[0].w[2] = l_mask[0].w[3] = 0u;
    mask_set(l_mask[0], r1);
    l_all[0] = max_mods;
    l_ind[0] = s_ind[top_row];
    int rin = top_row;  // row by which the current level was entered (root: last row)

    for (;;) {
        if (mask_empty(l_mask[d])) {
            if (d == 0) break;
            d--;
            packed >>= 8;
            rin = d == 0 ? top_row : l_path[d - 1];
            continue;
        }
        const int r = mask_pop_lowest(l_mask[d]);
        int child_all = 0, child_ind = 0;
        if (mode == MODE_EXACT) {
            const int ind_here = (r == rin) ? l_ind[d] : s_ind[r];
            const int mod = s_mod[r];
            if (mod && !(l_all[d] > 0 && ind_here > 0)) continue;
            child_all = l_all[d] - mod;
            child_ind = ind_here - mod;
        }
        const uint32_t m2 = l_m[d] - (uint32_t)s_w[r];
        if (d + 2 >= kMaxDepth) continue;  // cannot happen: the host checks the depth bound before launch
        l_path[d] = (uint8_t)r;
        if (m2 == 0u) {
            emit(d + 1, (packed << 8) | (uint64_t)r);
        } else if (mode != MODE_MEMO && m2 < two_wmin) {
            // the table bit says m2 is a sum of rows <= r, and it is too light for two: m2 == w_q, q <= r
            int lo = 1, hi = r;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((uint32_t)s_w[mid] < m2) lo = mid + 1;
                else hi = mid;
            }
            const int q = lo;
            bool ok = (uint32_t)s_w[q] == m2;
            if (ok && mode == MODE_EXACT) {
                const int ind_q = (q == r) ? child_ind : s_ind[q];
                if (s_mod[q] && !(child_all > 0 && ind_q > 0)) ok = false;
            }
            if (ok) {
                l_path[d + 1] = (uint8_t)q;
                emit(d + 2, (((packed << 8) | (uint64_t)r) << 8) | (uint64_t)q);
            }
        } else {
            packed = (packed << 8) | (uint64_t)r;
            d++;
            rin = r;
            l_m[d] = m2;
            l_mask[d] = child_mask(tv, mp, mode, p, m2, r);
            l_all[d] = child_all;
            l_ind[d] = child_ind;
            continue;
        }
        if (!FILL && count > per_item_cap) {  // combinatorial blow-up guard (the reference would never return)
            flags[0] = 1;
            break;
        }
    }
    return count;
}

// items -> compositions: count, chained scan, fill, in one launch.  item_comp_off[item] = first record of the
// item, item_comp_off[n_items] = totals[2] = number of compositions.  flags[0]: blow-up guard hit;
// flags[1]: records do not fit rec_capacity (nothing is written past it; the host grows and reruns).
__global__ void __launch_bounds__(kTile)
k_enumerate(TableView tv, RowMeta meta, PeakBatch pk, ItemList items, const unsigned long long* __restrict__ totals_in,
            unsigned long long* __restrict__ item_comp_off, uint8_t* __restrict__ recs, int W, MemoMap mp,
            unsigned long long per_item_cap, unsigned long long rec_capacity, ScanState ss,
            unsigned long long* __restrict__ totals, int* __restrict__ flags) {
    __shared__ int32_t s_w[kMaxRows];
    __shared__ int32_t s_ind[kMaxRows];
    __shared__ uint8_t s_mod[kMaxRows];
    for (int i = threadIdx.x; i < kMaxRows; i += blockDim.x) {
        s_w[i] = i < tv.R ? tv.weights[i] : 0;
        s_ind[i] = i < tv.R ? meta.ind[i] : 0;
        s_mod[i] = i < tv.R ? meta.is_mod[i] : 0;
    }
    __syncthreads();
    const int64_t n_items = (int64_t)totals_in[1];
    if (n_items == 0 || (unsigned long long)n_items > items.cap) {  // nothing to do / item pass overflowed
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            item_comp_off[0] = 0ULL;
            totals[2] = 0ULL;
        }
        return;
    }
    const int n_tiles = (int)((n_items + kTile - 1) / kTile);
    for (;;) {
        const int tile = next_tile(ss);
        if (tile >= n_tiles) break;
        const int64_t item = (int64_t)tile * kTile + threadIdx.x;
        unsigned long long n = 0;
        uint32_t v = 0;
        int64_t p = 0;
        int r1 = 0, mode = MODE_FREE, max_mods = 0;
        if (item < n_items) {
            v = items.v[item];
            p = items.peak[item];
            r1 = items.r[item];
            mode = pk.mode[p];
            max_mods = pk.max_mods[p];
            n = enumerate_item<false>(tv, mp, s_w, s_ind, s_mod, mode, p, v, r1, max_mods, nullptr, W, 0ULL, per_item_cap, flags);
        }
        unsigned long long tile_total;
        const unsigned long long off = chained_offset(ss, tile, n, &tile_total);
        if (item < n_items) {
            item_comp_off[item] = off;
            if (item == n_items - 1) {
                item_comp_off[n_items] = off + n;
                totals[2] = off + n;
            }
            if (off + n > rec_capacity) flags[1] = 1;
            else if (n) enumerate_item<true>(tv, mp, s_w, s_ind, s_mod, mode, p, v, r1, max_mods, recs, W, off, ~0ULL, flags);
        }
    }
}

// per-peak composition offsets: peak_off[p] = item_comp_off[root_item_off[root_off[p]]], peak_off[P] = total
__global__ void k_peak_offsets(const unsigned long long* __restrict__ root_off, const unsigned long long* __restrict__ root_item_off,
                               const unsigned long long* __restrict__ item_comp_off, int64_t P, unsigned long long item_capacity,
                               unsigned long long* __restrict__ peak_off) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > P) return;
    unsigned long long i = root_item_off[root_off[p]];
    if (i > item_capacity) i = item_capacity;  // only after an item overflow; the host repeats the run
    peak_off[p] = item_comp_off[i];
}

}  // namespace sst
