// N3 + N4 on the device: the ladder-difference generator and the explanation-based alphabet reduction.
//
// Replaces, for a frame of classified fragments that stays in device memory between the rounds,
//   Predictor.collect_explanations_per_side   (reference prediction.py:286-329: the two-pointer window over the sorted
//                                              standard-unit masses of a side, `diff > max_weight` restart, the l1 error
//                                              threshold of common.py:37-44, `explanations[diff] = expl` dedup by key)
//   Predictor.collect_diff_explanations_for_su (:261-284: START side, END side, then the singletons)
//   the observed-nucleoside union of Predictor.filter_by_explanation (:170-202) and
//   Predictor._reduce_alphabet's re-validation of every fragment against the rebuilt table (:204-227).
// Per round only the 128-bit mask of observed table rows and two counters cross the bus.
//
// The window.  For a side with sorted masses su[0..n) the reference visits, in this order, the pairs
//     (s, s+1), (s, s+2), ... (s, E(s))         E(s) = last e with su[e] - su[s] <= max_weight      for s = 0, 1, ...
// until a pair (s0, n-1) has been visited; from then on `end` stays on the last fragment and only `start` moves:
//     (s0+1, n-1), (s0+2, n-1), ... (n-2, n-1).
// (prediction.py:325-328: `if end == len(fragments) - 1: start += 1`.)  So the number of pairs of s is known from
// E(s) and s0 = the first s with E(s) = n-1: count -> scan -> fill, the pairs in the reference's order.
//
// Dedup.  The reference keeps ONE entry per float key: a pair's explanation enters the side's dict only when it has at
// least one explanation and replaces an earlier entry with the same key; the END dict overrides the START dict, and a
// singleton always enters (even with None) and overrides both.  Calls are laid out START pairs, END pairs, singletons,
// so "last entering call with this key wins": one atomicMax per call on a hash table keyed by the key's bits.
#pragma once
#include "sst_explain.cuh"

namespace sst {

enum : int { FRAG_START = 1, FRAG_END = 2, FRAG_SINGLE = 4 };  // flags of a fragment (breakage contains START / END, is_singleton)
enum : int { CALL_ENTERS = 1, CALL_WINS = 2 };

struct LadderFrame {
    const double* su;      // [F] standard-unit masses, ascending
    const double* obs;     // [F] observed masses
    const uint8_t* flags;  // [F] FRAG_*
    uint8_t* alive;        // [F] 1 while the fragment has survived every re-validation
    int64_t F;
};

// hdr layout (unsigned long long): [0] START fragments, [1] END fragments, [2] singletons, [3] s0 of START, [4] s0 of END,
// [5] START pairs, [6] END pairs, [7] calls, [8..11] row mask, [12] alive fragments, [13] out-of-table flag of the
// re-validation, [14] a generated call's window reaches beyond the table
constexpr int kLadderHdrWords = 16;

// ---- alive fragments -> the three ordered lists (stable compaction; one CTA of kPassThreads threads)
__global__ void __launch_bounds__(kPassThreads)
k_ladder_sides(LadderFrame fr, uint32_t* __restrict__ idx, unsigned long long* __restrict__ hdr) {
    unsigned long long run[3] = {0ULL, 0ULL, 0ULL};
    for (int64_t base = 0; base < fr.F; base += kPassThreads) {
        const int64_t f = base + threadIdx.x;
        const int fl = (f < fr.F && fr.alive[f]) ? fr.flags[f] : 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const unsigned long long bit = (fl >> k) & 1;
            unsigned long long tot;
            const unsigned long long ex = block_scan(bit, &tot);
            if (bit) idx[(size_t)k * fr.F + run[k] + ex] = (uint32_t)f;
            run[k] += tot;
        }
    }
    if (threadIdx.x < 3) hdr[threadIdx.x] = run[threadIdx.x];
    if (threadIdx.x >= 3 && threadIdx.x < 5) hdr[threadIdx.x] = ~0ULL;  // s0: none yet
    if (threadIdx.x >= 5 && threadIdx.x < kLadderHdrWords && threadIdx.x != 12 && threadIdx.x != 13) hdr[threadIdx.x] = 0ULL;
}

// ---- E(s) of every side element; s0 by atomicMin.  blockIdx.y = side.
__global__ void __launch_bounds__(256)
k_ladder_reach(LadderFrame fr, const uint32_t* __restrict__ idx, unsigned long long* __restrict__ hdr, double max_weight,
               uint32_t* __restrict__ reach) {
    const int side = blockIdx.y;
    const int64_t n = (int64_t)hdr[side];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t* list = idx + (size_t)side * fr.F;
    const double base = fr.su[list[s]];
    // last e in (s, n) with su[e] - su[s] <= max_weight (the rounded difference is monotone in su[e])
    int64_t lo = s, hi = n - 1;  // invariant: lo qualifies (or is s itself), everything above hi does not
    while (lo < hi) {
        const int64_t mid = (lo + hi + 1) >> 1;
        if (!(__dsub_rn(fr.su[list[mid]], base) > max_weight)) lo = mid;
        else hi = mid - 1;
    }
    reach[(size_t)side * fr.F + s] = (uint32_t)lo;
    if (lo == n - 1 && lo > s) atomicMin(hdr + 3 + side, (unsigned long long)s);
}

__device__ __forceinline__ unsigned long long ladder_pairs_of(int64_t s, int64_t n, unsigned long long s0, uint32_t reach) {
    if ((unsigned long long)s <= s0) return (unsigned long long)(reach - (uint32_t)s);
    return s < n - 1 ? 1ULL : 0ULL;
}

// ---- pairs per element -> first call of every element (exclusive scan over START then END; one CTA)
__global__ void __launch_bounds__(kPassThreads)
k_ladder_scan(LadderFrame fr, unsigned long long* __restrict__ hdr, const uint32_t* __restrict__ reach, unsigned long long* __restrict__ first) {
    unsigned long long run = 0ULL;
    for (int side = 0; side < 2; side++) {
        const int64_t n = (int64_t)hdr[side];
        const unsigned long long s0 = hdr[3 + side];
        const unsigned long long before = run;
        for (int64_t base = 0; base < n; base += kPassThreads) {
            const int64_t s = base + threadIdx.x;
            const unsigned long long c = s < n ? ladder_pairs_of(s, n, s0, reach[(size_t)side * fr.F + s]) : 0ULL;
            unsigned long long tot;
            const unsigned long long ex = block_scan(c, &tot);
            if (s < n) first[(size_t)side * fr.F + s] = run + ex;
            run += tot;
        }
        if (threadIdx.x == 0) hdr[5 + side] = run - before;
    }
    if (threadIdx.x == 0) hdr[7] = run + hdr[2];
}

// ---- the calls: (difference, l1 threshold) of every pair in the reference's order, then the singletons.
// blockIdx.y = 0 / 1: the sides (thread per element, it writes all its pairs); 2: the singletons.
__global__ void __launch_bounds__(256)
k_ladder_fill(LadderFrame fr, const uint32_t* __restrict__ idx, const unsigned long long* __restrict__ hdr, const uint32_t* __restrict__ reach,
              const unsigned long long* __restrict__ first, double tolerance, double* __restrict__ mass, double* __restrict__ thr) {
    const int what = blockIdx.y;
    const int64_t n = (int64_t)hdr[what];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t* list = idx + (size_t)what * fr.F;
    const uint32_t f = list[s];
    if (what == 2) {
        const unsigned long long at = hdr[5] + hdr[6] + (unsigned long long)s;
        mass[at] = fr.su[f];
        thr[at] = __dmul_rn(tolerance, fr.obs[f]);
        return;
    }
    const unsigned long long s0 = hdr[3 + what];
    const uint32_t E = reach[(size_t)what * fr.F + s];
    const unsigned long long c = ladder_pairs_of(s, n, s0, E);
    unsigned long long at = first[(size_t)what * fr.F + s];
    const double su_s = fr.su[f], obs_s = fr.obs[f];
    for (unsigned long long j = 0; j < c; j++, at++) {
        const int64_t e = (unsigned long long)s <= s0 ? s + 1 + (int64_t)j : n - 1;
        const uint32_t g = list[e];
        mass[at] = __dsub_rn(fr.su[g], su_s);
        thr[at] = __dmul_rn(tolerance, __dadd_rn(obs_s, fr.obs[g]));  // calculate_error_threshold, l1 norm (common.py:37-44)
    }
}

// ---- dedup by key: the last entering call with a key wins
struct KeyTable {
    unsigned long long* keys;  // 0 = empty, else the key's bits + 1
    unsigned int* last;        // index + 1 of the last entering call
    uint32_t cap_mask;
};
__device__ __forceinline__ uint32_t key_slot(const KeyTable& kt, unsigned long long tag) {
    uint32_t h = (uint32_t)mix64(tag) & kt.cap_mask;
    for (;;) {
        const unsigned long long k = atomicCAS(kt.keys + h, 0ULL, tag);
        if (k == 0ULL || k == tag) return h;
        h = (h + 1) & kt.cap_mask;
    }
}
__global__ void __launch_bounds__(256)
k_ladder_enter(const double* __restrict__ mass, const unsigned long long* __restrict__ peak_off, const uint8_t* __restrict__ status,
               unsigned long long n_calls, unsigned long long n_pairs, KeyTable kt, uint8_t* __restrict__ call_flags,
               unsigned long long* __restrict__ hdr) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_calls) return;
    if (status[i] & ST_OUT_OF_TABLE) atomicOr(hdr + 14, 1ULL);  // calculate_explanations of this call raises upstream
    const bool enters = i >= n_pairs || peak_off[i + 1] > peak_off[i];  // a singleton always, a pair with >= 1 explanation
    call_flags[i] = enters ? CALL_ENTERS : 0;
    if (!enters) return;
    const unsigned long long tag = (unsigned long long)__double_as_longlong(mass[i]) + 1ULL;
    atomicMax(kt.last + key_slot(kt, tag), (unsigned int)i + 1u);
}
// winners -> CALL_WINS, and the union of the table rows their compositions use (hdr[8..11])
__global__ void __launch_bounds__(256)
k_ladder_union(const double* __restrict__ mass, const unsigned long long* __restrict__ peak_off, const unsigned long long* __restrict__ recs,
               int NW /* 8-byte words per record */, unsigned long long n_calls, KeyTable kt, uint8_t* __restrict__ call_flags,
               unsigned long long* __restrict__ hdr) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m[4] = {0u, 0u, 0u, 0u};
    if (i < n_calls && (call_flags[i] & CALL_ENTERS)) {
        const unsigned long long tag = (unsigned long long)__double_as_longlong(mass[i]) + 1ULL;
        if (kt.last[key_slot(kt, tag)] == (unsigned int)i + 1u) {
            call_flags[i] |= CALL_WINS;
            for (unsigned long long r = peak_off[i]; r < peak_off[i + 1]; r++) {
                for (int q = 0; q < NW; q++) {
                    unsigned long long w = recs[r * NW + q];
                    while (w) {  // one byte per nucleotide: the row index (0 = padding)
                        const unsigned row = (unsigned)(w & 0xFF);
                        w >>= 8;
                        if (row) m[row >> 5] |= 1u << (row & 31);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        uint32_t v = m[k];
#pragma unroll
        for (int o = 16; o; o >>= 1) v |= __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicOr(reinterpret_cast<unsigned int*>(hdr + 8) + k, v);
    }
}

// ---- N4: every alive fragment against the (rebuilt) table: is_valid_mass(su, tolerance * observed), prediction.py:216-223
__global__ void __launch_bounds__(256)
k_ladder_revalidate(TableView tv, LadderFrame fr, double precision, double tolerance, unsigned long long* __restrict__ hdr) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false, oot = false;
    if (f < fr.F && fr.alive[f]) {
        int64_t t, h;
        integerise(fr.su[f], __dmul_rn(tolerance, fr.obs[f]), precision, tolerance, t, h);
        const uint8_t code = valid_code(tv, t, h);
        keep = code == 1;
        oot = code == 2;
        if (!keep) fr.alive[f] = 0;
    }
    const unsigned n = __popc(__ballot_sync(0xFFFFFFFFu, keep));
    const unsigned any_oot = __ballot_sync(0xFFFFFFFFu, oot);
    if ((threadIdx.x & 31) == 0) {
        if (n) atomicAdd(hdr + 12, (unsigned long long)n);
        if (any_oot) atomicOr(hdr + 13, 1ULL);
    }
}

}  // namespace sst
