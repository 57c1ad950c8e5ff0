// K1: 2-bit reachability table build (replaces set_up_bit_table, reference mass_table.py:207-248)
// K1t: row-major table -> mass-major 128-bit row masks used by the enumerator.
//
// Table semantics (row i >= 1, integer mass v; 32 masses per uint64, mass v%32 at bits 2*(31-v%32)+{1,0}):
//   bit0(i,v) = reach_{i-1}[v]          "v is a sum of weights of rows < i"
//   bit1(i,v) = reach_i[v - w_i]        "v - w_i is a sum of weights of rows <= i"
//   reach_0 = {0},  reach_i = reach_{i-1} U (reach_i + w_i)
// which is exactly what the reference's in-place ascending word loop leaves behind when every weight is
// >= 32 (step >= 1).  The last word of every row is AND-ed with the caller-supplied mask (:246).
//
// Parallel form: per word j, T_i(j) = funnel(x_i(j-step_i-1), x_i(j-step_i)) with x = (word|word>>1)&0x55..
// is the row's "bit1" contribution at even bit positions; bit0 of row i is reach_0 | OR_{i'<i} T_i'(j), a
// prefix-OR over rows.  A tile = all rows x 32 consecutive words; tile t only reads words
// <= 32t+31-step_min, i.e. tiles that are ~step_min/32 behind, so tiles are taken in ascending order and
// completion flags are checked before reading (no grid barrier).
#pragma once
#include "sst_common.cuh"

namespace sst {

constexpr int kBuildMaxWarps = 16;  // row groups per CTA (one warp each)
constexpr int kTileWords = 32;

// named barriers (0 is __syncthreads)
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// reach bits (bit0|bit1 of every cell, at the even bit positions) of one packed word, as two 32-bit halves.
// A cell never straddles the halves, so the shift by one stays inside each half.
struct Reach64 {
    uint32_t lo, hi;
};
__device__ __forceinline__ Reach64 reach_of(uint64_t w) {
    const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    Reach64 x;
    x.lo = (lo | (lo >> 1)) & 0x55555555u;
    x.hi = (hi | (hi >> 1)) & 0x55555555u;
    return x;
}

// 32x32 bit transpose across a warp (lane = row).  Stage s swaps the off-diagonal s x s blocks: a lane with bit s
// set takes (partner >> s) into the positions whose bit s is clear, a lane without it takes (partner << s) into
// the others.  On exactly those positions a shift equals a ROTATE, so each stage is one shuffle, one funnel shift
// by a per-lane amount and one bit-select LOP3 — amounts and masks are computed once per thread.
struct TransposePlan {
    uint32_t rot[5];   // rotate-left amount of stage k (s = 16 >> k)
    uint32_t keep[5];  // bits this lane keeps from its own word in stage k
};
__device__ __forceinline__ TransposePlan transpose_plan(int lane) {
    TransposePlan p;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const int s = 16 >> k;
        const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        p.rot[k] = (lane & s) ? 32 - s : s;
        uint32_t keep = (lane & s) ? ~m : m;
        asm volatile("mov.b32 %0, %0;" : "+r"(keep));  // keep it a register: ptxas otherwise rebuilds it as m ^ flag, 3 LOP3 per stage
        p.keep[k] = keep;
    }
    return p;
}
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, const TransposePlan& p) {
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, 16 >> k);
        const uint32_t yr = __funnelshift_l(y, y, p.rot[k]);
        asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(x) : "r"(x), "r"(yr), "r"(p.keep[k]));  // keep ? x : yr
    }
    return x;
}

// One tile of the build for one row group.  CHECKED=false is the interior fast path (every source word
// exists, no last column); CHECKED=true handles the first tiles and the last one.
// FUSE: the row's "bit1" word of every (row, table word) also goes to s_c — packed like k_transpose_masks packs it,
// P = (Thi << 1) | Tlo, masked like the table word in the last column — for the mass-major row masks the kernel writes
// after the tile has been handed on.
template <int RPW, bool CHECKED, bool FUSE>
__device__ __forceinline__ void build_tile_rows(uint64_t* __restrict__ tbl, int R, int64_t C, int64_t j, int lane, int g,
                                                const uint32_t (&src)[RPW], const int* s_step, const int* s_sh2,
                                                uint64_t last_mask, uint64_t (*s_tot)[32], int n_data_threads, uint32_t (*s_c)[33]) {
    // phase 1: gather every row's shifted reach word (independent L2 loads)
    uint64_t a[RPW], b0[RPW];
#pragma unroll
    for (int k = 0; k < RPW; k++) {
        const int r = 1 + g * RPW + k;
        a[k] = 0;
        b0[k] = 0;
        if (r < R) {
            if (CHECKED) {
                const int64_t col = j - s_step[r];
                if (col >= 0 && col < C) a[k] = __ldcg(tbl + src[k]);
                if (lane == 0 && col - 1 >= 0 && col - 1 < C) b0[k] = __ldcg(tbl + src[k] - 1);
            } else {
                a[k] = __ldcg(tbl + src[k]);
                if (lane == 0) b0[k] = __ldcg(tbl + src[k] - 1);
            }
        }
    }
    uint32_t Tlo[RPW], Thi[RPW];
    uint32_t tot_lo = 0, tot_hi = 0;
#pragma unroll
    for (int k = 0; k < RPW; k++) {
        const int r = 1 + g * RPW + k;
        const Reach64 xa = reach_of(a[k]);
        Reach64 xb;
        xb.lo = __shfl_up_sync(0xFFFFFFFFu, xa.lo, 1);
        xb.hi = __shfl_up_sync(0xFFFFFFFFu, xa.hi, 1);
        if (lane == 0) xb = reach_of(b0[k]);
        const int sh2 = r < R ? s_sh2[r] : 0;  // warp-uniform
        // (xb:xa) >> sh2, low 64 bits
        if (sh2 < 32) {
            Tlo[k] = __funnelshift_r(xa.lo, xa.hi, sh2);
            Thi[k] = __funnelshift_r(xa.hi, xb.lo, sh2);
        } else {
            Tlo[k] = __funnelshift_r(xa.hi, xb.lo, sh2 - 32);
            Thi[k] = __funnelshift_r(xb.lo, xb.hi, sh2 - 32);
        }
        tot_lo |= Tlo[k];
        tot_hi |= Thi[k];
        if (FUSE && r < R) {
            uint32_t P = (Thi[k] << 1) | Tlo[k];
            if (CHECKED) {
                if (j >= C) P = 0u;
                else if (j == C - 1) P &= ((uint32_t)(last_mask >> 32) & 0xAAAAAAAAu) | (((uint32_t)last_mask & 0xAAAAAAAAu) >> 1);
            }
            s_c[r][lane] = P;
        }
    }
    // phase 2: prefix-OR across the row groups (data warps only)
    s_tot[g][lane] = ((uint64_t)tot_hi << 32) | tot_lo;
    bar_sync(1, n_data_threads);
    uint64_t carry = (j == 0) ? 0x4000000000000000ULL : 0ULL;  // reach_0 = {0}
    for (int gg = 0; gg < g; gg++) carry |= s_tot[gg][lane];
    uint32_t c_lo = (uint32_t)carry, c_hi = (uint32_t)(carry >> 32);
    // phase 3: interleave and store
    if (!CHECKED || j < C) {
        if (g == 0) {
            uint64_t w0 = (j == 0) ? 0xC000000000000000ULL : 0ULL;
            if (CHECKED && j == C - 1) w0 &= last_mask;
            st_cg_u64(tbl + j, w0);
        }
#pragma unroll
        for (int k = 0; k < RPW; k++) {
            const int r = 1 + g * RPW + k;
            if (r < R) {
                uint64_t out = ((uint64_t)(c_hi | (Thi[k] << 1)) << 32) | (c_lo | (Tlo[k] << 1));
                if (CHECKED && j == C - 1) out &= last_mask;
                st_cg_u64(tbl + (src[k] + (uint32_t)s_step[r]), out);
                c_lo |= Tlo[k];
                c_hi |= Thi[k];
            }
        }
    }
}

// POLICY is 0 in the product.  tools/k1_probe.cu instantiates the (incorrect) variants 1 = no polling,
// 2 = no release fence, to measure what each mechanism costs.
template <int RPW, int POLICY = 0, bool FUSE = false>
__global__ void __launch_bounds__(kBuildMaxWarps * 32, 2)
k_build_table(uint64_t* __restrict__ tbl, int R, int64_t C, const int32_t* __restrict__ g_step,
              const int32_t* __restrict__ g_shift, uint64_t last_mask, int n_tiles, int* __restrict__ flags, uint4* __restrict__ H) {
    // blockDim = 32 * nd, nd = ceil((R-1)/RPW): one warp per group of RPW rows.
    // flags[t * nd + g] = 1 once row group g of tile t is stored.  Row r only ever reads row r of earlier
    // tiles, so the hand-off is per (tile, row group): each warp polls the <= 2*RPW flags its own rows need
    // and publishes its own flag; the only CTA-wide barrier is the prefix-OR exchange.  Tiles are assigned
    // round-robin (t = blockIdx + k*gridDim): the launch is cooperative so that all CTAs are co-resident
    // and a waiting CTA can never starve the one it waits for.
    // (Measured alternatives, tools/k1_probe.cu: a ticket per tile + one flag per tile, and a dedicated
    // publisher warp that takes the gpu-scope fence off the data warps, were both slower.)
    __shared__ uint64_t s_tot[2][kBuildMaxWarps][32];
    // FUSE: the tile's packed bit1 words, [row][table word], double-buffered like s_tot: a warp that is one tile ahead writes
    // the other buffer (it cannot be two ahead: the exchange barrier of a tile is passed by all warps together, and every
    // warp transposes tile k before it arrives at the barrier of tile k + 1).  Rows that do not exist stay zero.
    __shared__ uint32_t s_c[FUSE ? 2 : 1][FUSE ? kMaxRows : 1][33];
    __shared__ uint32_t s_srcoff[kMaxRows];  // r*C - step_r as a word index (R*C < 2^31, checked by the host)
    __shared__ int s_sh2[kMaxRows];
    __shared__ int s_step[kMaxRows];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int nd = blockDim.x >> 5;
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        s_step[i] = g_step[i];
        s_sh2[i] = 2 * g_shift[i];
        s_srcoff[i] = (uint32_t)((int64_t)i * C - g_step[i]);
    }
    if (FUSE)
        for (int i = threadIdx.x; i < 2 * kMaxRows * 33; i += blockDim.x) (&s_c[0][0][0])[i] = 0u;
    __syncthreads();
    const TransposePlan plan = transpose_plan(lane);
    const int mass_in_word = (lane & 1) ? 15 - (lane >> 1) : 31 - (lane >> 1);  // which mass bit `lane` of a packed word is
    const int step_max = s_step[R - 1];  // weights ascend
    // which row does this lane poll for?  lanes [0,RPW) the tile of word (j0 - step - 1), lanes [16,16+RPW) of (j0 + 31 - step)
    const int poll_k = lane & 15;
    const int poll_r = 1 + g * RPW + poll_k;
    const bool polls = poll_k < RPW && poll_r < R;
    const int poll_step = polls ? s_step[poll_r] : 0;
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, k++) {
        const int64_t j0 = (int64_t)t * kTileWords;
        const int64_t j = j0 + lane;
        uint32_t src[RPW];
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int r = 1 + g * RPW + q;
            src[q] = (r < R ? s_srcoff[r] : 0u) + (uint32_t)j;
        }
        if (polls && !(POLICY & 1)) {
            const int64_t w = (lane < 16) ? (j0 - poll_step - 1) : (j0 + (kTileWords - 1) - poll_step);
            if (w >= 0) {
                const int* f = flags + (w / kTileWords) * nd + g;
                while (ld_acquire(f) == 0) {
                }
            }
        }
        __syncwarp();
        const bool interior = (j0 - step_max - 1 >= 0) && (j0 + kTileWords < C);
        if (interior) build_tile_rows<RPW, false, FUSE>(tbl, R, C, j, lane, g, src, s_step, s_sh2, last_mask, s_tot[k & 1], nd * 32, s_c[FUSE ? (k & 1) : 0]);
        else build_tile_rows<RPW, true, FUSE>(tbl, R, C, j, lane, g, src, s_step, s_sh2, last_mask, s_tot[k & 1], nd * 32, s_c[FUSE ? (k & 1) : 0]);
        __syncwarp();
        if (lane == 0) {  // release: the group's stores (ordered before this by the warp barrier) become visible first
            if (POLICY & 2) *(volatile int*)(flags + (int64_t)t * nd + g) = 1;
            else st_release(flags + (int64_t)t * nd + g, 1);
        }
        if (FUSE) {  // off the dependency chain: the tile has been handed on.  Warp g takes table words g, g + nd, ... of the tile
            uint32_t (*c)[33] = s_c[k & 1];
            for (int l = g; l < kTileWords; l += nd) {
                uint4 d;
                d.x = warp_transpose32(c[lane][l], plan);
                d.y = warp_transpose32(c[32 + lane][l], plan);
                d.z = warp_transpose32(c[64 + lane][l], plan);
                d.w = warp_transpose32(c[96 + lane][l], plan);
                // lane q now holds, for bit q of the packed words, the rows that have it set; evict-first: the row masks must
                // not push the table words the next generations read out of L2
                if (j0 + l < C) __stcs(H + (j0 + l) * 32 + mass_in_word, d);
            }
        }
    }
}

// Generic single-CTA build for tables whose smallest weight is below 32*32 (tile hand-off needs a full
// tile of distance) — test alphabets only.  Row by row, target slabs of `step` words in ascending order.
__global__ void __launch_bounds__(1024)
k_build_table_small(uint64_t* __restrict__ tbl, int R, int64_t C, const int32_t* __restrict__ g_step,
                    const int32_t* __restrict__ g_shift, uint64_t last_mask) {
    for (int64_t j = threadIdx.x; j < C; j += blockDim.x) tbl[j] = (j == 0) ? 0xC000000000000000ULL : 0ULL;
    __syncthreads();
    for (int i = 1; i < R; i++) {
        uint64_t* row = tbl + (int64_t)i * C;
        const uint64_t* prev = tbl + (int64_t)(i - 1) * C;
        for (int64_t j = threadIdx.x; j < C; j += blockDim.x) {
            uint64_t p = prev[j];
            row[j] = (p | (p >> 1)) & kBit0Mask;
        }
        __syncthreads();
        const int64_t step = g_step[i];
        const int sh2 = 2 * g_shift[i];
        for (int64_t base = step; base < C; base += step) {
            const int64_t end = base + step < C ? base + step : C;
            for (int64_t jj = base + threadIdx.x; jj < end; jj += blockDim.x) {
                const uint64_t a = row[jj - step];
                const uint64_t b = (jj - step - 1 >= 0) ? row[jj - step - 1] : 0ULL;
                const uint64_t xa = (a | (a >> 1)) & kBit0Mask, xb = (b | (b >> 1)) & kBit0Mask;
                const uint64_t T = sh2 ? ((xa >> sh2) | (xb << (64 - sh2))) : xa;
                row[jj] |= T << 1;
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < R; i += blockDim.x) tbl[(int64_t)i * C + C - 1] &= last_mask;
}

// ---------------- K1t: mass-major row masks ----------------
// H[v] bit r = bit1(r, v) of the final (masked) table, r in 1..R-1.  One CTA per 32-word tile:
// coalesced row reads -> one 32-bit "bit1" word per (row, table word) in shared memory -> 32x32 bit transposes by
// warp shuffles -> one 16-byte store per mass.
//
// The kernel is bound by the (half-rate) integer ALU pipe, so the packing step matters: instead of compressing
// the 16 odd bits of each 32-bit half (10 operations per half), the two halves of a table word are INTERLEAVED —
// P = (hi & 0xAAAAAAAA) | ((lo & 0xAAAAAAAA) >> 1), one shift and one LOP3 — and the permutation this leaves in
// the bit order is undone for free in the store address: bit b of P is mass 15 - b/2 of the word if b is odd,
// 31 - b/2 if it is even.

__global__ void __launch_bounds__(256)
k_transpose_masks(const uint64_t* __restrict__ tbl, int R, int64_t C, uint4* __restrict__ H) {
    __shared__ uint32_t s_c[kMaxRows][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    {
        const int64_t j = j0 + lane;
        const bool col_ok = j < C;
        const uint64_t* src = tbl + (int64_t)w * C + j;
        const int64_t stride = 8 * C;
#pragma unroll 4
        for (int r = w; r < kMaxRows; r += 8, src += stride) {
            uint32_t c = 0;
            if (col_ok && r >= 1 && r < R) {
                const uint64_t word = __ldcs(src);
                c = ((uint32_t)(word >> 32) & 0xAAAAAAAAu) | (((uint32_t)word & 0xAAAAAAAAu) >> 1);
            }
            s_c[r][lane] = c;
        }
    }
    __syncthreads();
    const int mass_in_word = (lane & 1) ? 15 - (lane >> 1) : 31 - (lane >> 1);  // which mass bit `lane` of P is
    const TransposePlan plan = transpose_plan(lane);
#pragma unroll
    for (int l = w; l < 32; l += 8) {
        const int64_t j = j0 + l;
        uint4 d;
        d.x = warp_transpose32(s_c[lane][l], plan);
        d.y = warp_transpose32(s_c[32 + lane][l], plan);
        d.z = warp_transpose32(s_c[64 + lane][l], plan);
        d.w = warp_transpose32(s_c[96 + lane][l], plan);
        // lane q now holds, for bit q of the packed words, the rows that have it set
        if (j < C) H[j * 32 + mass_in_word] = d;
    }
}

}  // namespace sst
