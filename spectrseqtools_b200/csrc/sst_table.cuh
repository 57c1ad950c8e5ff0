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

// One tile of the build for one row group.  CHECKED=false is the interior fast path (every source word
// exists, no last column); CHECKED=true handles the first tiles and the last one.
template <int RPW, bool CHECKED>
__device__ __forceinline__ void build_tile_rows(uint64_t* __restrict__ tbl, int R, int64_t C, int64_t j, int lane, int g,
                                                const uint32_t (&src)[RPW], const int* s_step, const int* s_sh2,
                                                uint64_t last_mask, uint64_t (*s_tot)[32], int n_data_threads) {
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
template <int RPW, int POLICY = 0>
__global__ void __launch_bounds__(kBuildMaxWarps * 32, 2)
k_build_table(uint64_t* __restrict__ tbl, int R, int64_t C, const int32_t* __restrict__ g_step,
              const int32_t* __restrict__ g_shift, uint64_t last_mask, int n_tiles, int* __restrict__ flags) {
    // blockDim = 32 * nd, nd = ceil((R-1)/RPW): one warp per group of RPW rows.
    // flags[t * nd + g] = 1 once row group g of tile t is stored.  Row r only ever reads row r of earlier
    // tiles, so the hand-off is per (tile, row group): each warp polls the <= 2*RPW flags its own rows need
    // and publishes its own flag; the only CTA-wide barrier is the prefix-OR exchange.  Tiles are assigned
    // round-robin (t = blockIdx + k*gridDim): the launch is cooperative so that all CTAs are co-resident
    // and a waiting CTA can never starve the one it waits for.
    // (Measured alternatives, tools/k1_probe.cu: a ticket per tile + one flag per tile, and a dedicated
    // publisher warp that takes the gpu-scope fence off the data warps, were both slower.)
    __shared__ uint64_t s_tot[2][kBuildMaxWarps][32];
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
    __syncthreads();
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
        if (interior) build_tile_rows<RPW, false>(tbl, R, C, j, lane, g, src, s_step, s_sh2, last_mask, s_tot[k & 1], nd * 32);
        else build_tile_rows<RPW, true>(tbl, R, C, j, lane, g, src, s_step, s_sh2, last_mask, s_tot[k & 1], nd * 32);
        __syncwarp();
        if (lane == 0) {  // release: the group's stores (ordered before this by the warp barrier) become visible first
            if (POLICY & 2) *(volatile int*)(flags + (int64_t)t * nd + g) = 1;
            else st_release(flags + (int64_t)t * nd + g, 1);
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
// coalesced row reads -> 32-bit "bit1" columns in shared memory -> 32x32 bit transposes by warp shuffles
// -> one 16-byte store per mass.

__device__ __forceinline__ uint32_t compress_odd_bits(uint32_t v) {  // 16 odd bits -> low 16 bits
    v = (v >> 1) & 0x55555555u;
    v = (v | (v >> 1)) & 0x33333333u;
    v = (v | (v >> 2)) & 0x0F0F0F0Fu;
    v = (v | (v >> 4)) & 0x00FF00FFu;
    v = (v | (v >> 8)) & 0x0000FFFFu;
    return v;
}

__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xFFFFFFFFu, x, s);
        x = (lane & s) ? ((x & ~m) | ((y >> s) & m)) : ((x & m) | ((y << s) & ~m));
    }
    return x;
}

__global__ void __launch_bounds__(256)
k_transpose_masks(const uint64_t* __restrict__ tbl, int R, int64_t C, uint4* __restrict__ H) {
    __shared__ uint32_t s_c[kMaxRows][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t j0 = (int64_t)blockIdx.x * 32;
    for (int r = w; r < kMaxRows; r += 8) {
        uint32_t c = 0;
        const int64_t j = j0 + lane;
        if (r >= 1 && r < R && j < C) {
            const uint64_t word = __ldcs(tbl + (int64_t)r * C + j);
            c = (compress_odd_bits((uint32_t)(word >> 32)) << 16) | compress_odd_bits((uint32_t)word);
        }
        s_c[r][lane] = c;
    }
    __syncthreads();
    for (int l = w; l < 32; l += 8) {
        const int64_t j = j0 + l;
        if (j >= C) break;
        uint4 d;
        d.x = warp_transpose32(s_c[lane][l], lane);
        d.y = warp_transpose32(s_c[32 + lane][l], lane);
        d.z = warp_transpose32(s_c[64 + lane][l], lane);
        d.w = warp_transpose32(s_c[96 + lane][l], lane);
        // lane q now holds the rows of the mass whose compacted bit index is q, i.e. mass 31-q of word j
        H[j * 32 + (31 - lane)] = d;
    }
}

}  // namespace sst
