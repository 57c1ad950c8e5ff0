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

constexpr int kBuildWarps = 8;
constexpr int kTileWords = 32;

template <int RPW>
__global__ void __launch_bounds__(kBuildWarps * 32)
k_build_table(uint64_t* __restrict__ tbl, int R, int64_t C, const int32_t* __restrict__ g_step,
              const int32_t* __restrict__ g_shift, uint64_t last_mask, int n_tiles, int* __restrict__ flags) {
    // flags[t * kBuildWarps + g] = 1 once row group g of tile t is stored.  Row r only ever reads row r
    // of earlier tiles, so the hand-off is per (tile, row group): each warp polls the <= 2*RPW flags its
    // own rows need and publishes its own flag; the only CTA-wide barrier is the prefix-OR exchange.
    // Tiles are assigned round-robin (t = blockIdx + k*gridDim): the launch is cooperative so that all
    // CTAs are co-resident and a waiting CTA can never starve the one it waits for.
    __shared__ uint64_t s_tot[2][kBuildWarps][32];
    __shared__ int s_step[kMaxRows];
    __shared__ int s_shift[kMaxRows];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        s_step[i] = g_step[i];
        s_shift[i] = g_shift[i];
    }
    __syncthreads();
    // which row does this lane poll for?  lanes [0,RPW) the tile of word (j0 - step - 1), lanes [16,16+RPW) of (j0 + 31 - step)
    const int poll_k = lane & 15;
    const int poll_r = 1 + g * RPW + poll_k;
    const bool polls = poll_k < RPW && poll_r < R;
    int buf = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, buf ^= 1) {
        const int64_t j0 = (int64_t)t * kTileWords;
        const int64_t j = j0 + lane;

        if (polls) {
            const int64_t src = (lane < 16) ? (j0 - s_step[poll_r] - 1) : (j0 + (kTileWords - 1) - s_step[poll_r]);
            if (src >= 0) {
                const int* f = flags + (src / kTileWords) * kBuildWarps + g;
                while (ld_acquire(f) == 0) {
                }
            }
        }
        __syncwarp();

        // phase 1: gather every row's shifted reach word (independent L2 loads)
        uint64_t a[RPW], b0[RPW];
#pragma unroll
        for (int k = 0; k < RPW; k++) {
            const int r = 1 + g * RPW + k;
            a[k] = 0;
            b0[k] = 0;
            if (r < R) {
                const int64_t src = j - s_step[r];
                const uint64_t* row = tbl + (int64_t)r * C;
                if (src >= 0 && src < C) a[k] = __ldcg(row + src);
                if (lane == 0 && src - 1 >= 0 && src - 1 < C) b0[k] = __ldcg(row + src - 1);
            }
        }
        uint64_t T[RPW];
        uint64_t wtot = 0;
#pragma unroll
        for (int k = 0; k < RPW; k++) {
            const int r = 1 + g * RPW + k;
            uint64_t b = __shfl_up_sync(0xFFFFFFFFu, a[k], 1);
            if (lane == 0) b = b0[k];
            const uint64_t xa = (a[k] | (a[k] >> 1)) & kBit0Mask;
            const uint64_t xb = (b | (b >> 1)) & kBit0Mask;
            const int sh2 = r < R ? 2 * s_shift[r] : 0;
            T[k] = sh2 ? ((xa >> sh2) | (xb << (64 - sh2))) : xa;
            wtot |= T[k];
        }
        // phase 2: prefix-OR across the row groups (double-buffered: one barrier per tile)
        s_tot[buf][g][lane] = wtot;
        __syncthreads();
        uint64_t carry = (j == 0) ? 0x4000000000000000ULL : 0ULL;  // reach_0 = {0}
        for (int gg = 0; gg < g; gg++) carry |= s_tot[buf][gg][lane];
        // phase 3: interleave and store
        if (j < C) {
            if (g == 0) {
                uint64_t w0 = (j == 0) ? 0xC000000000000000ULL : 0ULL;
                if (j == C - 1) w0 &= last_mask;
                st_cg_u64(tbl + j, w0);
            }
#pragma unroll
            for (int k = 0; k < RPW; k++) {
                const int r = 1 + g * RPW + k;
                if (r < R) {
                    uint64_t out = carry | (T[k] << 1);
                    if (j == C - 1) out &= last_mask;
                    st_cg_u64(tbl + (int64_t)r * C + j, out);
                    carry |= T[k];
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            st_release(flags + (int64_t)t * kBuildWarps + g, 1);
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
