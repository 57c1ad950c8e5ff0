// Shared device helpers for the sm_100a mass-explanation kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sst {

constexpr int kWarp = 32;
constexpr int kMaxRows = 128;   // table rows incl. the leading 0 row; row masks are 128-bit
constexpr int kMaxDepth = 96;   // longest composition the enumerator can hold (nucleotides)

constexpr uint64_t kBit0Mask = 0x5555555555555555ULL;  // "reachable without this row" bit of every cell
constexpr uint64_t kBit1Mask = 0xAAAAAAAAAAAAAAAAULL;  // "one more copy of this row" bit of every cell

// acquire/release on global flags (tile hand-off in the table build)
__device__ __forceinline__ int ld_acquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// L2-only loads: data written by other SMs in the same launch must not be served from a stale L1 line
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
// streaming store: table words are written once and not re-read by this SM
__device__ __forceinline__ void st_cg_u64(uint64_t* p, uint64_t v) {
    asm volatile("st.global.cg.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// read-only 128-bit row-mask load
__device__ __forceinline__ uint4 ld_nc_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// 128-bit row-mask helpers (bit r <-> table row r, r in 1..127)
struct Mask128 {
    uint32_t w[4];
};
__device__ __forceinline__ Mask128 mk(uint4 v) {
    Mask128 m;
    m.w[0] = v.x; m.w[1] = v.y; m.w[2] = v.z; m.w[3] = v.w;
    return m;
}
__device__ __forceinline__ bool mask_empty(const Mask128& m) { return (m.w[0] | m.w[1] | m.w[2] | m.w[3]) == 0u; }
// keep rows <= r
__device__ __forceinline__ void mask_keep_le(Mask128& m, int r) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int lo = k * 32;
        uint32_t keep = (r >= lo + 31) ? 0xFFFFFFFFu : (r < lo ? 0u : (0xFFFFFFFFu >> (31 - (r - lo))));
        m.w[k] &= keep;
    }
}
// keep rows > r
__device__ __forceinline__ void mask_keep_gt(Mask128& m, int r) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int lo = k * 32;
        uint32_t drop = (r >= lo + 31) ? 0xFFFFFFFFu : (r < lo ? 0u : (0xFFFFFFFFu >> (31 - (r - lo))));
        m.w[k] &= ~drop;
    }
}
// pop the lowest set row; mask must be non-empty
__device__ __forceinline__ int mask_pop_lowest(Mask128& m) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (m.w[k]) {
            int b = __ffs(m.w[k]) - 1;
            m.w[k] &= m.w[k] - 1;
            return k * 32 + b;
        }
    }
    return -1;
}
__device__ __forceinline__ void mask_set(Mask128& m, int r) { m.w[r >> 5] |= 1u << (r & 31); }

// the two table bits of integer mass v in a packed 64-bit word (mass v%32 sits at bits 2*(31-v%32)+{1,0})
__device__ __forceinline__ uint32_t cell_bits(uint64_t word, int64_t v) {
    return (uint32_t)(word >> (2 * (31 - (int)(v & 31)))) & 3u;
}

}  // namespace sst
