// Probe of a table build whose tile hand-off is SELF-VALIDATING MESSAGES instead of flags + release fences
// (k_build_table_ll, defined here), against the product's tile kernel (k_build_table): byte-for-byte comparison on
// the device and best-of-N timings.  Every table word's reach bits travel as one 8-byte message {packed reach, tag}
// in an L2-resident ring; a consumer polls the message itself, so there is no fence and no separate flag on the
// generation-to-generation latency chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/k1_ll_probe tools/k1_ll_probe.cu
//   tools/k1_ll_probe tools/weights_full.txt [reps] [rows]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../spectrseqtools_b200/csrc/sst_table.cuh"
namespace sst {

constexpr int kRingShift = 15;              // messages per row ring: 32768 words = 1024 tiles
constexpr int kRing = 1 << kRingShift;
constexpr int kLapBits = 12;

__device__ __forceinline__ uint64_t ld_msg(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_msg(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_prog(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_prog(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// POLICY: 0 product candidate; 1 = no polling (incorrect: timing only); 2 = no ring-reuse guard (timing only)
template <int RPW, int POLICY>
__global__ void __launch_bounds__(kBuildMaxWarps * 32, 2)
k_build_table_ll(uint64_t* __restrict__ tbl, int R, int64_t C, const int32_t* __restrict__ g_step,
                 const int32_t* __restrict__ g_shift, uint64_t last_mask, int n_tiles, uint64_t* __restrict__ msg,
                 uint32_t epoch, uint32_t* __restrict__ prog) {
    __shared__ uint64_t s_tot[2][kBuildMaxWarps][32];
    __shared__ int s_sh2[kMaxRows];
    __shared__ int s_step[kMaxRows];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int nd = blockDim.x >> 5;
    for (int i = threadIdx.x; i < R; i += blockDim.x) {
        s_step[i] = g_step[i];
        s_sh2[i] = 2 * g_shift[i];
    }
    __syncthreads();
    const int step_max = s_step[R - 1];  // weights ascend
    const int G = gridDim.x;
    const uint32_t tag_hi = epoch << kLapBits;
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += G, k++) {
        const int64_t j0 = (int64_t)t * kTileWords;
        const int64_t j = j0 + lane;
        // ring-reuse guard (one warp): the messages this tile overwrites were for words j - kRing; their readers are
        // the tiles up to (j0 + 31 - kRing + step_max + 1) / 32.  All of those must have finished READING (prog[c] =
        // tiles of CTA c whose reads are complete).  Normally true long ago: one round of relaxed loads, no waiting.
        if (!(POLICY & 2) && g == 0) {
            const int64_t xw = j0 + 31 - kRing + step_max + 1;
            if (xw >= 0) {
                const int X = (int)(xw / kTileWords);  // all tiles <= X
                for (int c = lane; c < G; c += 32) {
                    const uint32_t need = X >= c ? (uint32_t)((X - c) / G + 1) : 0u;
                    while (ld_prog(prog + c) < need) {
                    }
                }
            }
        }
        // phase 1: poll the messages of the source words (self-validating: the tag says which lap of the ring wrote it)
        uint32_t P[RPW], Ph[RPW];  // packed reach of word j - step_r, and (lane 0) of word j - step_r - 1
        bool pending = false;
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            P[q] = 0u;
            Ph[q] = 0u;
        }
        uint32_t todo = 0u;  // bit q: message q still missing; bit 16+q: halo message q still missing
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int r = 1 + g * RPW + q;
            if (r < R) {
                const int64_t js = j - s_step[r];
                if (js >= 0 && js < C) todo |= 1u << q;
                if (lane == 0 && js - 1 >= 0 && js - 1 < C) todo |= 1u << (16 + q);
            }
        }
        do {
            uint64_t v[RPW], vh[RPW];
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                const int r = 1 + g * RPW + q;
                const int64_t js = j - (r < R ? s_step[r] : 0);
                v[q] = 0;
                vh[q] = 0;
                if (todo & (1u << q)) v[q] = ld_msg(msg + (size_t)r * kRing + (js & (kRing - 1)));
                if (todo & (1u << (16 + q))) vh[q] = ld_msg(msg + (size_t)r * kRing + ((js - 1) & (kRing - 1)));
            }
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                const int r = 1 + g * RPW + q;
                const int64_t js = j - (r < R ? s_step[r] : 0);
                if (todo & (1u << q)) {
                    const uint32_t want = tag_hi | (uint32_t)((js >> kRingShift) + 1);
                    if ((POLICY & 1) || (uint32_t)(v[q] >> 32) == want) {
                        P[q] = (uint32_t)v[q];
                        todo &= ~(1u << q);
                    }
                }
                if (todo & (1u << (16 + q))) {
                    const uint32_t want = tag_hi | (uint32_t)(((js - 1) >> kRingShift) + 1);
                    if ((POLICY & 1) || (uint32_t)(vh[q] >> 32) == want) {
                        Ph[q] = (uint32_t)vh[q];
                        todo &= ~(1u << (16 + q));
                    }
                }
            }
            pending = __any_sync(0xFFFFFFFFu, todo != 0u);
        } while (pending);
        // phase 2: this row group's contributions
        uint32_t Tlo[RPW], Thi[RPW];
        uint32_t tot_lo = 0, tot_hi = 0;
#pragma unroll
        for (int q = 0; q < RPW; q++) {
            const int r = 1 + g * RPW + q;
            Reach64 xa, xb;
            xa.lo = P[q] & 0x55555555u;
            xa.hi = (P[q] >> 1) & 0x55555555u;
            xb.lo = __shfl_up_sync(0xFFFFFFFFu, xa.lo, 1);
            xb.hi = __shfl_up_sync(0xFFFFFFFFu, xa.hi, 1);
            if (lane == 0) {
                xb.lo = Ph[q] & 0x55555555u;
                xb.hi = (Ph[q] >> 1) & 0x55555555u;
            }
            const int sh2 = r < R ? s_sh2[r] : 0;  // warp-uniform
            if (sh2 < 32) {
                Tlo[q] = __funnelshift_r(xa.lo, xa.hi, sh2);
                Thi[q] = __funnelshift_r(xa.hi, xb.lo, sh2);
            } else {
                Tlo[q] = __funnelshift_r(xa.hi, xb.lo, sh2 - 32);
                Thi[q] = __funnelshift_r(xb.lo, xb.hi, sh2 - 32);
            }
            tot_lo |= Tlo[q];
            tot_hi |= Thi[q];
        }
        // phase 3: prefix-OR across the row groups
        s_tot[k & 1][g][lane] = ((uint64_t)tot_hi << 32) | tot_lo;
        bar_sync(1, nd * 32);
        if (threadIdx.x == 0) st_prog(prog + blockIdx.x, (uint32_t)(k + 1));  // every read of this tile is done
        uint64_t carry = (j == 0) ? 0x4000000000000000ULL : 0ULL;  // reach_0 = {0}
        for (int gg = 0; gg < g; gg++) carry |= s_tot[k & 1][gg][lane];
        uint32_t c_lo = (uint32_t)carry, c_hi = (uint32_t)(carry >> 32);
        // phase 4: table words (streaming) and messages
        if (j < C) {
            const uint32_t tag = tag_hi | (uint32_t)((j >> kRingShift) + 1);
            if (g == 0) {
                uint64_t w0 = (j == 0) ? 0xC000000000000000ULL : 0ULL;
                if (j == C - 1) w0 &= last_mask;
                st_cg_u64(tbl + j, w0);
            }
#pragma unroll
            for (int q = 0; q < RPW; q++) {
                const int r = 1 + g * RPW + q;
                if (r < R) {
                    uint64_t out = ((uint64_t)(c_hi | (Thi[q] << 1)) << 32) | (c_lo | (Tlo[q] << 1));
                    if (j == C - 1) out &= last_mask;
                    c_lo |= Tlo[q];
                    c_hi |= Thi[q];
                    st_msg(msg + (size_t)r * kRing + (j & (kRing - 1)), ((uint64_t)tag << 32) | (uint64_t)((c_hi << 1) | c_lo));
                    st_cg_u64(tbl + (int64_t)r * C + j, out);
                }
            }
        }
    }
}

}  // namespace sst
using namespace sst;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_compare(const uint64_t* a, const uint64_t* b, size_t n, unsigned long long* out) {
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
        if (a[k] != b[k]) {
            atomicAdd(&out[0], 1ULL);
            atomicMin(&out[1], (unsigned long long)k);
        }
}

struct Prob {
    uint64_t *A, *B, *msg;
    int R; int64_t C; int32_t *d_step, *d_shift; int* flags; uint32_t* prog; int n_tiles; uint64_t last_mask; int step_min, step_max;
    unsigned long long* d_cmp; uint32_t epoch = 0; int sms;
};

template <int RPW, int POLICY>
void run_ll(Prob& p, int reps, int grid_cap) {
    auto kern = k_build_table_ll<RPW, POLICY>;
    const int nwarps = (p.R - 1 + RPW - 1) / RPW;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nwarps * 32, 0);
    int grid = p.sms * occ;
    const int indep = (p.step_min - 31) / 32;
    if (grid > indep) grid = indep;
    if (grid > grid_cap) grid = grid_cap;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemsetAsync(p.prog, 0, 4096));
        p.epoch++;
        void* args[] = {&p.B, &p.R, &p.C, &p.d_step, &p.d_shift, &p.last_mask, &p.n_tiles, &p.msg, &p.epoch, &p.prog};
        CK(cudaEventRecord(a));
        CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(nwarps * 32), args, 0, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    unsigned long long init[2] = {0ULL, ~0ULL};
    CK(cudaMemcpy(p.d_cmp, init, 16, cudaMemcpyHostToDevice));
    k_compare<<<1184, 256>>>(p.A, p.B, (size_t)p.R * p.C, p.d_cmp);
    unsigned long long h[2];
    CK(cudaMemcpy(h, p.d_cmp, 16, cudaMemcpyDeviceToHost));
    printf("LL rpw=%d policy=%d grid=%d (occ %d): %.3f ms -> %.0f GB/s   mismatches=%llu", RPW, POLICY, grid, occ, best,
           (double)p.R * p.C * 8 / best / 1e6, h[0]);
    if (h[0]) printf(" first at row %llu word %llu", h[1] / p.C, h[1] % p.C);
    printf("\n");
    CK(cudaMemset(p.B, 0xEE, (size_t)p.R * p.C * 8));
}

template <int RPW>
void run_tile(Prob& p, int reps) {
    auto kern = k_build_table<RPW, 0>;
    const int nwarps = (p.R - 1 + RPW - 1) / RPW;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nwarps * 32, 0);
    int grid = p.sms * occ;
    const int indep = (p.step_min - 31) / 32;
    if (grid > indep) grid = indep;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemsetAsync(p.flags, 0, (size_t)p.n_tiles * kBuildMaxWarps * sizeof(int)));
        uint4* no_masks = nullptr;
        void* args[] = {&p.A, &p.R, &p.C, &p.d_step, &p.d_shift, &p.last_mask, &p.n_tiles, &p.flags, &no_masks};
        CK(cudaEventRecord(a));
        CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(nwarps * 32), args, 0, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    printf("tile kernel rpw=%d grid=%d: %.3f ms -> %.0f GB/s\n", RPW, grid, best, (double)p.R * p.C * 8 / best / 1e6);
}

int main(int argc, char** argv) {
    std::vector<long long> w;
    FILE* f = fopen(argc > 1 ? argv[1] : "tools/weights_full.txt", "r");
    if (!f) { printf("no weights file\n"); return 1; }
    long long x; while (fscanf(f, "%lld", &x) == 1) w.push_back(x);
    fclose(f);
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    if (argc > 3) {  // keep the first `rows` rows and the heaviest one (same table width)
        int rows = atoi(argv[3]);
        long long last = w.back();
        w.resize(rows - 1);
        w.push_back(last);
    }
    Prob p;
    p.R = (int)w.size();
    long long max_mass = w.back() * 35;
    p.C = (max_mass + 1 + 31) / 32;
    p.n_tiles = (int)((p.C + 31) / 32);
    p.last_mask = ~0ULL << (2 * (31 - (int)(max_mass % 32)));
    std::vector<int32_t> st(p.R), sh(p.R);
    p.step_min = 1 << 30; p.step_max = 0;
    for (int i = 0; i < p.R; i++) {
        st[i] = (int32_t)(w[i] / 32); sh[i] = (int32_t)(w[i] % 32);
        if (i && st[i] < p.step_min) p.step_min = st[i];
        if (st[i] > p.step_max) p.step_max = st[i];
    }
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    p.sms = prop.multiProcessorCount;
    CK(cudaMalloc(&p.A, (size_t)p.R * p.C * 8)); CK(cudaMalloc(&p.B, (size_t)p.R * p.C * 8));
    CK(cudaMalloc(&p.msg, (size_t)p.R * kRing * 8)); CK(cudaMemset(p.msg, 0, (size_t)p.R * kRing * 8));
    CK(cudaMalloc(&p.d_step, p.R * 4)); CK(cudaMalloc(&p.d_shift, p.R * 4));
    CK(cudaMalloc(&p.flags, (size_t)(p.n_tiles + 1) * kBuildMaxWarps * sizeof(int)));
    CK(cudaMalloc(&p.prog, 4096));
    CK(cudaMalloc(&p.d_cmp, 16));
    CK(cudaMemcpy(p.d_step, st.data(), p.R * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(p.d_shift, sh.data(), p.R * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(p.B, 0xEE, (size_t)p.R * p.C * 8));
    printf("R=%d C=%lld step_min=%d step_max=%d sms=%d ring=%d words/row (%.1f MB)\n", p.R, (long long)p.C, p.step_min, p.step_max, p.sms, kRing,
           (double)p.R * kRing * 8 / 1e6);
    int rpw = 1;
    while ((p.R - 1 + rpw - 1) / rpw > kBuildMaxWarps) rpw *= 2;
    if (rpw == 8) {
        run_tile<8>(p, reps);
        run_ll<8, 0>(p, reps, 1 << 30);
        run_ll<8, 2>(p, reps, 1 << 30);
        run_ll<8, 1>(p, reps, 1 << 30);
        run_ll<8, 0>(p, reps, 148);
        run_ll<8, 0>(p, reps, 222);
    } else if (rpw == 2) {
        run_tile<2>(p, reps);
        run_ll<2, 0>(p, reps, 1 << 30);
    } else if (rpw == 1) {
        run_tile<1>(p, reps);
        run_ll<1, 0>(p, reps, 1 << 30);
    }
    return 0;
}
