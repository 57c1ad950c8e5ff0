// Probe of a ROW-PIPELINE table build (k_build_table_rows<G,U>, defined here) against the product's tile kernel
// (k_build_table<8>): byte-for-byte comparison on the device and best-of-N timings per configuration.
// Measured on B200 (round 1): bit-identical, but 0.86 ms against 0.375 ms - see DESIGN.md section 4, K1.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/k1_rows_probe tools/k1_rows_probe.cu
//   tools/k1_rows_probe tools/weights_full.txt [reps] [rows]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../spectrseqtools_b200/csrc/sst_table.cuh"
namespace sst {

// ---------------- K1 (row pipeline) ----------------
// One CTA per table row, rows chained like a systolic array.  Row i needs, for word j,
//   b0  = reach_{i-1}(j)                       = the reach bits of row i-1's word j   (from the CTA above)
//   T   = funnel(x_i(j-step_i-1), x_i(j-step_i))                                       (its OWN earlier words)
// so the long-distance dependency (>= step_min words back) never leaves the SM: every CTA keeps the reach
// words x_i of its last step_i+1 (+ in-flight) columns in a shared-memory ring, and the only traffic between
// SMs is the neighbour hand-off "row i-1 has stored block k" (release/acquire flag, the consumer re-reads the
// table words from L2).  The 73-generation chain of the tile kernel becomes a (R-1)-stage pipeline fill.
//
// A CTA is split into G groups of TG = 1024/G threads; group g owns blocks g, g+G, ... of B = TG*U words and
// runs  wait -> load -> compute -> store -> group barrier -> publish  on its own, so G blocks are in flight
// per SM and the hand-off latency is hidden after the fill.  Group progress inside a CTA goes through s_done[]
// (a block may only start when the blocks it reads history from, and the blocks whose ring slots it is about to
// overwrite, are complete).  Ring: N words, N a multiple of B and >= step_max + 1 + G*B.
constexpr int kRowThreads = 1024;

__device__ __forceinline__ int ld_acquire_cta_shared(const int* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_cta_shared(int* p, int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}

// One block of one row.  FAST: every source exists, no last column, the ring read window does not wrap.
// rs = ring slot of word (block start - step - 1), ws = ring slot of the block start (a multiple of B).
template <int TG, int U, bool FAST, bool FIRST_ROW>
__device__ __forceinline__ void rows_block(uint64_t* __restrict__ row, const uint64_t* __restrict__ prev, uint64_t* ring,
                                           int N, int64_t C, int64_t jb, int t, int step, int sh2, int rs, int ws,
                                           uint64_t last_mask) {
    uint64_t wp[U];
    uint64_t xa[U], xb[U];
    if (FAST) {
        const uint64_t* pp = prev + jb + t;
        const uint64_t* ra = ring + rs + t;  // [0] = x(j-step-1), [1] = x(j-step)
#pragma unroll
        for (int u = 0; u < U; u++) wp[u] = FIRST_ROW ? 0ULL : ld_cg_u64(pp + TG * u);
#pragma unroll
        for (int u = 0; u < U; u++) {
            xb[u] = ra[TG * u];
            xa[u] = ra[TG * u + 1];
        }
    } else {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t j = jb + t + TG * u;
            wp[u] = 0;
            if (j < C) {
                if (FIRST_ROW) wp[u] = (j == 0) ? 0x4000000000000000ULL : 0ULL;  // reach_0 = {0}
                else wp[u] = ld_cg_u64(prev + j);
            }
            int sb = rs + t + TG * u;
            if (sb >= N) sb -= N;
            int sa = sb + 1;
            if (sa >= N) sa -= N;
            xb[u] = (j - step - 1 >= 0) ? ring[sb] : 0ULL;
            xa[u] = (j - step >= 0) ? ring[sa] : 0ULL;
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const Reach64 b0 = reach_of(wp[u]);
        const uint32_t alo = (uint32_t)xa[u], ahi = (uint32_t)(xa[u] >> 32);
        const uint32_t blo = (uint32_t)xb[u], bhi = (uint32_t)(xb[u] >> 32);
        uint32_t Tlo, Thi;
        if (sh2 < 32) {  // uniform over the CTA
            Tlo = __funnelshift_r(alo, ahi, sh2);
            Thi = __funnelshift_r(ahi, blo, sh2);
        } else {
            Tlo = __funnelshift_r(ahi, blo, sh2 - 32);
            Thi = __funnelshift_r(blo, bhi, sh2 - 32);
        }
        const uint32_t xlo = b0.lo | Tlo, xhi = b0.hi | Thi;
        uint64_t out = ((uint64_t)(b0.hi | (Thi << 1)) << 32) | (b0.lo | (Tlo << 1));
        const uint64_t x = ((uint64_t)xhi << 32) | xlo;
        if (FAST) {
            st_cg_u64(row + jb + t + TG * u, out);
            ring[ws + t + TG * u] = x;
        } else {
            const int64_t j = jb + t + TG * u;
            if (j < C) {
                if (j == C - 1) out &= last_mask;
                st_cg_u64(row + j, out);
                ring[ws + t + TG * u] = x;
            }
        }
    }
}

// POLICY is 0 in the product; tools/k1_rows_probe.cu times (incorrect) variants: 1 = no polling of the row above,
// 2 = no gpu-scope release, 4 = no intra-CTA progress wait.
template <int G, int U, int POLICY = 0>
__global__ void __launch_bounds__(kRowThreads, 1)
k_build_table_rows(uint64_t* __restrict__ tbl, int R, int64_t C, const int32_t* __restrict__ g_step,
                   const int32_t* __restrict__ g_shift, uint64_t last_mask, int N, int* __restrict__ flags) {
    constexpr int TG = kRowThreads / G, B = TG * U;
    extern __shared__ __align__(16) uint64_t s_ring[];
    __shared__ int s_done[G];
    const int i = blockIdx.x, tid = threadIdx.x;
    uint64_t* row = tbl + (int64_t)i * C;
    if (i == 0) {  // row 0: nothing but mass 0
        for (int64_t j = tid; j < C; j += kRowThreads) {
            uint64_t w = (j == 0) ? 0xC000000000000000ULL : 0ULL;
            if (j == C - 1) w &= last_mask;
            st_cg_u64(row + j, w);
        }
        return;
    }
    const int step = g_step[i], sh2 = 2 * g_shift[i];
    for (int k = tid; k < N; k += kRowThreads) s_ring[k] = 0ULL;
    if (tid < G) s_done[tid] = 0;
    __syncthreads();
    const int g = tid / TG, t = tid % TG, lane = tid & 31;
    const uint64_t* prev = row - C;
    const int nblocks = (int)((C + B - 1) / B);
    const int hist_blocks = (step + 1 + B - 1) / B;  // block b reads history from blocks <= b + 1 - hist_blocks
    int* my_flag = flags + i * G + g;
    const int* up_flag = flags + (i - 1) * G + g;
    int ws = (int)(((int64_t)g * B) % N);  // N is a multiple of B, so a block's write slots never wrap
    const int adv = (int)(((int64_t)G * B) % N);
    for (int b = g, n = 0; b < nblocks; b += G, n++) {
        const int64_t jb = (int64_t)b * B;
        // (1) the blocks this one reads history from / whose ring slots it overwrites are complete
        const int L = max(b + 1 - hist_blocks, b - G);
        if (!(POLICY & 4) && L >= 0 && lane < G && lane != g) {
            const int need = (L >= lane) ? (L - lane) / G + 1 : 0;
            while (ld_acquire_cta_shared(&s_done[lane]) < need) {
            }
        }
        // (2) the row above has stored this block
        if (!(POLICY & 1) && i >= 2 && lane == 0) {
            while (ld_acquire(up_flag) < n + 1) {
            }
        }
        __syncwarp();
        int rs = ws - step - 1;  // step + 1 < N
        if (rs < 0) rs += N;
        const bool fast = (jb - step - 1 >= 0) && (jb + B < C) && (rs + B < N);
        if (i == 1) {
            if (fast) rows_block<TG, U, true, true>(row, prev, s_ring, N, C, jb, t, step, sh2, rs, ws, last_mask);
            else rows_block<TG, U, false, true>(row, prev, s_ring, N, C, jb, t, step, sh2, rs, ws, last_mask);
        } else {
            if (fast) rows_block<TG, U, true, false>(row, prev, s_ring, N, C, jb, t, step, sh2, rs, ws, last_mask);
            else rows_block<TG, U, false, false>(row, prev, s_ring, N, C, jb, t, step, sh2, rs, ws, last_mask);
        }
        // (3) publish: ring words to the other groups, table words to the row below
        bar_sync(1 + g, TG);
        if (t == 0) {
            st_release_cta_shared(&s_done[g], n + 1);
            if (POLICY & 2) *(volatile int*)my_flag = n + 1;
            else st_release(my_flag, n + 1);
        }
        ws += adv;
        if (ws >= N) ws -= N;
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
    uint64_t *A, *B;
    int R; int64_t C; int32_t *d_step, *d_shift; int* flags; int n_tiles; uint64_t last_mask; int step_min, step_max;
    unsigned long long* d_cmp;
};

template <int G, int U, int POLICY = 0>
void run_rows(Prob& p, int reps) {
    constexpr int TG = kRowThreads / G, B = TG * U;
    auto kern = k_build_table_rows<G, U, POLICY>;
    int N = p.step_max + 1 + G * B;
    N = (N + B - 1) / B * B;
    const size_t smem = (size_t)N * 8;
    if (smem > 232448 - 256 || p.step_min < B) { printf("G=%d U=%d: N=%d does not fit / step_min too small\n", G, U, N); return; }
    CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < reps; r++) {
        CK(cudaMemsetAsync(p.flags, 0, (size_t)p.R * 16 * sizeof(int)));
        void* args[] = {&p.B, &p.R, &p.C, &p.d_step, &p.d_shift, &p.last_mask, &N, &p.flags};
        CK(cudaEventRecord(a));
        CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(p.R), dim3(kRowThreads), args, smem, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    CK(cudaMemset(p.d_cmp, 0, 16));
    unsigned long long init[2] = {0ULL, ~0ULL};
    CK(cudaMemcpy(p.d_cmp, init, 16, cudaMemcpyHostToDevice));
    k_compare<<<1184, 256>>>(p.A, p.B, (size_t)p.R * p.C, p.d_cmp);
    unsigned long long h[2];
    CK(cudaMemcpy(h, p.d_cmp, 16, cudaMemcpyDeviceToHost));
    printf("rows policy=%d G=%d U=%d B=%d N=%d smem=%zu: %.3f ms -> %.0f GB/s   mismatches=%llu", POLICY, G, U, B, N, smem, best,
           (double)p.R * p.C * 8 / best / 1e6, h[0]);
    if (h[0]) printf(" first at row %llu word %llu", h[1] / p.C, h[1] % p.C);
    printf("\n");
    CK(cudaMemset(p.B, 0xEE, (size_t)p.R * p.C * 8));
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
    const int sms = prop.multiProcessorCount;
    CK(cudaMalloc(&p.A, (size_t)p.R * p.C * 8)); CK(cudaMalloc(&p.B, (size_t)p.R * p.C * 8));
    CK(cudaMalloc(&p.d_step, p.R * 4)); CK(cudaMalloc(&p.d_shift, p.R * 4));
    CK(cudaMalloc(&p.flags, (size_t)(p.n_tiles + 1) * kBuildMaxWarps * sizeof(int)));
    CK(cudaMalloc(&p.d_cmp, 16));
    CK(cudaMemcpy(p.d_step, st.data(), p.R * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(p.d_shift, sh.data(), p.R * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(p.B, 0xEE, (size_t)p.R * p.C * 8));
    printf("R=%d C=%lld step_min=%d step_max=%d sms=%d\n", p.R, (long long)p.C, p.step_min, p.step_max, sms);
    {   // the tile kernel: reference result in A
        int rpw = 1;
        while ((p.R - 1 + rpw - 1) / rpw > kBuildMaxWarps) rpw *= 2;
        const int nwarps = (p.R - 1 + rpw - 1) / rpw;
        auto go = [&](auto kern) {
            int occ = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nwarps * 32, 0);
            int grid = sms * occ;
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
            printf("tile kernel rpw=%d grid=%d: %.3f ms -> %.0f GB/s\n", rpw, grid, best, (double)p.R * p.C * 8 / best / 1e6);
        };
        if (rpw == 1) go(k_build_table<1>); else if (rpw == 2) go(k_build_table<2>); else if (rpw == 4) go(k_build_table<4>); else go(k_build_table<8>);
    }
    run_rows<8, 4>(p, reps);
    run_rows<8, 4, 1>(p, reps);
    run_rows<8, 4, 2>(p, reps);
    run_rows<8, 4, 4>(p, reps);
    run_rows<8, 4, 3>(p, reps);
    run_rows<8, 4, 7>(p, reps);
    run_rows<4, 4>(p, reps);
    run_rows<4, 4, 2>(p, reps);
    run_rows<4, 4, 7>(p, reps);
    return 0;
}
