// Micro-benchmark of the table-build kernel and of (deliberately incorrect) variants that drop one
// synchronisation mechanism each, to see where the per-generation latency goes.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/k1_probe tools/k1_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../spectrseqtools_b200/csrc/sst_table.cuh"
using namespace sst;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int RPW, int POLICY>
float run(uint64_t* tbl, int R, int64_t C, int32_t* d_step, int32_t* d_shift, int n_tiles, int* flags, int grid_cap, int reps, int sms) {
    auto kern = k_build_table<RPW, POLICY>;
    int occ = 1;
    const int nwarps = (R - 1 + RPW - 1) / RPW;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nwarps * 32, 0);
    int grid = sms * occ;
    if (grid > grid_cap) grid = grid_cap;
    uint64_t last_mask = ~0ULL << 24;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int i = 0; i < reps; i++) {
        CK(cudaMemset(flags, 0, (size_t)n_tiles * kBuildMaxWarps * sizeof(int)));
        uint4* no_masks = nullptr;
        void* args[] = {&tbl, &R, &C, &d_step, &d_shift, &last_mask, &n_tiles, &flags, &no_masks};
        CK(cudaEventRecord(a));
        CK(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(nwarps * 32), args, 0, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    printf("RPW=%d policy=%d grid=%d (occ %d): %.3f ms  -> %.0f GB/s\n", RPW, POLICY, grid, occ, best, (double)R * C * 8 / best / 1e6);
    return best;
}

int main(int argc, char** argv) {
    // full alphabet weights are read from a text file (one integer per line) written by the caller
    std::vector<long long> w;
    FILE* f = fopen(argc > 1 ? argv[1] : "gpurun_out/weights.txt", "r");
    if (!f) { printf("no weights file\n"); return 1; }
    long long x; while (fscanf(f, "%lld", &x) == 1) w.push_back(x);
    fclose(f);
    int R = (int)w.size();
    long long max_mass = w.back() * 35;
    int64_t C = (max_mass + 1 + 31) / 32;
    int n_tiles = (int)((C + 31) / 32);
    std::vector<int32_t> st(R), sh(R);
    long long step_min = 1LL << 60;
    for (int i = 0; i < R; i++) { st[i] = (int32_t)(w[i] / 32); sh[i] = (int32_t)(w[i] % 32); if (i && st[i] < step_min) step_min = st[i]; }
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    uint64_t* tbl; int32_t *d_step, *d_shift; int* flags;
    CK(cudaMalloc(&tbl, (size_t)R * C * 8)); CK(cudaMalloc(&d_step, R * 4)); CK(cudaMalloc(&d_shift, R * 4));
    CK(cudaMalloc(&flags, (size_t)n_tiles * kBuildMaxWarps * sizeof(int)));
    CK(cudaMemcpy(d_step, st.data(), R * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_shift, sh.data(), R * 4, cudaMemcpyHostToDevice));
    int indep = (int)((step_min - 31) / 32);
    printf("R=%d C=%lld tiles=%d step_min=%lld indep=%d sms=%d\n", R, (long long)C, n_tiles, step_min, indep, sms);
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    if (argc > 3) {  // profile mode: only the product variant
        run<8, 0>(tbl, R, C, d_step, d_shift, n_tiles, flags, indep, reps, sms);
        return 0;
    }
    run<8, 0>(tbl, R, C, d_step, d_shift, n_tiles, flags, indep, reps, sms);
    run<8, 2>(tbl, R, C, d_step, d_shift, n_tiles, flags, indep, reps, sms);   // no release fence
    run<8, 1>(tbl, R, C, d_step, d_shift, n_tiles, flags, indep, reps, sms);   // no polling
    run<8, 3>(tbl, R, C, d_step, d_shift, n_tiles, flags, indep, reps, sms);   // neither
    run<8, 0>(tbl, R, C, d_step, d_shift, n_tiles, flags, 148, reps, sms);     // one CTA per SM
    run<8, 3>(tbl, R, C, d_step, d_shift, n_tiles, flags, 100000, reps, sms);  // no sync, all resident CTAs
    // plain streaming write of the same bytes for reference
    {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        float best = 1e9;
        for (int i = 0; i < 5; i++) { cudaEventRecord(a); cudaMemsetAsync(tbl, 1, (size_t)R * C * 8); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
        printf("cudaMemset of the table: %.3f ms -> %.0f GB/s\n", best, (double)R * C * 8 / best / 1e6);
    }
    return 0;
}
