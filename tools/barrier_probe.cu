// Micro-benchmark of grid-barrier variants for the cooperative enumeration pass (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/barrier_probe tools/barrier_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// V0: the product's barrier.  V1: red.release + relaxed polling without sleep + one fence.  V2: red.release + acquire polling.
// V3: like V1, arrival spread over 8 counters (one per CTA group), thread 0..7 poll them in parallel.
template <int V>
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& gen) {
    __syncthreads();
    gen++;
    if (V == 3) {
        const unsigned want = gen * ((gridDim.x + 7) / 8);
        if (threadIdx.x == 0) red_release(bar + 32 * (blockIdx.x & 7), 1u);
        if (threadIdx.x < 8) {
            const unsigned members = (gridDim.x - threadIdx.x + 7) / 8;  // CTAs with blockIdx % 8 == threadIdx.x
            while ((int)(ld_relaxed(bar + 32 * threadIdx.x) - gen * members) < 0) {}
            __threadfence();
        }
        (void)want;
    } else if (threadIdx.x == 0) {
        const unsigned want = gen * gridDim.x;
        if (V == 0) {
            __threadfence();
            atomicAdd(bar, 1u);
            while ((int)(ld_relaxed(bar) - want) < 0) __nanosleep(20);
            __threadfence();
        } else if (V == 1) {
            red_release(bar, 1u);
            while ((int)(ld_relaxed(bar) - want) < 0) {}
            __threadfence();
        } else {
            red_release(bar, 1u);
            while ((int)(ld_acquire(bar) - want) < 0) {}
        }
    }
    __syncthreads();
}

template <int V>
__global__ void __launch_bounds__(1024) k_bar(unsigned* bar, unsigned* data, int iters, int work) {
    unsigned gen = 0;
    unsigned acc = 0;
    for (int it = 0; it < iters; it++) {
        // a little phase work: every thread writes one word and reads a word another CTA wrote in the previous phase
        const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
        if (work) {
            data[(it & 1) * gridDim.x * blockDim.x + idx] = idx + it;
        }
        grid_barrier<V>(bar, gen);
        if (work) {
            const unsigned other = (idx + 7919u * blockDim.x) % (gridDim.x * blockDim.x);
            acc += __ldcg(data + (it & 1) * gridDim.x * blockDim.x + other);
        }
    }
    if (acc == 0xFFFFFFFFu) data[0] = acc;
}

template <int V>
void run(unsigned* bar, unsigned* data, int grid, int threads, int work) {
    const int iters = 200;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 5; r++) {
        CK(cudaMemset(bar, 0, 4096));
        int it = iters;
        void* args[] = {&bar, &data, &it, &work};
        CK(cudaEventRecord(a));
        CK(cudaLaunchCooperativeKernel((const void*)k_bar<V>, dim3(grid), dim3(threads), args, 0, 0));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    printf("variant %d grid %d x %d work %d: %.2f us per barrier\n", V, grid, threads, work, best * 1e3 / iters);
}

int main() {
    unsigned *bar, *data;
    CK(cudaMalloc(&bar, 4096));
    CK(cudaMalloc(&data, 2 * 296 * 1024 * 4));
    for (int work = 0; work < 2; work++) {
        run<0>(bar, data, 296, 512, work); run<1>(bar, data, 296, 512, work); run<2>(bar, data, 296, 512, work); run<3>(bar, data, 296, 512, work);
        run<0>(bar, data, 148, 1024, work); run<1>(bar, data, 148, 1024, work); run<2>(bar, data, 148, 1024, work); run<3>(bar, data, 148, 1024, work);
        run<0>(bar, data, 148, 512, work); run<1>(bar, data, 148, 512, work);
    }
    return 0;
}
