// Probe: do copies and kernels of different streams overlap on this box, and what do pinned copies cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/overlap_probe tools/overlap_probe.cu && tools/overlap_probe
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
__global__ void spin(unsigned long long ns, unsigned long long* out) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
    if (out && threadIdx.x == 0 && blockIdx.x == 0) *out = t - t0;
}
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t MB = 1 << 20;
    char *h_in[4], *h_out[4], *d_in[4], *d_out[4];
    cudaStream_t st[4];
    for (int i = 0; i < 4; i++) {
        cudaHostAlloc(&h_in[i], 8 * MB, cudaHostAllocDefault);
        cudaHostAlloc(&h_out[i], 8 * MB, cudaHostAllocDefault);
        cudaMalloc(&d_in[i], 8 * MB);
        cudaMalloc(&d_out[i], 8 * MB);
        cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t sz : {(size_t)64 << 10, MB, 2 * MB, 4 * MB, 8 * MB}) {
        for (int dir = 0; dir < 2; dir++) {
            float best = 1e9;
            for (int r = 0; r < 10; r++) {
                cudaEventRecord(e0, st[0]);
                if (dir == 0) cudaMemcpyAsync(d_in[0], h_in[0], sz, cudaMemcpyHostToDevice, st[0]);
                else cudaMemcpyAsync(h_out[0], d_out[0], sz, cudaMemcpyDeviceToHost, st[0]);
                cudaEventRecord(e1, st[0]);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("%s %7zu KB: %7.1f us  %6.1f GB/s\n", dir ? "D2H" : "H2D", sz >> 10, best * 1e3, sz / (best * 1e-3) / 1e9);
        }
    }
    // launch cost seen by a pair of events around ONE kernel of 296 x 512 threads that spins for 50 us: plain vs cooperative
    for (int coop = 0; coop < 2; coop++) {
        float best = 1e9, sum = 0;
        for (int r = 0; r < 20; r++) {
            cudaDeviceSynchronize();
            unsigned long long ns = 50000; unsigned long long* outp = nullptr;
            cudaEventRecord(e0, st[0]);
            if (coop) {
                void* args[] = {&ns, &outp};
                cudaLaunchCooperativeKernel((const void*)spin, dim3(296), dim3(512), args, 0, st[0]);
            } else {
                spin<<<296, 512, 0, st[0]>>>(ns, outp);
            }
            cudaEventRecord(e1, st[0]);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
            if (r >= 4) sum += ms;
        }
        printf("%s launch of a 50 us kernel, event to event: best %.1f us, mean %.1f us\n", coop ? "cooperative" : "plain", best * 1e3, sum / 16 * 1e3);
    }
    // pipeline: per step H2D 2 MB -> kernel 100 us -> D2H 5 MB, on `depth` streams round robin, 200 steps
    for (int coop = 0; coop < 2; coop++)
    for (int depth = 1; depth <= 4; depth++) {
        cudaDeviceSynchronize();
        const int n = 200;
        double t0 = now_us();
        for (int i = 0; i < n; i++) {
            const int s = i % depth;
            if (i >= depth) cudaStreamSynchronize(st[s]);
            cudaMemcpyAsync(d_in[s], h_in[s], 2 * MB, cudaMemcpyHostToDevice, st[s]);
            unsigned long long ns = 100000; unsigned long long* outp = nullptr;
            if (coop) {
                void* args[] = {&ns, &outp};
                cudaLaunchCooperativeKernel((const void*)spin, dim3(296), dim3(512), args, 0, st[s]);
            } else {
                spin<<<296, 512, 0, st[s]>>>(ns, outp);
            }
            cudaMemcpyAsync(h_out[s], d_out[s], 5 * MB, cudaMemcpyDeviceToHost, st[s]);
        }
        cudaDeviceSynchronize();
        double t1 = now_us();
        printf("%s depth %d: %.1f us per step (H2D 2 MB + 100 us kernel + D2H 5 MB)\n", coop ? "cooperative" : "plain", depth, (t1 - t0) / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
