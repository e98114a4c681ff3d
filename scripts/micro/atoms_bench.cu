// Development aid: throughput of shared-memory atomics (ATOMS), global reductions (RED) and a mix of both, in lanes per
// clock per SM, for histogram-like access (pseudo-random 32-bit bins).  Build: nvcc -arch=sm_100a -O3 -o atoms_bench atoms_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kThreads = 512;
constexpr int kBins = 4096;

template <int MODE>   // 0 smem no return, 1 smem with return, 2 global RED, 3 half smem / half global, 4 = 3 of 4 smem, 5 = LDS+STS (no atomics)
__global__ void __launch_bounds__(kThreads) k(unsigned int *g, int iters, unsigned int *sink) {
    __shared__ unsigned int h[kBins];
    for (int i = threadIdx.x; i < kBins; i += kThreads) h[i] = 0;
    __syncthreads();
    unsigned int *gh = g + (size_t)blockIdx.x * kBins;
    unsigned int x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u, acc = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            x = x * 1664525u + 1013904223u;
            const unsigned int b = (x >> 12) & (kBins - 1);
            if (MODE == 0) atomicAdd(&h[b], 1u);
            else if (MODE == 1) acc += atomicAdd(&h[b], 1u);
            else if (MODE == 2) atomicAdd(&gh[b], 1u);
            else if (MODE == 3) { if (e & 1) atomicAdd(&gh[b], 1u); else atomicAdd(&h[b], 1u); }
            else if (MODE == 4) { if (e == 3) atomicAdd(&gh[b], 1u); else atomicAdd(&h[b], 1u); }
            else { acc += h[b]; h[(b + threadIdx.x) & (kBins - 1)] = acc; }
        }
    }
    __syncthreads();
    if (MODE == 1 || MODE == 5) { if (acc == 0xdeadbeefu) sink[0] = acc; }
    for (int i = threadIdx.x; i < kBins; i += kThreads) if (h[i] == 0xdeadbeefu) sink[1] = h[i];
}

template <int MODE>
void run(const char *name, int sms, int ctas_per_sm, unsigned int *g, unsigned int *sink, double ghz) {
    const int iters = 2000, grid = sms * ctas_per_sm;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<grid, kThreads>>>(g, 10, sink);
    cudaEventRecord(a);
    k<MODE><<<grid, kThreads>>>(g, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)grid * kThreads * iters * 4;
    printf("%-34s %d CTA/SM  %.3f ms  %.3f lanes/clk/SM (at %.2f GHz)\n", name, ctas_per_sm, ms, lanes / sms / (ms * 1e-3 * ghz * 1e9), ghz);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    unsigned int *g, *sink;
    cudaMalloc(&g, (size_t)p.multiProcessorCount * 4 * kBins * 4); cudaMemset(g, 0, (size_t)p.multiProcessorCount * 4 * kBins * 4);
    cudaMalloc(&sink, 64);
    for (int c = 2; c <= 4; c += 2) {
        run<0>("smem atomicAdd, no return", p.multiProcessorCount, c, g, sink, ghz);
        run<1>("smem atomicAdd, with return", p.multiProcessorCount, c, g, sink, ghz);
        run<2>("global RED (L2-resident, per CTA)", p.multiProcessorCount, c, g, sink, ghz);
        run<3>("half smem / half global", p.multiProcessorCount, c, g, sink, ghz);
        run<4>("3 of 4 smem / 1 of 4 global", p.multiProcessorCount, c, g, sink, ghz);
        run<5>("LDS + STS (no atomics)", p.multiProcessorCount, c, g, sink, ghz);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
