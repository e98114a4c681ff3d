// Micro-benchmark (B200): cost per value of three ways to bin 32-bit keys inside one CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/hist_micro scripts/micro/hist_micro.cu && /tmp/hist_micro
#include <cstdio>
#include <cuda_runtime.h>
constexpr int T = 512, B = 1024, PER = 16, REP = 64;

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void k_atoms(unsigned *out, long long *clk) {
    __shared__ unsigned h[B];
    for (int i = threadIdx.x; i < B; i += T) h[i] = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < REP; ++r)
#pragma unroll
        for (int e = 0; e < PER; ++e) atomicAdd(&h[hash(threadIdx.x * 977 + e * 131 + r * 7 + blockIdx.x) & (B - 1)], 1u);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * T + threadIdx.x] = h[threadIdx.x];
}
__global__ void k_atoms_ret(unsigned *out, long long *clk) {
    __shared__ unsigned h[B];
    for (int i = threadIdx.x; i < B; i += T) h[i] = 0;
    __syncthreads();
    unsigned acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < REP; ++r)
#pragma unroll
        for (int e = 0; e < PER; ++e) acc += atomicAdd(&h[hash(threadIdx.x * 977 + e * 131 + r * 7 + blockIdx.x) & (B - 1)], 1u);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * T + threadIdx.x] = h[threadIdx.x] + acc;
}
// warp-private histogram (u16), intra-warp conflicts resolved with match_any
__global__ void k_match(unsigned *out, long long *clk) {
    __shared__ unsigned short h[T / 32][B];
    for (int i = threadIdx.x; i < (T / 32) * B; i += T) (&h[0][0])[i] = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < REP; ++r)
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const unsigned b = hash(threadIdx.x * 977 + e * 131 + r * 7 + blockIdx.x) & (B - 1);
            const unsigned m = __match_any_sync(0xffffffffu, b);
            const int leader = __ffs(m) - 1;
            unsigned old = 0;
            if (lane == leader) { old = h[w][b]; h[w][b] = (unsigned short)(old + __popc(m)); }
            old = __shfl_sync(0xffffffffu, old, leader);
            acc += old + __popc(m & ((1u << lane) - 1u));
            __syncwarp();
        }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * T + threadIdx.x] = h[w][lane] + acc;
}
// thread-private counters over 32 coarse bins (u16, [bin][thread] layout: conflict free)
__global__ void k_private(unsigned *out, long long *clk) {
    __shared__ unsigned short h[32][T];
    for (int i = threadIdx.x; i < 32 * T; i += T) (&h[0][0])[i] = 0;
    __syncthreads();
    unsigned acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < REP; ++r)
#pragma unroll
        for (int e = 0; e < PER; ++e) {
            const unsigned b = hash(threadIdx.x * 977 + e * 131 + r * 7 + blockIdx.x) & 31;
            const unsigned old = h[b][threadIdx.x];
            h[b][threadIdx.x] = (unsigned short)(old + 1);
            acc += old;
        }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * T + threadIdx.x] = h[threadIdx.x & 31][threadIdx.x] + acc;
}
// hash only (baseline)
__global__ void k_base(unsigned *out, long long *clk) {
    unsigned acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < REP; ++r)
#pragma unroll
        for (int e = 0; e < PER; ++e) acc += hash(threadIdx.x * 977 + e * 131 + r * 7 + blockIdx.x) & (B - 1);
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
    out[blockIdx.x * T + threadIdx.x] = acc;
}
template <class K>
void run(const char *name, K k, int ctas_per_sm) {
    int sms = 148;
    unsigned *out; long long *clk;
    cudaMalloc(&out, sizeof(unsigned) * T * sms * ctas_per_sm);
    cudaMalloc(&clk, sizeof(long long) * sms * ctas_per_sm);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<sms * ctas_per_sm, T>>>(out, clk);
    cudaEventRecord(a);
    k<<<sms * ctas_per_sm, T>>>(out, clk);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[8]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
    const double values_per_sm = (double)T * PER * REP * ctas_per_sm;
    printf("%-12s ctas/sm=%d  %8.1f us  cta cycles=%lld  -> %.3f SM-cycles per value (at 1.965 GHz, from wall time)  err=%s\n", name, ctas_per_sm,
           ms * 1e3, h[0], ms * 1e-3 * 1.965e9 / values_per_sm, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(clk);
}
int main() {
    for (int c = 1; c <= 2; ++c) {
        run("base", k_base, c);
        run("atoms", k_atoms, c);
        run("atoms_ret", k_atoms_ret, c);
        run("match_any", k_match, c);
        run("private", k_private, c);
    }
    return 0;
}
