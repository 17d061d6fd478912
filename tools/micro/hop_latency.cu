// Microbenchmark: cost of one hop of the forest walk = one ATOMG (64-bit, with return) on a far-away word, with and
// without the dependent 8-byte record load, L2-cold and L2-warm.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void chain(unsigned long long *wq, const int2 *pe, int start, int hops, int mode, long long *cycles, int *sink)
{
    if (threadIdx.x != 0) return;
    int p = start;
    unsigned long long acc = 0;
    long long t0 = clock64();
    for (int h = 0; h < hops; h++) {
        int2 r = make_int2(0, 0);
        unsigned long long old = 0;
        if (mode & 1) old = atomicAdd(&wq[p], 1ull);          // ATOMG with return
        if (mode & 2) r = pe[p];                              // the record: next node
        else r.x = p + 40960;                                 // same stride without the load
        acc += old;
        p = r.x + (int)(old >> 62);                           // both results feed the next address
    }
    long long t1 = clock64();
    *cycles = t1 - t0;
    *sink = p + (int)acc;
}

__global__ void fill(int2 *pe, unsigned long long *wq, int n, int stride)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { pe[i] = make_int2((i + stride) % n, 0); wq[i] = 0; }
}
__global__ void trash(float *b, size_t n) { size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; if (i < n) b[i] = 1.0f; }
__global__ void touch(unsigned long long *wq, const int2 *pe, int start, int hops, int stride, int n, int *sink)
{
    int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < hops) { int p = (int)(((long long)start + (long long)h * stride) % n); atomicAdd(&wq[p], 0ull); atomicAdd(sink, pe[p].x & 1); }
}

int main()
{
    const int n = 10485760, stride = 40960, hops = 250;      // 84 MB arrays, 327 KB between hops (one tile row of 64 tiles)
    int2 *pe; unsigned long long *wq; long long *cyc; int *sink; float *big;
    cudaMalloc(&pe, sizeof(int2) * n); cudaMalloc(&wq, 8ull * n); cudaMalloc(&cyc, 8); cudaMalloc(&sink, 4);
    const size_t nb = 128ull << 20; cudaMalloc(&big, nb * 4);
    fill<<<(n + 255) / 256, 256>>>(pe, wq, n, stride);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char *names[4] = {"", "ATOMG only", "LDG.64 only", "ATOMG + LDG (walk hop)"};
    for (int warm = 0; warm < 2; warm++)
        for (int mode = 1; mode <= 3; mode++) {
            long long best = 1ll << 62;
            for (int rep = 0; rep < 5; rep++) {
                trash<<<(unsigned)((nb + 255) / 256), 256>>>(big, nb);        // 512 MB through L2
                if (warm) touch<<<1, 256>>>(wq, pe, 7 + rep, hops, stride, n, sink);
                chain<<<1, 32>>>(wq, pe, 7 + rep, hops, mode, cyc, sink);
                long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
                if (c < best) best = c;
            }
            printf("%-26s %s: %6.0f cycles/hop = %.3f us/hop (SM clock %d MHz)\n", names[mode], warm ? "L2-warm" : "L2-cold",
                   (double)best / hops, (double)best / hops / (clk * 1e-3), clk / 1000);
        }
    return 0;
}
