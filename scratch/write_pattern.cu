// Write-bandwidth ceiling of the z sweep's store pattern on this box: every warp walks z downwards writing one 128-byte line
// per array per step (stride = one slice), against the same volume written in address order.  usage: write_pattern [X Y Z]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k_zwalk(int *a, int *b, int X, int Y, int Z, int zchunk)
{
    // item = (z chunk, y, x group); chunk-major order
    const int lane = threadIdx.x & 31;
    const int XG = X / 32;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5), gw = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nchunks = Z / zchunk;
    const long long items = (long long)nchunks * Y * XG;
    const size_t slice = (size_t)X * Y;
    for (long long it = gw; it < items; it += nwarps) {
        const int c = (int)(it / ((long long)Y * XG)), rem = (int)(it % ((long long)Y * XG));
        const int y = rem / XG, x = (rem % XG) * 32 + lane;
        const int z_hi = Z - 1 - c * zchunk;
        size_t o = (size_t)z_hi * slice + (size_t)y * X + x;
        for (int u = 0; u < zchunk; u++) { __stcs(a + o, u); __stcs(b + o, u + 1); o -= slice; }
    }
}
// The same stores issued the way the z sweep issues them: warps pull (row, x group) items from a counter; optionally a forward
// phase first (NS slices of two input arrays, 8 loads per dependent batch) and a shared-memory read every POPEVERY steps.
template <int FWD, int POPEVERY>
__global__ void __launch_bounds__(256, 4) k_zwalk_dyn(int *a, int *b, const int *g, const int *c, int X, int Y, int Z, int ns, int *counter)
{
    __shared__ int ring[8 * 32 * 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int XG = X / 32, n_items = Y * XG;
    const size_t slice = (size_t)X * Y;
    int *my = ring + wid * 32 * 32 + lane;
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        const int y = item / XG, x = (item % XG) * 32 + lane;
        const size_t base = (size_t)y * X + x;
        int acc = 0;
        if (FWD) {
            for (int j0 = 0; j0 < ns; j0 += 8) {
                int hh[8], cc[8];
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int kk = min(j0 + k, ns - 1) * (Z / ns);
                    hh[k] = __ldcs(&g[base + (size_t)kk * slice]);
                    cc[k] = __ldcs(&c[base + (size_t)kk * slice]);
                }
#pragma unroll
                for (int k = 0; k < 8; k++) { acc += hh[k] ^ cc[k]; my[((j0 + k) & 31) * 32] = acc; }
            }
        }
        size_t o = base + (size_t)(Z - 1) * slice;
        int *pa = a + o, *pb = b + o;
        int h = acc, t = Z - 3;
        for (int v = Z - 1; v >= 0; v--) {
            const int d = v - 7;
            __stcs(pa, d * d + h); __stcs(pb, h);
            if (POPEVERY && v == t) { h = my[(v & 31) * 32]; t = v - POPEVERY - (h & 1); }
            if (POPEVERY) __syncwarp();
            pa -= slice; pb -= slice;
        }
    }
}
__global__ void k_linear(int *a, int *b, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { __stcs(a + i, 1); __stcs(b + i, 2); }
}
int main(int argc, char **argv)
{
    int X = argc > 1 ? atoi(argv[1]) : 512, Y = argc > 2 ? atoi(argv[2]) : 512, Z = argc > 3 ? atoi(argv[3]) : 512;
    size_t n = (size_t)X * Y * Z;
    int *a, *b;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, auto launch) {
        for (int i = 0; i < 3; i++) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; i++) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
        printf("%-44s %.3f ms  %.0f GB/s\n", name, ms, n * 8 / ms / 1e6);
    };
    timeit("address order (grid-stride, 4 B stores)", [&] { k_linear<<<148 * 16, 256>>>(a, b, n); });
    for (int zc : { 512, 128, 32, 8 }) {
        if (zc > Z) continue;
        char name[96];
        snprintf(name, sizeof name, "z walk, chunk %d, 148x4 CTAs x 8 warps", zc);
        timeit(name, [&] { k_zwalk<<<148 * 4, 256>>>(a, b, X, Y, Z, zc); });
        snprintf(name, sizeof name, "z walk, chunk %d, 148x8 CTAs x 8 warps", zc);
        timeit(name, [&] { k_zwalk<<<148 * 8, 256>>>(a, b, X, Y, Z, zc); });
    }
    int *g, *c, *counter;
    cudaMalloc(&g, n * 4); cudaMalloc(&c, n * 4); cudaMalloc(&counter, 4);
    cudaMemset(g, 0, n * 4); cudaMemset(c, 0, n * 4);
    const int ns = 55;
    for (int mult : { 4, 8 }) {
        char name[96];
        snprintf(name, sizeof name, "dynamic items, stores only, 148x%d CTAs", mult);
        timeit(name, [&] { cudaMemsetAsync(counter, 0, 4); k_zwalk_dyn<0, 0><<<148 * mult, 256>>>(a, b, g, c, X, Y, Z, ns, counter); });
        snprintf(name, sizeof name, "dynamic + forward loads (55 slices), 148x%d", mult);
        timeit(name, [&] { cudaMemsetAsync(counter, 0, 4); k_zwalk_dyn<1, 0><<<148 * mult, 256>>>(a, b, g, c, X, Y, Z, ns, counter); });
        snprintf(name, sizeof name, "dynamic + forward + pop every ~3 steps, 148x%d", mult);
        timeit(name, [&] { cudaMemsetAsync(counter, 0, 4); k_zwalk_dyn<1, 3><<<148 * mult, 256>>>(a, b, g, c, X, Y, Z, ns, counter); });
        snprintf(name, sizeof name, "dynamic + pop every ~3 steps (no forward), 148x%d", mult);
        timeit(name, [&] { cudaMemsetAsync(counter, 0, 4); k_zwalk_dyn<0, 3><<<148 * mult, 256>>>(a, b, g, c, X, Y, Z, ns, counter); });
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
