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
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
