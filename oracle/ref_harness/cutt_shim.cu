// cutt_shim.cu — stand-in for the cuTT binary blob (reference lib/libcutt_x86.a holds only an sm_75 cubin and cannot
// load on a B200).  A tensor permutation has no arithmetic, so a plain gather kernel is semantically exact:
// out[index in permuted order] = in[index].  API: reference include/cutt/cutt.h:57-101.  Test infrastructure only.
#include <cutt/cutt.h>
#include <vector>
struct ShimPlan { int rank; int dim[3]; int perm[3]; size_t elem; cudaStream_t stream; };
static std::vector<ShimPlan> g_plans;

__global__ void permute_kernel(const int *in, int *out, int d0, int d1, int d2, int p0, int p1, int p2)
{
    size_t n = (size_t)d0 * d1 * d2;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int idx[3] = { (int)(i % d0), (int)((i / d0) % d1), (int)(i / ((size_t)d0 * d1)) };
    int dims[3] = { d0, d1, d2 };
    // output dimension j has extent dims[perm[j]] and index idx[perm[j]]; first dimension fastest
    size_t o = (size_t)idx[p0] + (size_t)dims[p0] * ((size_t)idx[p1] + (size_t)dims[p1] * (size_t)idx[p2]);
    out[o] = in[i];
}
cuttResult cuttPlan(cuttHandle *handle, int rank, int *dim, int *permutation, size_t sizeofType, cudaStream_t stream)
{
    if (rank < 2 || rank > 3 || sizeofType != 4) return CUTT_INVALID_PARAMETER;
    ShimPlan p{};
    p.rank = rank; p.elem = sizeofType; p.stream = stream;
    for (int i = 0; i < 3; i++) { p.dim[i] = i < rank ? dim[i] : 1; p.perm[i] = i < rank ? permutation[i] : i; }
    g_plans.push_back(p);
    *handle = (cuttHandle)g_plans.size();
    return CUTT_SUCCESS;
}
cuttResult cuttPlanMeasure(cuttHandle *handle, int rank, int *dim, int *permutation, size_t sizeofType, cudaStream_t stream, void *, void *)
{
    return cuttPlan(handle, rank, dim, permutation, sizeofType, stream);
}
cuttResult cuttDestroy(cuttHandle) { return CUTT_SUCCESS; }
cuttResult cuttExecute(cuttHandle handle, void *idata, void *odata)
{
    if (handle == 0 || handle > g_plans.size()) return CUTT_INVALID_PLAN;
    const ShimPlan &p = g_plans[handle - 1];
    size_t n = (size_t)p.dim[0] * p.dim[1] * p.dim[2];
    permute_kernel<<<(unsigned)((n + 255) / 256), 256, 0, p.stream>>>((const int *)idata, (int *)odata, p.dim[0], p.dim[1], p.dim[2],
                                                                      p.perm[0], p.perm[1], p.perm[2]);
    return cudaGetLastError() == cudaSuccess ? CUTT_SUCCESS : CUTT_INTERNAL_ERROR;
}
