// Stand-in for include/cutt/cutt.h.  The reference builds three rank-3 permutation plans
// (src/volumetric_mapper.cpp:344-373) and executes them six times per frame (src/kernel/edt/local_edt.cu:12-25).
// The engine's sweep kernels transpose in shared memory, so a plan is only a handle that remembers its arguments.
#pragma once
#include <cstddef>
typedef unsigned int cuttHandle;
typedef enum cuttResult_t { CUTT_SUCCESS, CUTT_INVALID_PLAN, CUTT_INVALID_PARAMETER, CUTT_INVALID_DEVICE,
                            CUTT_INTERNAL_ERROR, CUTT_UNDEFINED_ERROR } cuttResult;
inline cuttResult cuttPlan(cuttHandle *handle, int rank, int *dim, int *permutation, size_t /*sizeofType*/, void * /*stream*/)
{
    if (!handle || rank < 1 || !dim || !permutation) return CUTT_INVALID_PARAMETER;
    static cuttHandle next = 1;
    *handle = next++;
    return CUTT_SUCCESS;
}
inline cuttResult cuttPlanMeasure(cuttHandle *handle, int rank, int *dim, int *permutation, size_t s, void *stream, void *, void *)
{
    return cuttPlan(handle, rank, dim, permutation, s, stream);
}
inline cuttResult cuttDestroy(cuttHandle) { return CUTT_SUCCESS; }
#define cuttCheck(stmt) do { cuttResult e_ = (stmt); (void)e_; } while (0)
