// gmg_common.cuh -- level geometry and the reference arithmetic shared by every GMG kernel file.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mgb {

struct LevelGeom {
    int w;        // points per side of this level (global)          domain.hpp:49 width
    int rows;     // rows owned by this rank
    int row0;     // global index of the first owned row
    int pitch;    // doubles per stored row (multiple of 16)
    double diag;  // 4*alpha/k, k = (h*2^l)^2                        linear_system.hpp:27-28
    double off;   // -alpha/k                                        linear_system.hpp:38
};

// ---- reference arithmetic, never contracted ------------------------------------------------
// solvers.hpp:33-48 / 64-83: sum = off*up + off*left + off*right + off*down (that order, from 0);
// u = (b - sum) / diag
__device__ __forceinline__ double smooth_point(double b, double up, double left, double right,
                                               double down, double off, double diag)
{
    double sum = __dmul_rn(off, up);
    sum = __dadd_rn(sum, __dmul_rn(off, left));
    sum = __dadd_rn(sum, __dmul_rn(off, right));
    sum = __dadd_rn(sum, __dmul_rn(off, down));
    return __ddiv_rn(__dsub_rn(b, sum), diag);
}
// solvers.hpp:257-296: sum over up,left,centre,right,down; r = b - sum
__device__ __forceinline__ double resid_point(double b, double up, double left, double c,
                                              double right, double down, double off, double diag)
{
    double sum = __dmul_rn(off, up);
    sum = __dadd_rn(sum, __dmul_rn(off, left));
    sum = __dadd_rn(sum, __dmul_rn(diag, c));
    sum = __dadd_rn(sum, __dmul_rn(off, right));
    sum = __dadd_rn(sum, __dmul_rn(off, down));
    return __dsub_rn(b, sum);
}

__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over the CTA, result valid in thread 0.  `red` holds >= blockDim.x/32 doubles.
__device__ __forceinline__ double block_sum(double v, double *red)
{
    v = warp_sum(v);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double t = 0.;
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        t = lane < nw ? red[lane] : 0.;
        t = warp_sum(t);
    }
    __syncthreads();
    return t;
}

}  // namespace mgb
