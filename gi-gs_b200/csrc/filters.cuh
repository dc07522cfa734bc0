// 3x3 median network shared by screen.cu (standalone filters, geometry chain) and deferred.cu (fused frame passes).
#pragma once
#include "common.cuh"

namespace gigs {

__device__ __forceinline__ void cswap(float& a, float& b)
{
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo;
    b = hi;
}
__device__ __forceinline__ float median9(float v[9])
{
    bool bad = false;
#pragma unroll
    for (int i = 0; i < 9; ++i) bad |= !isfinite(v[i]);
    if (bad) return __int_as_float(0x7fc00000);
    // 19-exchange median-of-9 network
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[1]); cswap(v[3], v[4]); cswap(v[6], v[7]);
    cswap(v[1], v[2]); cswap(v[4], v[5]); cswap(v[7], v[8]);
    cswap(v[0], v[3]); cswap(v[5], v[8]); cswap(v[4], v[7]);
    cswap(v[3], v[6]); cswap(v[1], v[4]); cswap(v[2], v[5]);
    cswap(v[4], v[7]); cswap(v[4], v[2]); cswap(v[6], v[4]);
    cswap(v[4], v[2]);
    return v[4];
}

// window index (row-major, 0..8) of the first element equal to the median m, or -1 (NaN median)
__device__ __forceinline__ int median9_select(const float v[9], const float m)
{
    int sel = -1;
#pragma unroll
    for (int k = 8; k >= 0; --k)
        if (v[k] == m) sel = k;
    return sel;
}

}  // namespace gigs
