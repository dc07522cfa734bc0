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

// pixel -> view-space point at `depth` (reference depth_to_normal: forward.cu back-projection of the pixel grid)
__device__ __forceinline__ float3 back_project(int x, int y, float cx, float cy, float fx, float fy, float depth)
{
    const float3 dir = make_float3((float(x) - cx) / fx, (float(y) - cy) / fy, 1.0f);
    return make_float3(dir.x * depth, dir.y * depth, dir.z * depth);
}

// One pixel of the geometry chain's position output, median3x3(position(median3x3(depth))), evaluated straight from the
// raw depth map with the chain's own expressions and padding rules (screen.cu: geometry_chain_kernel): the fused frame
// uses it for the rare pixel that needs its position when the chain itself was skipped (GigsFrame.skip_geometry).
static __device__ __noinline__ float3 depth_pos_pixel(int x, int y, int W, int H, float fx, float fy, const float* __restrict__ depth)
{
    const float cx = float(W) / 2.0f, cy = float(H) / 2.0f;
    float px[9], py[9], pz[9];
#pragma unroll 1
    for (int k = 0; k < 9; ++k) {
        const int qx = x + (k % 3) - 1, qy = y + (k / 3) - 1;
        float3 pos = make_float3(0.f, 0.f, 0.f);                      // outside the image / on its border: zero
        if (qx > 0 && qx < W - 1 && qy > 0 && qy < H - 1) {
            float v[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const int rx = qx + (j % 3) - 1, ry = qy + (j / 3) - 1;   // inside the image here (qx, qy are interior)
                v[j] = depth[(size_t)ry * W + rx];
            }
            pos = back_project(qx, qy, cx, cy, fx, fy, median9(v));
        }
        px[k] = pos.x; py[k] = pos.y; pz[k] = pos.z;
    }
    return make_float3(median9(px), median9(py), median9(pz));
}

}  // namespace gigs
