// SSR's per-pixel epilogue (reference cuda_rasterizer/forward.cu SSR kernel tail, ssr.h:13-16), shared by the march
// kernels (gi_march.cu) and by the fused frame's shade kernel (deferred.cu), which evaluates it itself when no
// direction is marched (start >= step: the gathered radiance is 0 and only the sign / NaN pattern of kD survives).
#pragma once
#include "common.cuh"

namespace gigs {

// true when ssr_epilogue_px returns +0 whatever the normal and the position are (see the comment inside it)
__device__ __forceinline__ bool ssr_epilogue_is_zero(const float3 F0, const float metallic, const float3 diffuse,
                                                     const float nrSamples)
{
    return nrSamples > 0.0f && __float_as_uint(diffuse.x) == 0u && __float_as_uint(diffuse.y) == 0u &&
           __float_as_uint(diffuse.z) == 0u && F0.x >= 0.f && F0.x <= 1.f && F0.y >= 0.f && F0.y <= 1.f && F0.z >= 0.f &&
           F0.z <= 1.f && metallic >= 0.f && metallic <= 1.f;
}

// diffuse: the radiance gathered over the marched directions; nrSamples: their number.
// color = diffuse' * albedo, abd = diffuse' (the factor the backward multiplies the colour gradient with).
__device__ __forceinline__ void ssr_epilogue_px(const float3 normal, const float3 pos, const float3 albedo, const float3 F0,
                                                const float metallic, float3 diffuse, const float nrSamples,
                                                float3& color, float3& abd)
{
    // Nothing gathered (every component +0.0: always the case when no direction is marched) and F0, metallic inside
    // [0, 1] (activated parameters are): the chain below is then +0 exactly, whatever the normal and the position are -
    // fpow lies in (0, 1] for every input (the base is clamped to [1e-6, 1]), so F = F0 + (1 - F0) * fpow <= 1 in float
    // with or without the contraction (rn(F0 + rn(1 - F0)) == 1), kD = (1 - F)(1 - metallic) is a finite value >= +0,
    // and pi * (+0) * (1/n) * kD = +0. The double-precision pow is skipped; NaN / out-of-range inputs take the full path.
    if (ssr_epilogue_is_zero(F0, metallic, diffuse, nrSamples)) {
        abd = make_float3(0.f, 0.f, 0.f);
        color = make_float3(0.f * albedo.x, 0.f * albedo.y, 0.f * albedo.z);
        return;
    }
    const float3 Vd = normalize3(make_float3(-pos.x, -pos.y, -pos.z));
    // fresnelSchlick (ssr.h:13-16): the un-suffixed literals make the base a double subtraction and
    // the power a double pow, rounded to float before the float3 multiply
    const float cosTheta = fmaxf(dot3(normal, Vd), 0.0000001);
    const float fbase = fminf(fmaxf(1.0 - cosTheta, 0.000001), 1.0);
    const float fpow = pow((double)fbase, 5.0);
    float3 F;
    F.x = F0.x + (1.0f - F0.x) * fpow;
    F.y = F0.y + (1.0f - F0.y) * fpow;
    F.z = F0.z + (1.0f - F0.z) * fpow;
    float3 kD = {(float)(1.0 - F.x), (float)(1.0 - F.y), (float)(1.0 - F.z)};
    kD.x *= 1.0 - metallic;
    kD.y *= 1.0 - metallic;
    kD.z *= 1.0 - metallic;
    float3 gd;
    if (nrSamples > 0.0) {
        gd.x = M_PIf * diffuse.x * (1.0 / float(nrSamples)) * kD.x;
        gd.y = M_PIf * diffuse.y * (1.0 / float(nrSamples)) * kD.y;
        gd.z = M_PIf * diffuse.z * (1.0 / float(nrSamples)) * kD.z;
        diffuse.x = gd.x * albedo.x;
        diffuse.y = gd.y * albedo.y;
        diffuse.z = gd.z * albedo.z;
    } else {
        diffuse.x = diffuse.y = diffuse.z = 0.0000001;
        gd.x = gd.y = gd.z = 0.0000001;
    }
    color = diffuse;
    abd = gd;
}

}  // namespace gigs
