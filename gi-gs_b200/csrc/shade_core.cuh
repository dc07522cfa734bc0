// Device core of the split-sum shading: texture sampling (cube / LUT), mip selection, tone curves and the per-pixel
// evaluation shared by shade.cu (reference-shaped pbr_shading op) and deferred.cu (fused deferred frame pass).
#pragma once
#include "common.cuh"

namespace gigs {

struct CubeTaps {
    int idx[4];   // linear texel index into [6,res,res] (multiply by 3 for channels), -1 = unused
    float w[4];
};

// direction -> face, (u,v) in [0,1]; returns -1 for a non-finite / zero direction
__device__ __forceinline__ int cube_face_uv(float x, float y, float z, float& u, float& v)
{
    const float ax = fabsf(x), ay = fabsf(y), az = fabsf(z);
    int idx;
    float c;
    if (az > fmaxf(ax, ay)) { idx = 4; c = z; }
    else if (ay > ax) { idx = 2; c = y; y = z; }
    else { idx = 0; c = x; x = z; }
    if (c < 0.f) idx += 1;
    const float m = __frcp_rn(fabsf(c)) * .5f;
    const float m0 = (idx == 0 || idx == 5) ? -m : m;
    const float m1 = (idx != 2) ? -m : m;
    u = x * m0 + .5f;
    v = y * m1 + .5f;
    if (!isfinite(u) || !isfinite(v)) return -1;
    u = fminf(fmaxf(u, 0.f), 1.f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    return idx;
}

// texel (iu,iv) of `face`, possibly one step outside the face, -> linear index on the adjacent face.
// Works in doubled integer coordinates: texel centres are the odd integers in [-w+1, w-1], face planes at +-w.
// (the reference statement of the fold; the kernels use cube_texel below, which gigs_cube_wrap_selfcheck checks
// against this function)
static __device__ __noinline__ int cube_wrap_texel(int face, int iu, int iv, int w)
{
    const bool ou = (iu < 0 || iu >= w), ov = (iv < 0 || iv >= w);
    if (!ou && !ov) return (face * w + iv) * w + iu;
    if (ou && ov) return -1;  // cube corner: no such texel
    const int s = 2 * iu + 1 - w, t = 2 * iv + 1 - w;
    int p[3];
    switch (face) {
        case 0: p[0] = w;  p[1] = -t; p[2] = -s; break;
        case 1: p[0] = -w; p[1] = -t; p[2] = s;  break;
        case 2: p[0] = s;  p[1] = w;  p[2] = t;  break;
        case 3: p[0] = s;  p[1] = -w; p[2] = -t; break;
        case 4: p[0] = s;  p[1] = -t; p[2] = w;  break;
        default: p[0] = -s; p[1] = -t; p[2] = -w; break;
    }
    const int major = face >> 1;
    int over = -1;
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (a != major && (p[a] > w || p[a] < -w)) over = a;
    // fold across the edge: the overflowing axis becomes the new face plane, the old plane steps one texel in
    p[major] = (p[major] > 0) ? (w - 1) : -(w - 1);
    p[over] = (p[over] > 0) ? w : -w;
    const int nf = 2 * over + (p[over] < 0 ? 1 : 0);
    int s2, t2;
    switch (nf) {
        case 0: t2 = -p[1]; s2 = -p[2]; break;
        case 1: t2 = -p[1]; s2 = p[2];  break;
        case 2: s2 = p[0];  t2 = p[2];  break;
        case 3: s2 = p[0];  t2 = -p[2]; break;
        case 4: s2 = p[0];  t2 = -p[1]; break;
        default: s2 = -p[0]; t2 = -p[1]; break;
    }
    const int iu2 = (s2 + w - 1) >> 1, iv2 = (t2 + w - 1) >> 1;
    return (nf * w + iv2) * w + iu2;
}

// The same map as cube_wrap_texel for the taps the bilinear footprint can produce (at most one step outside the face),
// from a 24-entry table: across edge e of face f the texel at position i along the edge lands on face nf at
// (iu2, iv2) = (ku (w-1) + bu i, kv (w-1) + bv i) with bu, bv in {-1, 0, 1} - affine in i for every face, edge and
// resolution (the table was derived from cube_wrap_texel itself; gigs_cube_wrap_selfcheck compares the two on every
// tap of a level). ~15 instructions inline instead of a ~50-instruction call: a warp pays the edge path whenever ONE
// of its lanes straddles an edge, which at the small mip levels is nearly always (measured with the folding compiled
// out: 0.020 ms of the deferred backward and 0.017 ms of the shade kernel were this path).
// entry: nf | ku << 3 | (bu + 1) << 4 | kv << 6 | (bv + 1) << 7, index face * 4 + {iu < 0, iu >= w, iv < 0, iv >= w}
static __device__ const uint32_t CUBE_EDGE[24] = {0x11c, 0x115, 0x05a, 0x11b, 0x11d, 0x114, 0x112, 0x053, 0x0a1, 0x088, 0x08d, 0x0a4, 0x0c9, 0x0e0, 0x0e4, 0x0cd, 0x119, 0x110, 0x0e2, 0x0a3, 0x118, 0x111, 0x08a, 0x0cb};
__device__ __forceinline__ int cube_texel(int face, int iu, int iv, int w)
{
    const bool ou = (unsigned)iu >= (unsigned)w, ov = (unsigned)iv >= (unsigned)w;
    if (!(ou || ov)) return (face * w + iv) * w + iu;
    if (ou && ov) return -1;  // cube corner: no such texel
    const int edge = ou ? (iu < 0 ? 0 : 1) : (iv < 0 ? 2 : 3);
    const int i = ou ? iv : iu;
    const uint32_t e = CUBE_EDGE[face * 4 + edge];
    const int nf = (int)(e & 7u);
    const int iu2 = (int)((e >> 3) & 1u) * (w - 1) + ((int)((e >> 4) & 3u) - 1) * i;
    const int iv2 = (int)((e >> 6) & 1u) * (w - 1) + ((int)((e >> 7) & 3u) - 1) * i;
    return (nf * w + iv2) * w + iu2;
}

__device__ __forceinline__ CubeTaps cube_taps(float dx, float dy, float dz, int w)
{
    CubeTaps T;
    float u, v;
    const int face = cube_face_uv(dx, dy, dz, u, v);
    if (face < 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { T.idx[k] = -1; T.w[k] = 0.f; }
        return T;
    }
    u = u * (float)w - 0.5f;
    v = v * (float)w - 0.5f;
    const int iu0 = __float2int_rd(u), iv0 = __float2int_rd(v);
    const float fu = u - (float)iu0, fv = v - (float)iv0;
    if (iu0 >= 0 && iv0 >= 0 && iu0 + 1 < w && iv0 + 1 < w) {
        // all four taps inside the face (the common case): no edge / corner folding
        const int base = (face * w + iv0) * w + iu0;
        T.idx[0] = base;         T.w[0] = (1.f - fu) * (1.f - fv);
        T.idx[1] = base + 1;     T.w[1] = fu * (1.f - fv);
        T.idx[2] = base + w;     T.w[2] = (1.f - fu) * fv;
        T.idx[3] = base + w + 1; T.w[3] = fu * fv;
        return T;
    }
    T.idx[0] = cube_texel(face, iu0, iv0, w);             T.w[0] = (1.f - fu) * (1.f - fv);
    T.idx[1] = cube_texel(face, iu0 + 1, iv0, w);         T.w[1] = fu * (1.f - fv);
    T.idx[2] = cube_texel(face, iu0, iv0 + 1, w);         T.w[2] = (1.f - fu) * fv;
    T.idx[3] = cube_texel(face, iu0 + 1, iv0 + 1, w); T.w[3] = fu * fv;
    // at a cube corner the missing texel is the mean of the other three
    int missing = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (T.idx[k] < 0) missing = k;
    if (missing >= 0) {
        const float share = T.w[missing] * 0.33333333f;
#pragma unroll
        for (int k = 0; k < 4; ++k) T.w[k] = (k == missing) ? 0.f : T.w[k] + share;
    }
    return T;
}

__device__ __forceinline__ float3 cube_fetch(const float* __restrict__ tex, const CubeTaps& T)
{
    float3 r = make_float3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (T.idx[k] >= 0) {
            const float* p = tex + 3 * (size_t)T.idx[k];
            r.x += T.w[k] * p[0];
            r.y += T.w[k] * p[1];
            r.z += T.w[k] * p[2];
        }
    }
    return r;
}

struct LutTaps {
    int i00, i10, i01, i11;
    float fu, fv;
    bool clampV;
};
__device__ __forceinline__ LutTaps lut_taps(float u, float v, int res)
{
    LutTaps L;
    u = u * (float)res - 0.5f;
    v = v * (float)res - 0.5f;
    u = fminf(fmaxf(u, 0.f), res - 1.f);
    v = fminf(fmaxf(v, 0.f), res - 1.f);
    const bool clampU = (u == 0.f || u == res - 1.f);
    L.clampV = (v == 0.f || v == res - 1.f);
    const int iu0 = __float2int_rd(u), iv0 = __float2int_rd(v);
    const int iu1 = iu0 + (clampU ? 0 : 1), iv1 = iv0 + (L.clampV ? 0 : 1);
    L.fu = u - (float)iu0;
    L.fv = v - (float)iv0;
    L.i00 = iv0 * res + iu0; L.i10 = iv0 * res + iu1;
    L.i01 = iv1 * res + iu0; L.i11 = iv1 * res + iu1;
    return L;
}

__device__ __forceinline__ float mip_level(float r, float rmin, float rmax, int nlev, float& dlevel_dr)
{
    // pbr/light.py:142-152
    float lvl;
    if (r < rmax) {
        const float rc = fminf(fmaxf(r, rmin), rmax);
        lvl = (rc - rmin) / (rmax - rmin) * (float)(nlev - 2);
        dlevel_dr = (r >= rmin && r <= rmax) ? (float)(nlev - 2) / (rmax - rmin) : 0.f;
    } else {
        const float rc = fminf(fmaxf(r, rmax), 1.0f);
        lvl = (rc - rmax) / (1.0f - rmax) + (float)nlev - 2.f;
        dlevel_dr = (r >= rmax && r <= 1.0f) ? 1.f / (1.0f - rmax) : 0.f;
    }
    return lvl;
}

// x^y for the colour-space curves (x > 0): exp2(y * log2 x) with the hardware log2 / exp2 units. Relative error
// ~1e-6 against powf's 2 ulp, two orders below the 1e-4 image gate, at ~8 instructions instead of ~70 (powf was
// 23 % of the deferred shading kernel's stall samples).
__device__ __forceinline__ float fast_pow(float x, float y) { return exp2f(y * __log2f(x)); }

__device__ __forceinline__ float srgb_fwd(float x)
{
    // pbr/shade.py:46-52
    const float eps = 1.1920928955078125e-07f;
    const float s0 = (323.f / 25.f) * x;
    const float s1 = (211.f * fast_pow(fmaxf(x, eps), 5.f / 12.f) - 11.f) / 200.f;
    return (x <= 0.0031308f) ? s0 : s1;
}
__device__ __forceinline__ float srgb_bwd(float x)
{
    const float eps = 1.1920928955078125e-07f;
    if (x <= 0.0031308f) return 323.f / 25.f;
    if (x < eps) return 0.f;
    return (211.f / 200.f) * (5.f / 12.f) * fast_pow(x, 5.f / 12.f - 1.f);
}
__device__ __forceinline__ float aces_raw(float x)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return (x * (a * x + b)) / (x * (c * x + d) + e);
}
__device__ __forceinline__ float aces_raw_bwd(float x)
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    const float num = x * (a * x + b), den = x * (c * x + d) + e;
    return ((2.f * a * x + b) * den - num * (2.f * c * x + d)) / (den * den);
}

struct ShadeParams {
    int W, H, n_lev, diffuse_res, lut_res, tone, gamma;
    int spec_res[8];
    const float* spec[8];
    const float* diffuse;
    const float* lut;
    float rmin, rmax;
    const float *normals, *view_dirs, *albedo, *roughness, *metallic, *occlusion, *background;
    const uint8_t* mask;
    float *render_rgb, *diffuse_rgb, *specular_rgb, *diffuse_light;
    const float *g_render, *g_diffuse, *g_specular;
    float *g_albedo, *g_roughness, *g_metallic, *g_diffuse_tex;
    float* g_spec[8];
    // Privatised accumulation of the small textures (deferred.cu): a stride > 0 means the pointer above addresses
    // TEX_COPIES consecutive copies of the texture, `stride` floats apart, and a CTA adds into copy
    // blockIdx.x % TEX_COPIES. A 16x16 or 32x32 cube level takes millions of reductions on a few thousand
    // addresses per frame, which serialise in the L2 (measured: 0.09 of the 0.21 ms of the backward kernel).
    uint32_t g_spec_stride[8];
    uint32_t g_diffuse_stride;
};

// everything the forward computes for one pixel, kept for reuse by the backward
struct PixelShade {
    float3 alb, dl, spec, F0, diffuse_rgb, specular_rgb, lin;  // lin = pre-tone linear sum
    float rough, metal, occ, fgx, fgy, flevel, dlevel_dr;
    int l0, l1;
    CubeTaps td, t0, t1;
    LutTaps lt;
    float3 s0, s1;
};

struct ShadeIn {
    float3 n, v, alb;
    float rough, metal, occ;
};

__device__ __forceinline__ void shade_eval(const ShadeParams& p, const ShadeIn& in, PixelShade& S);

__device__ __forceinline__ void shade_pixel(const ShadeParams& p, size_t id, size_t HW, PixelShade& S)
{
    ShadeIn in;
    in.n = make_float3(p.normals[id], p.normals[HW + id], p.normals[2 * HW + id]);
    in.v = make_float3(p.view_dirs[id], p.view_dirs[HW + id], p.view_dirs[2 * HW + id]);
    in.alb = make_float3(p.albedo[id], p.albedo[HW + id], p.albedo[2 * HW + id]);
    in.rough = p.roughness[id];
    in.metal = p.metallic ? p.metallic[id] : 0.f;
    in.occ = p.occlusion ? p.occlusion[id] : 1.f;
    shade_eval(p, in, S);
}

__device__ __forceinline__ void shade_eval(const ShadeParams& p, const ShadeIn& in, PixelShade& S)
{
    const float3 n = in.n, v = in.v;
    S.alb = in.alb;
    S.rough = in.rough;
    S.metal = in.metal;
    S.occ = in.occ;

    const float ndv = fmaxf(n.x * v.x + n.y * v.y + n.z * v.z, 0.0f);
    const float3 ref = make_float3(2.0f * ndv * n.x - v.x, 2.0f * ndv * n.y - v.y, 2.0f * ndv * n.z - v.z);
    // x @ T^T with T = [[0,-1,0],[0,0,1],[-1,0,0]]  ->  (-x.y, x.z, -x.x)
    const float3 nT = make_float3(-n.y, n.z, -n.x);
    const float3 vT = make_float3(-v.y, v.z, -v.x);
    const float3 rT = make_float3(-ref.y, ref.z, -ref.x);

    S.td = cube_taps(nT.x, nT.y, nT.z, p.diffuse_res);
    S.dl = cube_fetch(p.diffuse, S.td);
    if (p.occlusion) { S.dl.x *= S.occ; S.dl.y *= S.occ; S.dl.z *= S.occ; }
    S.diffuse_rgb = make_float3(S.dl.x * S.alb.x, S.dl.y * S.alb.y, S.dl.z * S.alb.z);

    const float NoV = fminf(fmaxf(nT.x * vT.x + nT.y * vT.y + nT.z * vT.z, 1e-4f), 1.0f);
    S.lt = lut_taps(NoV, S.rough, p.lut_res);
    {
        const float2* L = reinterpret_cast<const float2*>(p.lut);
        const float2 a00 = L[S.lt.i00], a10 = L[S.lt.i10], a01 = L[S.lt.i01], a11 = L[S.lt.i11];
        const float bx0 = a00.x + S.lt.fu * (a10.x - a00.x), bx1 = a01.x + S.lt.fu * (a11.x - a01.x);
        const float by0 = a00.y + S.lt.fu * (a10.y - a00.y), by1 = a01.y + S.lt.fu * (a11.y - a01.y);
        S.fgx = bx0 + S.lt.fv * (bx1 - bx0);
        S.fgy = by0 + S.lt.fv * (by1 - by0);
    }
    float lvl = mip_level(S.rough, p.rmin, p.rmax, p.n_lev, S.dlevel_dr);
    const float lmax = (float)(p.n_lev - 1);
    if (lvl < 0.f || lvl > lmax) S.dlevel_dr = 0.f;
    lvl = fminf(fmaxf(lvl, 0.f), lmax);
    S.l0 = __float2int_rd(lvl);
    S.l1 = min(S.l0 + 1, p.n_lev - 1);
    S.flevel = lvl - (float)S.l0;
    S.t0 = cube_taps(rT.x, rT.y, rT.z, p.spec_res[S.l0]);
    S.s0 = cube_fetch(p.spec[S.l0], S.t0);
    if (S.l1 != S.l0) {
        S.t1 = cube_taps(rT.x, rT.y, rT.z, p.spec_res[S.l1]);
        S.s1 = cube_fetch(p.spec[S.l1], S.t1);
        S.spec = make_float3(S.s0.x + S.flevel * (S.s1.x - S.s0.x), S.s0.y + S.flevel * (S.s1.y - S.s0.y),
                             S.s0.z + S.flevel * (S.s1.z - S.s0.z));
    } else {
        S.s1 = S.s0;
        S.spec = S.s0;
    }
    if (p.metallic)
        S.F0 = make_float3((1.0f - S.metal) * 0.04f + S.alb.x * S.metal, (1.0f - S.metal) * 0.04f + S.alb.y * S.metal,
                           (1.0f - S.metal) * 0.04f + S.alb.z * S.metal);
    else
        S.F0 = make_float3(0.04f, 0.04f, 0.04f);
    const float3 refl = make_float3(S.F0.x * S.fgx + S.fgy, S.F0.y * S.fgx + S.fgy, S.F0.z * S.fgx + S.fgy);
    S.specular_rgb = make_float3(S.spec.x * refl.x, S.spec.y * refl.y, S.spec.z * refl.z);
    S.lin = make_float3(S.diffuse_rgb.x + S.specular_rgb.x, S.diffuse_rgb.y + S.specular_rgb.y,
                        S.diffuse_rgb.z + S.specular_rgb.z);
}

// ------------------------------------------------------------------------------------------------
// Backward pieces shared by shade_backward_kernel (shade.cu) and deferred_backward_kernel (deferred.cu)
// ------------------------------------------------------------------------------------------------
struct ShadeGrad {
    float g_alb[3], g_rough, g_metal;   // dL/d albedo, roughness (the value the LUT / mip lookup saw), metallic
    float g_dl[3], g_spec[3];           // dL/d (diffuse cube sample), dL/d (specular cube sample)
};

// Chain through gamma and tone/clamp: g (dL/d render value of channel k, already masked) -> dL/d linear sum.
__device__ __forceinline__ float shade_tone_bwd(const ShadeParams& p, float x, float g)
{
    float y, dy_dx;
    if (p.tone) {
        const float r = aces_raw(x);
        y = fminf(fmaxf(r, 0.f), 1.f);
        dy_dx = (r >= 0.f && r <= 1.f) ? aces_raw_bwd(x) : 0.f;
    } else {
        y = fminf(fmaxf(x, 0.f), 1.f);
        dy_dx = (x >= 0.f && x <= 1.f) ? 1.f : 0.f;
    }
    if (p.gamma) g *= srgb_bwd(y);
    return g * dy_dx;
}

// gd / gs: dL/d diffuse_rgb and dL/d specular_rgb (linear) of the pixel S was evaluated for.
__device__ __forceinline__ void shade_material_bwd(const ShadeParams& p, const PixelShade& S, const float gd[3],
                                                   const float gs[3], ShadeGrad& G)
{
    const float alb[3] = {S.alb.x, S.alb.y, S.alb.z};
    const float dl[3] = {S.dl.x, S.dl.y, S.dl.z};
    const float spec[3] = {S.spec.x, S.spec.y, S.spec.z};
    const float F0[3] = {S.F0.x, S.F0.y, S.F0.z};
    const float s0[3] = {S.s0.x, S.s0.y, S.s0.z}, s1[3] = {S.s1.x, S.s1.y, S.s1.z};
    float g_fgx = 0.f, g_fgy = 0.f, g_metal = 0.f, g_level = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        G.g_alb[k] = gd[k] * dl[k];
        G.g_dl[k] = gd[k] * alb[k] * (p.occlusion ? S.occ : 1.f);
        const float refl = F0[k] * S.fgx + S.fgy;
        G.g_spec[k] = gs[k] * refl;
        const float g_refl = gs[k] * spec[k];
        const float g_F0 = g_refl * S.fgx;
        g_fgx += g_refl * F0[k];
        g_fgy += g_refl;
        if (p.metallic) {
            G.g_alb[k] += g_F0 * S.metal;
            g_metal += g_F0 * (alb[k] - 0.04f);
        }
        g_level += G.g_spec[k] * (s1[k] - s0[k]);
    }
    // roughness: LUT v-coordinate + mip level
    float g_rough = 0.f;
    if (!S.lt.clampV) {
        const float2* L = reinterpret_cast<const float2*>(p.lut);
        const float2 a00 = L[S.lt.i00], a10 = L[S.lt.i10], a01 = L[S.lt.i01], a11 = L[S.lt.i11];
        const float dfx = ((a01.x - a00.x) * (1.f - S.lt.fu) + (a11.x - a10.x) * S.lt.fu) * (float)p.lut_res;
        const float dfy = ((a01.y - a00.y) * (1.f - S.lt.fu) + (a11.y - a10.y) * S.lt.fu) * (float)p.lut_res;
        g_rough += g_fgx * dfx + g_fgy * dfy;
    }
    if (S.l1 != S.l0) g_rough += g_level * S.dlevel_dr;
    G.g_rough = g_rough;
    G.g_metal = g_metal;
}

// a lane that shades nothing: taps that add nothing, so it can still take part in the warp reductions
__device__ __forceinline__ void shade_dead_lane(PixelShade& S, ShadeGrad& G)
{
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        S.td.idx[t] = S.t0.idx[t] = S.t1.idx[t] = -1;
        S.td.w[t] = S.t0.w[t] = S.t1.w[t] = 0.f;
    }
    S.l0 = S.l1 = 0;
    S.flevel = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) G.g_dl[k] = G.g_spec[k] = 0.f;
}

// Texel gradients are the hot spot of the shading backward: neighbouring pixels hit the same few texels (a 16x16
// face texel covers ~50 px at 800x800), so per-lane atomics serialise 32-way. Each tap is therefore reduced over
// runs of equal texel index inside the warp first (segmented scan, 5 shuffle steps, flags shared by the three
// channels) and only the last lane of a run issues the atomic.  key < 0 = nothing to add. All 32 lanes must call.
constexpr int SHB_MAX_DIFFUSE = 6 * 16 * 16 * 3;
constexpr int TEX_COPIES = 32;

__device__ __forceinline__ void warp_run_reduce3(const int key, float a, float b, float c, float* dst, const bool shared,
                                                 const int lane)
{
    const unsigned full = 0xffffffffu;
    if (__all_sync(full, key < 0)) return;
    const int prev = __shfl_up_sync(full, key, 1);
    int f = (lane == 0 || prev != key) ? 1 : 0;  // a run head lies within the last d lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float au = __shfl_up_sync(full, a, d), bu = __shfl_up_sync(full, b, d), cu = __shfl_up_sync(full, c, d);
        const int fu = __shfl_up_sync(full, f, d);
        if (lane >= d && !f) {
            a += au;
            b += bu;
            c += cu;
            f = fu;
        }
    }
    const int next = __shfl_down_sync(full, key, 1);
    const bool tail = (lane == 31) || (next != key);
    if (tail && key >= 0) {
        if (shared) {
            if (a != 0.f) atomicAdd(dst + 0, a);
            if (b != 0.f) atomicAdd(dst + 1, b);
            if (c != 0.f) atomicAdd(dst + 2, c);
        } else {
            if (a != 0.f) red_add_f32(dst + 0, a);
            if (b != 0.f) red_add_f32(dst + 1, b);
            if (c != 0.f) red_add_f32(dst + 2, c);
        }
    }
}

// Warp-uniform: every lane of the warp must reach this call (dead lanes via shade_dead_lane).
// The diffuse texture accumulates in shared memory (s_dtex, flushed by the caller) when use_smem.
__device__ __forceinline__ void shade_texel_scatter(const ShadeParams& p, const PixelShade& S, const ShadeGrad& G,
                                                    float* s_dtex, const bool use_smem, const int lane)
{
    const uint32_t copy = blockIdx.x & (TEX_COPIES - 1);
    if (p.g_diffuse_tex) {
        float* dbase = p.g_diffuse_tex + (size_t)copy * p.g_diffuse_stride;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int key = (S.td.w[t] != 0.f) ? S.td.idx[t] : -1;
            float* dst = use_smem ? (s_dtex + 3 * max(key, 0)) : (dbase + 3 * (size_t)max(key, 0));
            warp_run_reduce3(key, G.g_dl[0] * S.td.w[t], G.g_dl[1] * S.td.w[t], G.g_dl[2] * S.td.w[t], dst, use_smem,
                             lane);
        }
    }
    const float w0 = (S.l1 != S.l0) ? (1.f - S.flevel) : 1.f;
    float* tex0 = p.g_spec[S.l0];
    if (tex0) tex0 += (size_t)copy * p.g_spec_stride[S.l0];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int idx = (tex0 != nullptr) ? S.t0.idx[t] : -1;
        const int key = (idx >= 0) ? ((S.l0 << 24) | idx) : -1;
        const float w = S.t0.w[t] * w0;
        warp_run_reduce3(key, G.g_spec[0] * w, G.g_spec[1] * w, G.g_spec[2] * w,
                         tex0 ? tex0 + 3 * (size_t)max(idx, 0) : nullptr, false, lane);
    }
    float* tex1 = (S.l1 != S.l0) ? p.g_spec[S.l1] : nullptr;
    if (tex1) tex1 += (size_t)copy * p.g_spec_stride[S.l1];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int idx = (tex1 != nullptr) ? S.t1.idx[t] : -1;
        const int key = (idx >= 0) ? ((S.l1 << 24) | idx) : -1;
        const float w = S.t1.w[t] * S.flevel;
        warp_run_reduce3(key, G.g_spec[0] * w, G.g_spec[1] * w, G.g_spec[2] * w,
                         tex1 ? tex1 + 3 * (size_t)max(idx, 0) : nullptr, false, lane);
    }
}

}  // namespace gigs
