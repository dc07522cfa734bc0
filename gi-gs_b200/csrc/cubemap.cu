// Cubemap prefilter of the environment light: CubemapLight.build_mips (reference pbr/light.py:154-170), the step that
// produces the textures the split-sum shading consumes and that the reference runs once per PBR training step
// (train.py:340). SURVEY.md §8f-1.
//
// Follows (reference, read-only):
//   pbr/renderutils/c_src/cubemap.cu:17-47      pixel_area, cube_to_dir
//   pbr/renderutils/c_src/cubemap.cu:110-168    diffuse (cosine) filter, forward and backward
//   pbr/renderutils/c_src/cubemap.cu:173-246    brute-force bounds of the GGX lobe per texel and face
//   pbr/renderutils/c_src/cubemap.cu:248-350    GGX split-sum specular filter, forward and backward
//   pbr/light.py:54-79                          cubemap_mip: 2x2 average pool; its backward bilinearly samples 0.25*dout
//   pbr/renderutils/ops.py:391-459              autograd wrappers (the col / wsum division happens in Python there)
//
// What is ours:
//  * per-texel direction and solid-angle weight come from a table built once per resolution instead of a normalise +
//    4 atan per (output, input) pair;
//  * ONE filter launch covers every level of the light (5 GGX levels + the cosine filter at base_res 256): a plan of
//    segments, each level given G = 1..32 lanes per output texel so that every lane walks ~64-160 partner texels
//    (the reference runs one thread per texel: 1536 threads x 1536 partners at the 16^2 level). Lanes of one texel
//    stride the columns of each row of the cone's bounding box and fold their partial sums with shuffles;
//  * both backward filters are GATHERS over the same cone bounds (the cone test dot(L, V) >= cutoff is symmetric), so
//    there are no atomics and the gradients are deterministic (the reference scatters 3 atomicAdd per pair);
//  * the col / wsum division of ops.py:456 is fused into the forward and its chain rule into the backward;
//  * the mip chain (4 average pools) is one kernel that also emits the 16-byte padded texels the filter reads.
// Per-pair weights follow the reference's expression forms (true divisions in safeNormalize, same FMA shapes): at
// roughness 0.08 the GGX term is ill-conditioned in dot(V,H) (one ulp moves the central weights by 0.3 %), so the
// arithmetic that feeds it is kept identical; only the final alphaSqr / (d*d*pi) runs in float instead of double
// (<= 1.5 ulp on a weight, no amplification) and the sums run in a different order.
#include <algorithm>
#include <cstring>
#include "shade_core.cuh"

namespace gigs {

__device__ __forceinline__ float cm_pixel_area(int x, int y, int N)
{
    if (N > 1) {
        const int H = N / 2;
        x = abs(x - H);
        y = abs(y - H);
        const float dx = atanf((float)(x + 1) / (float)H) - atanf((float)x / (float)H);
        const float dy = atanf((float)(y + 1) / (float)H) - atanf((float)y / (float)H);
        return dx * dy;
    }
    return 1.f;
}

__device__ __forceinline__ float3 cm_safe_normalize(float x, float y, float z)
{
    const float l = sqrtf(x * x + y * y + z * z);
    return l > 0.0f ? make_float3(x / l, y / l, z / l) : make_float3(0.f, 0.f, 0.f);
}

__device__ __forceinline__ float3 cm_cube_to_dir(int x, int y, int side, int N)
{
    const float fx = 2.0f * (((float)x + 0.5f) / (float)N) - 1.0f;
    const float fy = 2.0f * (((float)y + 0.5f) / (float)N) - 1.0f;
    switch (side) {
        case 0: return cm_safe_normalize(1, -fy, -fx);
        case 1: return cm_safe_normalize(-1, -fy, fx);
        case 2: return cm_safe_normalize(fx, 1, fy);
        case 3: return cm_safe_normalize(fx, -1, -fy);
        case 4: return cm_safe_normalize(fx, -fy, 1);
        default: return cm_safe_normalize(-fx, -fy, -1);
    }
}

// table[(s*N + y)*N + x] = (unit direction of the texel centre, pixel_area)
__global__ void __launch_bounds__(256) cm_table_kernel(const int N, float4* __restrict__ table)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 6 * N * N) return;
    const int x = i % N, y = (i / N) % N, s = i / (N * N);
    const float3 d = cm_cube_to_dir(x, y, s, N);
    table[i] = make_float4(d.x, d.y, d.z, cm_pixel_area(x, y, N));
}

// ---------------------------------------------------------------------------------------------
// Bounds of the cone {L : dot(L, V) >= cutoff} on every face, for every texel direction V.
// Same brute force as the reference, including its 16x16-tile interval cull (so the boxes are identical); the
// corners of a tile are texel coordinates up to N inclusive, hence computed here and not read from the table.
// bounds[(texel*6 + face)] = (xmin, xmax, ymin, ymax); empty = (N-1, 0, N-1, 0).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
cm_bounds_kernel(const int N, const float cutoff, const float4* __restrict__ table, short4* __restrict__ bounds)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= 6 * N * N) return;
    const float4 tv = table[i];
    const float3 V = make_float3(tv.x, tv.y, tv.z);
    constexpr int TS = 16;
    const int nt = (N + TS - 1) / TS;
    for (int s = 0; s < 6; ++s) {
        int mnx = N - 1, mxx = 0, mny = N - 1, mxy = 0;
        for (int tx = 0; tx < nt; ++tx) {
            for (int ty = 0; ty < nt; ++ty) {
                const int tsx = tx * TS, tsy = ty * TS;
                const int tex = min((tx + 1) * TS, N), tey = min((ty + 1) * TS, N);
                const float3 L0 = cm_cube_to_dir(tsx, tsy, s, N), L1 = cm_cube_to_dir(tex, tsy, s, N);
                const float3 L2 = cm_cube_to_dir(tsx, tey, s, N), L3 = cm_cube_to_dir(tex, tey, s, N);
                const float minx = fminf(fminf(L0.x, L1.x), fminf(L2.x, L3.x)), maxx = fmaxf(fmaxf(L0.x, L1.x), fmaxf(L2.x, L3.x));
                const float miny = fminf(fminf(L0.y, L1.y), fminf(L2.y, L3.y)), maxy = fmaxf(fmaxf(L0.y, L1.y), fmaxf(L2.y, L3.y));
                const float minz = fminf(fminf(L0.z, L1.z), fminf(L2.z, L3.z)), maxz = fmaxf(fmaxf(L0.z, L1.z), fmaxf(L2.z, L3.z));
                const float maxdp = fmaxf(minx * V.x, maxx * V.x) + fmaxf(miny * V.y, maxy * V.y) + fmaxf(minz * V.z, maxz * V.z);
                if (maxdp >= cutoff) {
                    for (int y = tsy; y < tey; ++y)
                        for (int x = tsx; x < tex; ++x) {
                            const float4 tl = table[(s * N + y) * N + x];
                            if (tl.x * V.x + tl.y * V.y + tl.z * V.z >= cutoff) {
                                mnx = min(mnx, x); mxx = max(mxx, x);
                                mny = min(mny, y); mxy = max(mxy, y);
                            }
                        }
                }
            }
        }
        bounds[(size_t)i * 6 + s] = make_short4((short)mnx, (short)mxx, (short)mny, (short)mxy);
    }
}

// ---------------------------------------------------------------------------------------------
// The filter plan: one launch, several segments (a level of the GGX chain or the cosine filter each).
// ---------------------------------------------------------------------------------------------
struct CmSeg {
    int cta_begin;        // first CTA of the segment (segments are ordered by cta_begin)
    int N;                // face resolution
    int lg;               // log2 of the lanes per output texel
    int kind;             // 0 = GGX specular (bounds + cutoff), 1 = cosine diffuse (all texels)
    float alpha_sqr;      // roughness^4
    float cutoff;
    const float4* table;  // [6*N*N] (dir, area)
    const short4* bounds; // [6*N*N][6]
    const void* src;      // PACKED: float4 [6*N*N] (rgb, -); else float [6*N*N][3]
    float* dst;           // [6*N*N][3]
    float* wsum;          // forward: out; unpacked backward: in; else unused
};
constexpr int CM_MAX_SEG = 12;
struct CmPlan {
    int nseg;
    int pad;
    CmSeg seg[CM_MAX_SEG];
};

__device__ __forceinline__ float cm_spec_weight(const float4 L, const float4 V, const float dotLV, const float alphaSqr)
{
    // H = safeNormalize(L + VNR); w = max(dot(L,VNR),0) * ndfGGX(alphaSqr, max(dot(VNR,H),0)) * pixel_area(L) / 4
    const float3 H = cm_safe_normalize(L.x + V.x, L.y + V.y, L.z + V.z);
    const float wiDotN = fmaxf(dotLV, 0.0f);
    const float VdotH = fmaxf(V.x * H.x + V.y * H.y + V.z * H.z, 0.0f);
    const float c = fminf(fmaxf(VdotH, 0.0f), 1.0f);
    const float d = (c * alphaSqr - c) * c + 1.0f;
    const float ndf = alphaSqr / (d * d * 3.14159274f);
    return wiDotN * ndf * L.w / 4.0f;
}

template <bool PACKED>
__device__ __forceinline__ float3 cm_load3(const void* src, const int j)
{
    if (PACKED) {
        const float4 c = reinterpret_cast<const float4*>(src)[j];
        return make_float3(c.x, c.y, c.z);
    }
    const float* p = reinterpret_cast<const float*>(src) + 3 * (size_t)j;
    return make_float3(p[0], p[1], p[2]);
}

// forward : dst(V) = sum_L src(L) w(V,L) / wsum(V), wsum(V) = sum_L w(V,L)           (kind 0; kind 1 has no wsum)
// backward: dst(L) = sum_V q(V) w(V,L); PACKED: src holds q = g / wsum; else src = g and wsum is divided per pair
template <bool BACKWARD, bool PACKED>
__global__ void __launch_bounds__(256) cm_filter_kernel(const __grid_constant__ CmPlan plan)
{
    pdl_enter();
    int k = 0;
    for (int q = 1; q < plan.nseg; ++q)
        if ((int)blockIdx.x >= plan.seg[q].cta_begin) k = q;
    const CmSeg& S = plan.seg[k];
    const int N = S.N, lg = S.lg, G = 1 << lg;
    const int n = 6 * N * N;
    const int t = ((int)blockIdx.x - S.cta_begin) * 256 + (int)threadIdx.x;
    int i = t >> lg;
    const int sub = t & (G - 1);
    const bool live = i < n;
    if (!live) i = n - 1;   // keep the lane in the shuffles below
    const float4* __restrict__ table = S.table;
    const void* __restrict__ src = S.src;
    const float4 tv = table[i];
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, ws = 0.f;
    if (S.kind == 0) {
        const float alphaSqr = S.alpha_sqr, cutoff = S.cutoff;
        const short4* __restrict__ bnd = S.bounds + (size_t)i * 6;
        for (int s = 0; s < 6; ++s) {
            const short4 b = bnd[s];
            if (b.x > b.y) continue;
            for (int y = b.z; y <= b.w; ++y) {
                const int row = (s * N + y) * N;
                for (int x = b.x + sub; x <= b.y; x += G) {
                    const int j = row + x;
                    const float4 tl = table[j];
                    const float dotLV = tl.x * tv.x + tl.y * tv.y + tl.z * tv.z;
                    if (dotLV >= cutoff) {
                        // forward: this thread is the output direction V, j the light texel L; backward: swapped
                        float w = BACKWARD ? cm_spec_weight(tv, tl, dotLV, alphaSqr) : cm_spec_weight(tl, tv, dotLV, alphaSqr);
                        if (BACKWARD && !PACKED) w = w / S.wsum[j];
                        const float3 c = cm_load3<PACKED>(src, j);
                        c0 += c.x * w;
                        c1 += c.y * w;
                        c2 += c.z * w;
                        if (!BACKWARD) ws += w;
                    }
                }
            }
        }
    } else {
        for (int j = sub; j < n; j += G) {
            const float4 tl = table[j];
            const float costheta = fminf(fmaxf(tv.x * tl.x + tv.y * tl.y + tv.z * tl.z, 0.0f), 0.999f);
            // forward: area of the light texel j; backward: this thread IS the light texel
            const float w = costheta * (BACKWARD ? tv.w : tl.w) / 3.141592f;
            const float3 c = cm_load3<PACKED>(src, j);
            c0 += c.x * w;
            c1 += c.y * w;
            c2 += c.z * w;
        }
    }
    for (int off = G >> 1; off > 0; off >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, off);
        c1 += __shfl_xor_sync(0xffffffffu, c1, off);
        c2 += __shfl_xor_sync(0xffffffffu, c2, off);
        ws += __shfl_xor_sync(0xffffffffu, ws, off);
    }
    if (live && sub == 0) {
        float* o = S.dst + 3 * (size_t)i;
        if (!BACKWARD && S.kind == 0) {
            o[0] = c0 / ws; o[1] = c1 / ws; o[2] = c2 / ws;
            S.wsum[i] = ws;
        } else {
            o[0] = c0; o[1] = c1; o[2] = c2;
        }
    }
}

// lanes per output texel: ~64+ partner texels per lane, never wider than a row of the face
static int cm_lanes_log2(int N, int kind, float cutoff)
{
    const double texels = 6.0 * N * N;
    const double pairs = kind == 0 ? texels * (1.0 - (double)cutoff) * 0.5 * 1.27 : texels;
    int lg = 0;
    while (lg < 5 && (2 << lg) <= N && pairs / (double)(2 << lg) >= 64.0) ++lg;
    return lg;
}

static void cm_plan_add(CmPlan& plan, int& ctas, int N, int kind, float roughness, float cutoff, const float* table,
                        const int16_t* bounds, const void* src, float* dst, float* wsum)
{
    CmSeg& S = plan.seg[plan.nseg++];
    S.cta_begin = ctas;
    S.N = N;
    S.kind = kind;
    S.lg = cm_lanes_log2(N, kind, cutoff);
    const float alpha = roughness * roughness;
    S.alpha_sqr = alpha * alpha;
    S.cutoff = cutoff;
    S.table = (const float4*)table;
    S.bounds = (const short4*)bounds;
    S.src = src;
    S.dst = dst;
    S.wsum = wsum;
    const long long lanes = (long long)6 * N * N << S.lg;
    ctas += (int)((lanes + 255) / 256);
}

// ---------------------------------------------------------------------------------------------
// Mip chain: up to 4 successive 2x2 average pools per launch (a 16x16 block of source texels per CTA), each level
// also written as 16-byte padded texels for the filter. pbr/light.py:56-60 (avg_pool2d: ((a+b)+c)+d, then /4).
// ---------------------------------------------------------------------------------------------
struct CmChainArgs {
    int R;             // source resolution (multiple of 16)
    int nh;            // halvings to do, 0..4
    const float* src;  // SRC4 ? float4 [6,R,R] : float [6,R,R,3]
    float4* pack0;     // padded copy of the source level (may be null)
    float4* out[4];    // padded levels R/2, R/4, ...
};

template <bool SRC4>
__global__ void __launch_bounds__(256) cm_chain_kernel(const __grid_constant__ CmChainArgs a)
{
    pdl_enter();
    __shared__ float sm[3][16][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int f = blockIdx.z, R = a.R;
    const int x = blockIdx.x * 16 + tx, y = blockIdx.y * 16 + ty;
    const size_t idx = ((size_t)f * R + y) * R + x;
    float3 v;
    if (SRC4) {
        const float4 q = reinterpret_cast<const float4*>(a.src)[idx];
        v = make_float3(q.x, q.y, q.z);
    } else {
        v = make_float3(a.src[3 * idx], a.src[3 * idx + 1], a.src[3 * idx + 2]);
        if (a.pack0) a.pack0[idx] = make_float4(v.x, v.y, v.z, 0.f);
    }
    sm[0][ty][tx] = v.x; sm[1][ty][tx] = v.y; sm[2][ty][tx] = v.z;
    __syncthreads();
    for (int h = 1; h <= a.nh; ++h) {
        const int w = 16 >> h;
        const bool act = tx < w && ty < w;
        float r[3];
        if (act) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                r[c] = (sm[c][2 * ty][2 * tx] + sm[c][2 * ty][2 * tx + 1] + sm[c][2 * ty + 1][2 * tx] + sm[c][2 * ty + 1][2 * tx + 1]) * 0.25f;
        }
        __syncthreads();
        if (act) {
            sm[0][ty][tx] = r[0]; sm[1][ty][tx] = r[1]; sm[2][ty][tx] = r[2];
            const int Rh = R >> h;
            a.out[h - 1][((size_t)f * Rh + (blockIdx.y * w + ty)) * Rh + blockIdx.x * w + tx] = make_float4(r[0], r[1], r[2], 0.f);
        }
        __syncthreads();
    }
}

// plain (unpadded) 2x2 pool for the stand-alone cubemap_mip op
__global__ void __launch_bounds__(256) cm_mip_forward_kernel(const int No, const float* __restrict__ in, float* __restrict__ out)
{
    pdl_enter();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 6 * No * No) return;
    const int x = i % No, y = (i / No) % No, s = i / (No * No);
    const int Ni = 2 * No;
    const float* p = in + ((size_t)(s * Ni + 2 * y) * Ni + 2 * x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        out[3 * (size_t)i + c] = (p[c] + p[3 + c] + p[(size_t)Ni * 3 + c] + p[(size_t)Ni * 3 + 3 + c]) * 0.25f;
}

// The reference's backward of the pool is NOT its adjoint but a seamless bilinear cube lookup of 0.25 * dout at the
// fine texel-centre directions (pbr/light.py:62-79).  out(i) = [accumulate ? out(i) : 0] + [add ? add(i) : 0]
//                                                              + 0.25 * lookup(coarse_a [+ coarse_b])(dir_i)
__global__ void __launch_bounds__(256)
cm_mip_backward_kernel(const int Nc, const float* __restrict__ coarse_a, const float* __restrict__ coarse_b,
                       const float* __restrict__ add, float* __restrict__ out, const bool accumulate)
{
    pdl_enter();
    const int Nf = 2 * Nc;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 6 * Nf * Nf) return;
    const int x = i % Nf, y = (i / Nf) % Nf, s = i / (Nf * Nf);
    // linspace(-1 + 1/res, 1 - 1/res, res)[k] = -1 + (2k+1)/res : the texel centres
    const float gx = -1.0f + (float)(2 * x + 1) / (float)Nf, gy = -1.0f + (float)(2 * y + 1) / (float)Nf;
    float3 d;
    switch (s) {
        case 0: d = make_float3(1.f, -gy, -gx); break;
        case 1: d = make_float3(-1.f, -gy, gx); break;
        case 2: d = make_float3(gx, 1.f, gy); break;
        case 3: d = make_float3(gx, -1.f, -gy); break;
        case 4: d = make_float3(gx, -gy, 1.f); break;
        default: d = make_float3(-gx, -gy, -1.f); break;
    }
    const float n = fmaxf(sqrtf(d.x * d.x + d.y * d.y + d.z * d.z), 1e-12f);   // F.normalize
    const CubeTaps T = cube_taps(d.x / n, d.y / n, d.z / n, Nc);
    float3 v = cube_fetch(coarse_a, T);
    if (coarse_b) {
        const float3 u = cube_fetch(coarse_b, T);
        v.x += u.x; v.y += u.y; v.z += u.z;
    }
    float* o = out + 3 * (size_t)i;
    float3 r = make_float3(v.x * 0.25f, v.y * 0.25f, v.z * 0.25f);
    if (add) { r.x += add[3 * (size_t)i]; r.y += add[3 * (size_t)i + 1]; r.z += add[3 * (size_t)i + 2]; }
    if (accumulate) { r.x += o[0]; r.y += o[1]; r.z += o[2]; }
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}

// q = g / wsum, padded to 16 bytes, for every level in one launch (wsum == null: q = g, the cosine filter)
struct CmPrepArgs {
    int n;
    int begin[CM_MAX_SEG + 1];   // texel offsets of the segments in the launch
    const float* g[CM_MAX_SEG];
    const float* wsum[CM_MAX_SEG];
    float4* q[CM_MAX_SEG];
};

__global__ void __launch_bounds__(256) cm_prep_kernel(const __grid_constant__ CmPrepArgs a)
{
    pdl_enter();
    const int t = blockIdx.x * 256 + threadIdx.x;
    if (t >= a.begin[a.n]) return;
    int k = 0;
    for (int q = 1; q < a.n; ++q)
        if (t >= a.begin[q]) k = q;
    const int i = t - a.begin[k];
    const float* g = a.g[k] + 3 * (size_t)i;
    float3 v = make_float3(g[0], g[1], g[2]);
    if (a.wsum[k]) {
        const float w = a.wsum[k][i];
        v.x /= w; v.y /= w; v.z /= w;
    }
    a.q[k][i] = make_float4(v.x, v.y, v.z, 0.f);
}

__global__ void __launch_bounds__(256)
cm_add_kernel(const int n, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, const bool accumulate)
{
    pdl_enter();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    float r = a[i] + (b ? b[i] : 0.f);
    if (accumulate) r += out[i];
    out[i] = r;
}


// ---------------------------------------------------------------------------------------------
// The filters as stored sparse operators. Every weight w(V, L) depends only on (resolution, roughness, cutoff), never
// on the light itself, and the reference rebuilds the mips every training step: so the operator is evaluated ONCE,
// with the reference's arithmetic to the bit (including ndfGGX's double-precision division), into HBM, and a step
// streams it: out(V) = sum_L src(L) * W[V,L]. 180 GB of HBM3e make the ~1.5 GB (base_res 256, forward + transposed
// backward operator) a non-issue; the per-step cost drops from ~100 issue slots per pair to a fraction of a 16-byte load.
// Format (per filter), register-blocked by 4: a BLOCK is 4 consecutive output texels of one face row; their cones are the
// same cone shifted by a texel, so they share partner texels. The nonzeros of a block are RUNS of consecutive partner
// texels (the union, over the 4 texels, of the interval in which a face row of the light meets the cone — cone /\ face
// plane is convex), cut into pieces of at most G texels (G = the lanes that work on one block):
//   rowptr[block] .. rowptr[block+1] : the block's run records   recs[r] = (first partner texel | length << 21, offset)
//   W[offset + k] = float4 (w_0..w_3)                             : the 4 texels' weights for partner k of the run
// One partner gather (16 B) and one 16-byte weight load feed 12 FMAs. The backward operators have the same structure
// with blocks of light texels and weights w(V=j, L=i_b) / wsum(j) (the cone test is symmetric).
// ---------------------------------------------------------------------------------------------
constexpr int CM_RUN_SHIFT = 21;                       // 6*512*512 < 2^21 partner texels, run length <= 512 < 2^10
constexpr uint32_t CM_RUN_MASK = (1u << CM_RUN_SHIFT) - 1u;
constexpr int CM_B = 4;

struct CmBuildSeg {
    int N, kind;
    int lg;                 // log2 of the piece length G the runs are cut into
    int pad;
    float alpha_sqr, cutoff;
    const float4* table;
    const short4* bounds;
    uint2* counts;          // [blocks] (run pieces, weights) per block
    uint32_t* rowptr;       // [blocks+1]
    uint32_t* wptr;         // [blocks+1], in float4 units
    uint2* recs;
    float4* W;              // forward or backward operator
    float* wsum;            // forward: written (reference summation order); backward: read
};

// exact restatement of the reference's per-pair weight (c_src/cubemap.cu:173-178, 274-281), double division included
__device__ __forceinline__ float cm_spec_weight_exact(const float4 L, const float4 V, const float dotLV, const float alphaSqr)
{
    const float3 H = cm_safe_normalize(L.x + V.x, L.y + V.y, L.z + V.z);
    const float wiDotN = fmaxf(dotLV, 0.0f);
    const float VdotH = fmaxf(V.x * H.x + V.y * H.y + V.z * H.z, 0.0f);
    const float c = fminf(fmaxf(VdotH, 0.0f), 1.0f);
    const float d = (c * alphaSqr - c) * c + 1.0f;
    const float ndf = (float)((double)alphaSqr / ((double)(d * d) * 3.14159265358979323846));
    return wiDotN * ndf * L.w / 4.0f;
}

__device__ __forceinline__ bool cm_in_box(const short4 b, const int x, const int y)
{
    return x >= b.x && x <= b.y && y >= b.z && y <= b.w;
}

// One thread per block of 4 texels, serial over the block's cone (prepare-time only).
// PASS 0: count run pieces / weights. PASS 1: forward operator + run records + wsum. PASS 2: backward operator.
template <int PASS>
__global__ void __launch_bounds__(128) cm_operator_kernel(const CmBuildSeg S)
{
    const int N = S.N, n = 6 * N * N, G = 1 << S.lg;
    const int blk = blockIdx.x * 128 + threadIdx.x;
    if (blk >= n / CM_B) return;
    const int i0 = blk * CM_B;
    float4 tv[CM_B];
#pragma unroll
    for (int b = 0; b < CM_B; ++b) tv[b] = S.table[i0 + b];
    uint32_t nrun = 0, nw = 0, rp = 0, wp = 0;
    if (PASS > 0) { rp = S.rowptr[blk]; wp = S.wptr[blk]; }
    float ws[CM_B] = {0.f, 0.f, 0.f, 0.f};
    const float cutoff = S.cutoff, alphaSqr = S.alpha_sqr;
    for (int s = 0; s < 6; ++s) {
        short4 bb[CM_B];
        int ux0 = N, ux1 = -1, uy0 = N, uy1 = -1;
        if (S.kind == 1) { ux0 = 0; ux1 = N - 1; uy0 = 0; uy1 = N - 1; }
        else {
#pragma unroll
            for (int b = 0; b < CM_B; ++b) {
                bb[b] = S.bounds[(size_t)(i0 + b) * 6 + s];
                if (bb[b].x <= bb[b].y) {
                    ux0 = min(ux0, (int)bb[b].x); ux1 = max(ux1, (int)bb[b].y);
                    uy0 = min(uy0, (int)bb[b].z); uy1 = max(uy1, (int)bb[b].w);
                }
            }
        }
        if (ux0 > ux1) continue;
        for (int y = uy0; y <= uy1; ++y) {
            const int row = (s * N + y) * N;
            int x0 = -1, x1 = -1;
            if (S.kind == 1) { x0 = 0; x1 = N - 1; }
            else {
                for (int x = ux0; x <= ux1; ++x) {
                    const float4 tl = S.table[row + x];
                    bool any = false;
#pragma unroll
                    for (int b = 0; b < CM_B; ++b)
                        any |= cm_in_box(bb[b], x, y) && (tl.x * tv[b].x + tl.y * tv[b].y + tl.z * tv[b].z >= cutoff);
                    if (any) { if (x0 < 0) x0 = x; x1 = x; }
                }
            }
            if (x0 < 0) continue;
            const int len = x1 - x0 + 1;
            if (PASS == 0) { nrun += (uint32_t)((len + G - 1) >> S.lg); nw += (uint32_t)len; continue; }
            if (PASS == 1)
                for (int p = 0; p < len; p += G)
                    S.recs[rp++] = make_uint2((uint32_t)(row + x0 + p) | ((uint32_t)min(G, len - p) << CM_RUN_SHIFT), wp + (uint32_t)p);
            for (int x = x0; x <= x1; ++x) {
                const int j = row + x;
                const float4 tl = S.table[j];
                float w[CM_B];
#pragma unroll
                for (int b = 0; b < CM_B; ++b) {
                    w[b] = 0.f;
                    if (S.kind == 1) {
                        // cosine filter: forward weighs by the partner's solid angle, backward by the block texel's
                        const float costheta = fminf(fmaxf(tv[b].x * tl.x + tv[b].y * tl.y + tv[b].z * tl.z, 0.0f), 0.999f);
                        w[b] = costheta * (PASS == 1 ? tl.w : tv[b].w) / 3.141592f;
                    } else {
                        const float dotLV = tl.x * tv[b].x + tl.y * tv[b].y + tl.z * tv[b].z;
                        if (cm_in_box(bb[b], x, y) && dotLV >= cutoff) {
                            if (PASS == 1) { w[b] = cm_spec_weight_exact(tl, tv[b], dotLV, alphaSqr); ws[b] += w[b]; }
                            else w[b] = cm_spec_weight_exact(tv[b], tl, dotLV, alphaSqr) / S.wsum[j];
                        }
                    }
                }
                S.W[wp++] = make_float4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    if (PASS == 0) S.counts[blk] = make_uint2(nrun, nw);
    if (PASS == 1 && S.kind == 0) {
#pragma unroll
        for (int b = 0; b < CM_B; ++b) S.wsum[i0 + b] = ws[b];
    }
}

// exclusive prefix sums of the per-texel counts (one CTA; prepare-time only)
__global__ void __launch_bounds__(1024) cm_scan_kernel(const int T, const uint2* __restrict__ counts, uint32_t* __restrict__ rowptr,
                                                       uint32_t* __restrict__ wptr, unsigned long long* __restrict__ totals)
{
    __shared__ unsigned long long s_a[32], s_b[32];
    __shared__ unsigned long long carry_a, carry_b;
    if (threadIdx.x == 0) { carry_a = 0; carry_b = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < T; base += 1024) {
        const int i = base + threadIdx.x;
        const uint2 c = i < T ? counts[i] : make_uint2(0, 0);
        unsigned long long a = c.x, b = c.y;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long ta = __shfl_up_sync(0xffffffffu, a, off), tb = __shfl_up_sync(0xffffffffu, b, off);
            if (lane >= off) { a += ta; b += tb; }
        }
        if (lane == 31) { s_a[warp] = a; s_b[warp] = b; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long wa = s_a[lane], wb = s_b[lane];
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long ta = __shfl_up_sync(0xffffffffu, wa, off), tb = __shfl_up_sync(0xffffffffu, wb, off);
                if (lane >= off) { wa += ta; wb += tb; }
            }
            s_a[lane] = wa; s_b[lane] = wb;
        }
        __syncthreads();
        const unsigned long long pa = carry_a + (warp ? s_a[warp - 1] : 0), pb = carry_b + (warp ? s_b[warp - 1] : 0);
        if (i < T) { rowptr[i] = (uint32_t)(pa + a - c.x); wptr[i] = (uint32_t)(pb + b - c.y); }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_a = pa + a; carry_b = pb + b; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rowptr[T] = (uint32_t)carry_a; wptr[T] = (uint32_t)carry_b;
        totals[0] = carry_a; totals[1] = carry_b;
    }
}

struct CmSpSeg {
    int cta_begin, NB, lg, pad;   // NB blocks, 2^lg lanes per block (= the piece length the runs were cut into)
    const uint32_t* rowptr;
    const uint32_t* wptr;  // [NB+1] first weight entry of each block
    const uint2* recs;
    const float4* W;
    const float4* src;     // padded texels
    float* dst;            // [4*NB][3]
    const float* wsum;     // divide the result by it (forward GGX levels) or null
};
struct CmSpPlan {
    int nseg, pad;
    CmSpSeg seg[CM_MAX_SEG];
};

// dst(V_b) = sum over the runs of the block of src(partner) * W_b. G = 2^lg lanes per block; lane k of a group takes
// partner k of every run piece (pieces are at most G long), two pieces per iteration: per lane and iteration two
// 16-byte weight loads (the HBM stream: 1 KB per warp in flight) and two 16-byte partner gathers (L1/L2) feed 24 FMAs.
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) cm_sparse_kernel(const __grid_constant__ CmSpPlan plan)
{
    pdl_enter();
    int k = 0;
    for (int q = 1; q < plan.nseg; ++q)
        if ((int)blockIdx.x >= plan.seg[q].cta_begin) k = q;
    const CmSpSeg& S = plan.seg[k];
    const int lg = S.lg, G = 1 << lg;
    const int t = ((int)blockIdx.x - S.cta_begin) * 256 + (int)threadIdx.x;
    int blk = t >> lg;
    const int sub = t & (G - 1);
    const bool live = blk < S.NB;
    if (!live) blk = S.NB - 1;   // keep the lane in the shuffles below
    if (threadIdx.x == 0) {
        // the CTA's blocks are consecutive, so are their weights: one bulk prefetch pulls the whole slab (7..120 KB)
        // into L2 while the warps start on it — the copy engine, not per-warp loads in flight, covers the HBM latency
        const int b0 = blk, b1 = min(b0 + (256 >> lg), S.NB);
        const uint32_t w0 = S.wptr[b0], w1 = S.wptr[b1];
        if (w1 > w0) bulk_prefetch_l2(S.W + w0, (w1 - w0) * 16u);
    }
    const uint2* __restrict__ recs = S.recs;
    const float4* __restrict__ W = S.W;
    const float4* __restrict__ src = S.src;
    const uint32_t r0 = S.rowptr[blk], r1 = S.rowptr[blk + 1];
    float acc[CM_B][3];
#pragma unroll
    for (int b = 0; b < CM_B; ++b) acc[b][0] = acc[b][1] = acc[b][2] = 0.f;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint2 z2 = make_uint2(0u, 0u);
    // Run records: lane k of the group loads record k of a chunk of G records (one coalesced load, the NEXT chunk's issued
    // a whole chunk ahead) and the group hands them round by shuffle — a per-piece record load put one full memory
    // latency in front of every weight load (66 % of the stall samples of the first version of this kernel).
    uint2 mine = r0 + sub < r1 ? recs[r0 + sub] : z2;
    // a warp may hold two or more groups (G < 32) with different record counts: trip counts are made warp-uniform so
    // that every lane reaches every shuffle
    const int nall = (int)__reduce_max_sync(0xffffffffu, r1 - r0);
    for (int c = 0; c < nall; c += G) {
        const uint32_t rbase = r0 + (uint32_t)c;
        const uint2 ahead = rbase + G + sub < r1 ? recs[rbase + G + sub] : z2;
        const int nrec = min(max((int)(r1 - r0) - c, 0), G);
        const int ntrip = min(nall - c, G);
        for (int p = 0; p < ntrip; p += U) {
            float4 w[U], s[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t rx = __shfl_sync(0xffffffffu, mine.x, p + u, G);    // p + u >= nrec reads a zero record of
                const uint32_t ry = __shfl_sync(0xffffffffu, mine.y, p + u, G);    // this chunk or wraps to a real one:
                const bool in = p + u < nrec && sub < (int)(rx >> CM_RUN_SHIFT);   // masked here
                w[u] = in ? W[ry + sub] : z4;
                s[u] = in ? src[(rx & CM_RUN_MASK) + sub] : z4;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc[0][0] += s[u].x * w[u].x; acc[0][1] += s[u].y * w[u].x; acc[0][2] += s[u].z * w[u].x;
                acc[1][0] += s[u].x * w[u].y; acc[1][1] += s[u].y * w[u].y; acc[1][2] += s[u].z * w[u].y;
                acc[2][0] += s[u].x * w[u].z; acc[2][1] += s[u].y * w[u].z; acc[2][2] += s[u].z * w[u].z;
                acc[3][0] += s[u].x * w[u].w; acc[3][1] += s[u].y * w[u].w; acc[3][2] += s[u].z * w[u].w;
            }
        }
        mine = ahead;
    }
    for (int off = G >> 1; off > 0; off >>= 1) {
#pragma unroll
        for (int b = 0; b < CM_B; ++b) {
            acc[b][0] += __shfl_xor_sync(0xffffffffu, acc[b][0], off);
            acc[b][1] += __shfl_xor_sync(0xffffffffu, acc[b][1], off);
            acc[b][2] += __shfl_xor_sync(0xffffffffu, acc[b][2], off);
        }
    }
    if (live && sub < CM_B) {
        float c0 = acc[0][0], c1 = acc[0][1], c2 = acc[0][2];
        if (sub == 1) { c0 = acc[1][0]; c1 = acc[1][1]; c2 = acc[1][2]; }
        if (sub == 2) { c0 = acc[2][0]; c1 = acc[2][1]; c2 = acc[2][2]; }
        if (sub == 3) { c0 = acc[3][0]; c1 = acc[3][1]; c2 = acc[3][2]; }
        const int i = blk * CM_B + sub;
        float* o = S.dst + 3 * (size_t)i;
        if (S.wsum) {
            const float ws = S.wsum[i];
            o[0] = c0 / ws; o[1] = c1 / ws; o[2] = c2 / ws;
        } else {
            o[0] = c0; o[1] = c1; o[2] = c2;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Env-map smoothness prior (train.py:406-420): envmap = seamless bilinear lookup of the BASE cubemap along a fixed
// [EH,EW] lat-long grid of directions; loss = mean((env[1:] - env[:-1])^2) + mean((env[:,1:] - env[:,:-1])^2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
env_lookup_kernel(const int N, const float* __restrict__ base, const float* __restrict__ dirs, const int n, float* __restrict__ env)
{
    pdl_enter();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const CubeTaps T = cube_taps(dirs[3 * (size_t)i], dirs[3 * (size_t)i + 1], dirs[3 * (size_t)i + 2], N);
    const float3 v = cube_fetch(base, T);
    env[3 * (size_t)i] = v.x; env[3 * (size_t)i + 1] = v.y; env[3 * (size_t)i + 2] = v.z;
}

// per pixel: its two forward differences (loss partials, one per CTA) and, if grad_base, d loss / d env(pixel) from the
// (up to) four differences that touch it, scattered through the pixel's four bilinear taps
__global__ void __launch_bounds__(256)
env_tv_kernel(const int N, const int EH, const int EW, const float* __restrict__ env, const float* __restrict__ dirs,
              const float scale /*weight * loss_scale*/, float* __restrict__ partials, float* __restrict__ grad_base)
{
    pdl_enter();
    __shared__ float s_red[8];
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = EH * EW;
    float part = 0.f;
    if (i < n) {
        const int x = i % EW, y = i / EW;
        const float kh = 1.0f / (3.0f * (float)(EH - 1) * (float)EW), kw = 1.0f / (3.0f * (float)EH * (float)(EW - 1));
        const float* e = env + 3 * (size_t)i;
        float g[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = e[c];
            if (y + 1 < EH) { const float d = e[3 * EW + c] - v; part += d * d * kh; g[c] -= 2.f * kh * d; }
            if (x + 1 < EW) { const float d = e[3 + c] - v; part += d * d * kw; g[c] -= 2.f * kw * d; }
            if (y > 0) g[c] += 2.f * kh * (v - e[c - 3 * EW]);
            if (x > 0) g[c] += 2.f * kw * (v - e[c - 3]);
        }
        if (grad_base) {
            const CubeTaps T = cube_taps(dirs[3 * (size_t)i], dirs[3 * (size_t)i + 1], dirs[3 * (size_t)i + 2], N);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (T.idx[k] >= 0 && T.w[k] != 0.f) {
                    float* o = grad_base + 3 * (size_t)T.idx[k];
                    const float w = T.w[k] * scale;
                    red_add_f32(o, g[0] * w); red_add_f32(o + 1, g[1] * w); red_add_f32(o + 2, g[2] * w);
                }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_red[w];
        partials[blockIdx.x] = t;
    }
}

// fixed-order sum of the per-CTA partials (deterministic); loss_out = [accumulate ? loss_out : 0] + scale * sum
__global__ void __launch_bounds__(256)
env_tv_sum_kernel(const int nblk, const float* __restrict__ partials, const float scale, float* __restrict__ loss_out,
                  const bool accumulate)
{
    pdl_enter();
    __shared__ float s_red[8];
    float a = 0.f;
    for (int i = threadIdx.x; i < nblk; i += 256) a += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_red[w];
        loss_out[0] = (accumulate ? loss_out[0] : 0.f) + scale * t;
    }
}

}  // namespace gigs

using namespace gigs;

namespace {

template <bool BACKWARD, bool PACKED>
int cm_launch_plan(const CmPlan& plan, int ctas, cudaStream_t st)
{
    GIGS_CUDA(launch_k(cm_filter_kernel<BACKWARD, PACKED>, dim3(ctas), dim3(256), (size_t)(0), st, plan));
    GIGS_LAUNCH_CHECK("cm_filter_kernel");
    return 0;
}

bool cm_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

inline char* at(void* ws, uint64_t off) { return reinterpret_cast<char*>(ws) + off; }

}  // namespace

// ---------------------------------------------------------------------------------------------
// latlong_to_cubemap (relight.py:92-112, render.py:64-84): every cube texel looks its direction up in an
// equirectangular HDR map, (u, v) = (atan2(x, -z) / 2pi + 0.5, acos(clamp(y)) / pi), bilinear, both axes wrapping
// (the reference samples with nvdiffrast's dr.texture(filter_mode="linear"), whose default boundary mode is "wrap":
// texel centres at (i + 0.5) / size).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cm_latlong_kernel(const int EH, const int EW, const int Cn, const float* __restrict__ env, const int R,
                  float* __restrict__ cube)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= 6 * R * R) return;
    const int x = i % R, y = (i / R) % R, s = i / (R * R);
    const float gx = -1.0f + (float)(2 * x + 1) / (float)R, gy = -1.0f + (float)(2 * y + 1) / (float)R;
    float3 d;
    switch (s) {
        case 0: d = make_float3(1.f, -gy, -gx); break;
        case 1: d = make_float3(-1.f, -gy, gx); break;
        case 2: d = make_float3(gx, 1.f, gy); break;
        case 3: d = make_float3(gx, -1.f, -gy); break;
        case 4: d = make_float3(gx, -gy, 1.f); break;
        default: d = make_float3(-gx, -gy, -1.f); break;
    }
    const float n = fmaxf(torch_norm_inner3(d.x, d.y, d.z), 1e-12f);   // F.normalize(dim=-1)
    d = make_float3(d.x / n, d.y / n, d.z / n);
    const float tu = atan2f(d.x, -d.z) / (2.0f * 3.14159265358979323846f) + 0.5f;
    const float tv = acosf(fminf(fmaxf(d.y, -1.0f), 1.0f)) / 3.14159265358979323846f;
    const float fxp = tu * (float)EW - 0.5f, fyp = tv * (float)EH - 0.5f;
    const float x0f = floorf(fxp), y0f = floorf(fyp);
    const float ax = fxp - x0f, ay = fyp - y0f;
    int x0 = (int)x0f % EW, y0 = (int)y0f % EH;
    if (x0 < 0) x0 += EW;
    if (y0 < 0) y0 += EH;
    const int x1 = (x0 + 1 == EW) ? 0 : x0 + 1, y1 = (y0 + 1 == EH) ? 0 : y0 + 1;
    const float w00 = (1.f - ax) * (1.f - ay), w10 = ax * (1.f - ay), w01 = (1.f - ax) * ay, w11 = ax * ay;
    const float* p00 = env + ((size_t)y0 * EW + x0) * Cn;
    const float* p10 = env + ((size_t)y0 * EW + x1) * Cn;
    const float* p01 = env + ((size_t)y1 * EW + x0) * Cn;
    const float* p11 = env + ((size_t)y1 * EW + x1) * Cn;
    float* o = cube + (size_t)i * Cn;
    for (int c = 0; c < Cn; ++c) o[c] = p00[c] * w00 + p10[c] * w10 + p01[c] * w01 + p11[c] * w11;
}

extern "C" {

int gigs_latlong_to_cubemap(int32_t env_h, int32_t env_w, int32_t channels, const float* env, int32_t res, float* cube,
                             void* stream)
{
    if (env_h <= 0 || env_w <= 0 || channels <= 0 || res <= 0 || res > 4096 || !env || !cube) { set_error("gigs_latlong_to_cubemap: bad arguments"); return -1; }
    const int n = 6 * res * res;
    ProfScope ps(ST_CUBEMAP, (cudaStream_t)stream);
    cm_latlong_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(env_h, env_w, channels, env, res, cube);
    GIGS_LAUNCH_CHECK("cm_latlong_kernel");
    return 0;
}

int gigs_cubemap_table(int32_t res, float* table, void* stream)
{
    if (res <= 0 || res > 4096 || !table) { set_error("gigs_cubemap_table: bad arguments"); return -1; }
    const int n = 6 * res * res;
    cm_table_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(res, (float4*)table);
    GIGS_LAUNCH_CHECK("cm_table_kernel");
    return 0;
}

int gigs_specular_bounds(int32_t res, float costheta_cutoff, const float* table, int16_t* bounds, void* stream)
{
    if (res <= 0 || res > 4096 || !table || !bounds) { set_error("gigs_specular_bounds: bad arguments"); return -1; }
    const int n = 6 * res * res;
    cm_bounds_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(res, costheta_cutoff, (const float4*)table, (short4*)bounds);
    GIGS_LAUNCH_CHECK("cm_bounds_kernel");
    return 0;
}

int gigs_cubemap_mip_forward(int32_t res_out, const float* in, float* out, void* stream)
{
    if (res_out <= 0 || !in || !out) { set_error("gigs_cubemap_mip_forward: bad arguments"); return -1; }
    const int n = 6 * res_out * res_out;
    GIGS_CUDA(launch_k(cm_mip_forward_kernel, dim3((n + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, res_out, in, out));
    GIGS_LAUNCH_CHECK("cm_mip_forward_kernel");
    return 0;
}

int gigs_cubemap_mip_backward(int32_t res_coarse, const float* grad_coarse, float* grad_fine, int32_t accumulate, void* stream)
{
    if (res_coarse <= 0 || !grad_coarse || !grad_fine) { set_error("gigs_cubemap_mip_backward: bad arguments"); return -1; }
    const int n = 6 * 4 * res_coarse * res_coarse;
    GIGS_CUDA(launch_k(cm_mip_backward_kernel, dim3((n + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, res_coarse, grad_coarse, nullptr, nullptr, grad_fine,
                                                                              accumulate != 0));
    GIGS_LAUNCH_CHECK("cm_mip_backward_kernel");
    return 0;
}

static int cm_single(bool backward, int kind, int32_t res, const float* table, const int16_t* bounds, float roughness,
                     float cutoff, const float* src, float* dst, float* wsum, void* stream)
{
    CmPlan plan{};
    int ctas = 0;
    cm_plan_add(plan, ctas, res, kind, roughness, cutoff, table, bounds, src, dst, wsum);
    ProfScope ps(ST_CUBEMAP, (cudaStream_t)stream);
    return backward ? cm_launch_plan<true, false>(plan, ctas, (cudaStream_t)stream)
                    : cm_launch_plan<false, false>(plan, ctas, (cudaStream_t)stream);
}

int gigs_diffuse_cubemap_forward(int32_t res, const float* table, const float* cubemap, float* out, void* stream)
{
    if (res <= 0 || !table || !cubemap || !out) { set_error("gigs_diffuse_cubemap_forward: bad arguments"); return -1; }
    return cm_single(false, 1, res, table, nullptr, 1.f, 0.f, cubemap, out, nullptr, stream);
}

int gigs_diffuse_cubemap_backward(int32_t res, const float* table, const float* grad_out, float* grad_in, void* stream)
{
    if (res <= 0 || !table || !grad_out || !grad_in) { set_error("gigs_diffuse_cubemap_backward: bad arguments"); return -1; }
    return cm_single(true, 1, res, table, nullptr, 1.f, 0.f, grad_out, grad_in, nullptr, stream);
}

int gigs_specular_cubemap_forward(int32_t res, const float* table, const int16_t* bounds, float roughness, float cutoff,
                                  const float* cubemap, float* out, float* wsum, void* stream)
{
    if (res <= 0 || !table || !bounds || !cubemap || !out || !wsum) { set_error("gigs_specular_cubemap_forward: bad arguments"); return -1; }
    return cm_single(false, 0, res, table, bounds, roughness, cutoff, cubemap, out, wsum, stream);
}

int gigs_specular_cubemap_backward(int32_t res, const float* table, const int16_t* bounds, float roughness, float cutoff,
                                   const float* grad_out, const float* wsum, float* grad_in, void* stream)
{
    if (res <= 0 || !table || !bounds || !grad_out || !wsum || !grad_in) { set_error("gigs_specular_cubemap_backward: bad arguments"); return -1; }
    return cm_single(true, 0, res, table, bounds, roughness, cutoff, grad_out, grad_in, const_cast<float*>(wsum), stream);
}

// ---- the whole light: CubemapLight.build_mips and its backward --------------------------------------------------
// filter index f: 0..n_levels-1 = the GGX levels, n_levels = the cosine filter (on the coarsest level)
static inline int cm_filter_res(const GigsLightLayout* L, int f) { return L->res[f < L->n_levels ? f : L->n_levels - 1]; }

int gigs_light_layout(int32_t base_res, int32_t min_res, GigsLightLayout* L)
{
    if (!L || !cm_pow2(base_res) || !cm_pow2(min_res) || min_res < 16 || base_res < min_res || base_res > 512) {
        set_error("gigs_light_layout: base_res and min_res must be powers of two, 16 <= min_res <= base_res <= 512");
        return -1;
    }
    std::memset(L, 0, sizeof(*L));
    int n = 0;
    for (int r = base_res; r >= min_res; r >>= 1) {
        if (n >= GIGS_MAX_LIGHT_LEVELS) { set_error("gigs_light_layout: more than %d levels", GIGS_MAX_LIGHT_LEVELS); return -1; }
        L->res[n++] = r;
    }
    if (n == 2) { set_error("gigs_light_layout: two levels divide by zero in the roughness schedule (pbr/light.py:166)"); return -1; }
    L->n_levels = n;
    // pbr/light.py:165-170 (python doubles, then passed down as float)
    for (int i = 0; i + 1 < n; ++i) L->roughness[i] = (float)(((double)i / (double)(n - 2)) * (0.5 - 0.08) + 0.08);
    L->roughness[n - 1] = 1.0f;
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { const uint64_t o = off; off += (bytes + 255) & ~uint64_t(255); return o; };
    for (int i = 0; i < n; ++i) {
        const uint64_t t = (uint64_t)6 * L->res[i] * L->res[i];
        L->table[i] = take(16 * t);
        L->bounds[i] = take(48 * t);
        L->chain[i] = take(16 * t);
        L->spec[i] = take(12 * t);
        L->wsum[i] = take(4 * t);
        L->gq[i] = take(16 * t);
        L->g_chain[i] = take(12 * t);
    }
    for (int f = 0; f <= n; ++f) {
        const uint64_t t = (uint64_t)6 * cm_filter_res(L, f) * cm_filter_res(L, f);
        L->rowptr[f] = take(4 * (t + 1));
        L->wptr[f] = take(4 * (t + 1));
    }
    L->counts = take(8 * (uint64_t)6 * base_res * base_res);
    L->totals = take(16 * (GIGS_MAX_LIGHT_LEVELS + 1));
    const uint64_t tl = (uint64_t)6 * min_res * min_res;
    L->diffuse = take(12 * tl);
    L->gq_diffuse = take(16 * tl);
    L->g_diffuse_in = take(12 * tl);
    // the texture gradients the shading backward accumulates into: one contiguous span, cleared by one memset
    L->grad_begin = off;
    for (int i = 0; i < n; ++i) L->g_spec[i] = take(12 * (uint64_t)6 * L->res[i] * L->res[i]);
    L->g_diffuse = take(12 * tl);
    L->grad_bytes = off - L->grad_begin;
    L->total_bytes = off;
    return 0;
}

static CmBuildSeg cm_build_seg(const GigsLightLayout* L, void* ws, void* weights, int f, int pass)
{
    const int n = L->n_levels, lvl = f < n ? f : n - 1;
    CmBuildSeg S{};
    S.N = L->res[lvl];
    S.kind = f < n ? 0 : 1;
    S.lg = L->lanes_log2[f];
    const float alpha = L->roughness[lvl] * L->roughness[lvl];
    S.alpha_sqr = alpha * alpha;
    S.cutoff = L->cutoff[lvl];
    S.table = (const float4*)at(ws, L->table[lvl]);
    S.bounds = (const short4*)at(ws, L->bounds[lvl]);
    S.counts = (uint2*)at(ws, L->counts);
    S.rowptr = (uint32_t*)at(ws, L->rowptr[f]);
    S.wptr = (uint32_t*)at(ws, L->wptr[f]);
    if (weights) {
        S.recs = (uint2*)at(weights, L->w_rows[f]);
        S.W = (float4*)at(weights, pass == 2 ? L->w_bwd[f] : L->w_fwd[f]);
    }
    S.wsum = (float*)at(ws, L->wsum[lvl]);
    return S;
}

int gigs_light_prepare(GigsLightLayout* L, void* ws, void* stream)
{
    if (!L || !ws || L->n_levels <= 0) { set_error("gigs_light_prepare: bad arguments"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int n = L->n_levels;
    for (int i = 0; i < n; ++i) {
        if (!(L->cutoff[i] > -1.f && L->cutoff[i] <= 1.f)) { set_error("gigs_light_prepare: cutoff[%d] not set", i); return -1; }
        int rc = gigs_cubemap_table(L->res[i], (float*)at(ws, L->table[i]), stream);
        if (rc) return rc;
        rc = gigs_specular_bounds(L->res[i], L->cutoff[i], (const float*)at(ws, L->table[i]), (int16_t*)at(ws, L->bounds[i]), stream);
        if (rc) return rc;
    }
    // structure of the stored operators. Round 0 counts uncut runs to pick the lanes per block (= piece length) of each
    // filter from its average run; round 1 counts the pieces and takes the prefix sums.
    unsigned long long tot[2 * (GIGS_MAX_LIGHT_LEVELS + 1)];
    for (int round = 0; round < 2; ++round) {
        for (int f = 0; f <= n; ++f) {
            if (round == 0) L->lanes_log2[f] = 30;    // pieces never cut
            const CmBuildSeg S = cm_build_seg(L, ws, nullptr, f, 0);
            const int NB = 6 * S.N * S.N / CM_B;
            cm_operator_kernel<0><<<(NB + 127) / 128, 128, 0, st>>>(S);
            GIGS_LAUNCH_CHECK("cm_operator_kernel<0>");
            cm_scan_kernel<<<1, 1024, 0, st>>>(NB, S.counts, S.rowptr, S.wptr, (unsigned long long*)at(ws, L->totals) + 2 * f);
            GIGS_LAUNCH_CHECK("cm_scan_kernel");
        }
        GIGS_CUDA(cudaMemcpyAsync(tot, at(ws, L->totals), sizeof(unsigned long long) * 2 * (n + 1), cudaMemcpyDeviceToHost, st));
        GIGS_CUDA(cudaStreamSynchronize(st));
        if (round == 0)
            for (int f = 0; f <= n; ++f) {
                const double avg = tot[2 * f] ? (double)tot[2 * f + 1] / (double)tot[2 * f] : 1.0;
                int lg = 2;
                while (lg < 5 && (double)(1 << lg) < avg) ++lg;
                L->lanes_log2[f] = lg;
            }
    }
    uint64_t off = 0;
    auto take = [&](uint64_t bytes) { const uint64_t o = off; off += (bytes + 255) & ~uint64_t(255); return o; };
    for (int f = 0; f <= n; ++f) {
        L->n_runs[f] = tot[2 * f];
        L->n_weights[f] = tot[2 * f + 1];
        if (L->n_weights[f] >= (1ull << 32)) { set_error("gigs_light_prepare: operator %d has %llu weight entries (>= 2^32)", f, tot[2 * f + 1]); return -1; }
        L->w_rows[f] = take(8 * L->n_runs[f]);
        L->w_fwd[f] = take(16 * L->n_weights[f]);
        L->w_bwd[f] = take(16 * L->n_weights[f]);
    }
    L->weights_bytes = off;
    GIGS_CUDA(cudaMemsetAsync(at(ws, L->grad_begin), 0, L->grad_bytes, st));
    return 0;
}

int gigs_light_weights(const GigsLightLayout* L, void* ws, void* weights, void* stream)
{
    if (!L || !ws || !weights || L->n_levels <= 0 || L->weights_bytes == 0) { set_error("gigs_light_weights: bad arguments (prepare first)"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    for (int pass = 1; pass <= 2; ++pass)
        for (int f = 0; f <= L->n_levels; ++f) {
            const CmBuildSeg S = cm_build_seg(L, ws, weights, f, pass);
            const int NB = 6 * S.N * S.N / CM_B;
            if (pass == 1) cm_operator_kernel<1><<<(NB + 127) / 128, 128, 0, st>>>(S);
            else cm_operator_kernel<2><<<(NB + 127) / 128, 128, 0, st>>>(S);
            GIGS_LAUNCH_CHECK("cm_operator_kernel");
        }
    return 0;
}

// all filters as one launch of the stored operators; longest blocks (coarse levels) first
static int cm_sparse_launch(const GigsLightLayout* L, void* ws, void* weights, bool backward, cudaStream_t st)
{
    const int n = L->n_levels;
    CmSpPlan plan{};
    int ctas = 0;
    int order[GIGS_MAX_LIGHT_LEVELS + 1];
    for (int f = 0; f <= n; ++f) order[f] = f;
    auto per_block = [&](int f) { return (double)L->n_runs[f] / (6.0 * cm_filter_res(L, f) * cm_filter_res(L, f)); };
    std::sort(order, order + n + 1, [&](int a, int b) { return per_block(a) > per_block(b); });
    for (int q = 0; q <= n; ++q) {
        const int f = order[q], lvl = f < n ? f : n - 1;
        CmSpSeg& S = plan.seg[plan.nseg++];
        S.cta_begin = ctas;
        S.NB = 6 * L->res[lvl] * L->res[lvl] / CM_B;
        S.lg = L->lanes_log2[f];
        S.rowptr = (const uint32_t*)at(ws, L->rowptr[f]);
        S.wptr = (const uint32_t*)at(ws, L->wptr[f]);
        S.recs = (const uint2*)at(weights, L->w_rows[f]);
        S.W = (const float4*)at(weights, backward ? L->w_bwd[f] : L->w_fwd[f]);
        if (!backward) {
            S.src = (const float4*)at(ws, L->chain[lvl]);
            S.dst = (float*)at(ws, f < n ? L->spec[f] : L->diffuse);
            S.wsum = f < n ? (const float*)at(ws, L->wsum[f]) : nullptr;
        } else {
            S.src = (const float4*)at(ws, f < n ? L->gq[f] : L->gq_diffuse);
            S.dst = (float*)at(ws, f < n ? L->g_chain[f] : L->g_diffuse_in);
            S.wsum = nullptr;
        }
        const int bpc = 256 >> S.lg;
        ctas += (S.NB + bpc - 1) / bpc;
    }
    // 2 pieces per iteration, 4 CTAs per SM: measured equal (within 3 %) to 3-4 pieces at 3 CTAs and better than 1 piece
    // at 6 CTAs or 2 at 5; with the weight loads stubbed out the kernel still takes 76 % of its time, so it is bound by
    // the gathers and the ~30 instructions per piece, not by the HBM stream (3.6 TB/s)
    GIGS_CUDA(launch_k(cm_sparse_kernel<2, 4>, dim3(ctas), dim3(256), (size_t)(0), st, plan));
    GIGS_LAUNCH_CHECK("cm_sparse_kernel");
    return 0;
}

// compute-on-the-fly variant of the same launch (weights == NULL: no stored operators)
static int cm_compute_launch(const GigsLightLayout* L, void* ws, bool backward, cudaStream_t st)
{
    const int n = L->n_levels;
    CmPlan plan{};
    int ctas = 0;
    int order[GIGS_MAX_LIGHT_LEVELS];
    for (int i = 0; i < n; ++i) order[i] = i;
    auto per_lane = [&](int i) { return 6.0 * L->res[i] * L->res[i] * (1.0 - L->cutoff[i]) / (double)(1 << cm_lanes_log2(L->res[i], 0, L->cutoff[i])); };
    std::sort(order, order + n, [&](int a, int b) { return per_lane(a) > per_lane(b); });
    cm_plan_add(plan, ctas, L->res[n - 1], 1, 1.f, 0.f, (const float*)at(ws, L->table[n - 1]), nullptr,
                at(ws, backward ? L->gq_diffuse : L->chain[n - 1]), (float*)at(ws, backward ? L->g_diffuse_in : L->diffuse), nullptr);
    for (int q = 0; q < n; ++q) {
        const int i = order[q];
        cm_plan_add(plan, ctas, L->res[i], 0, L->roughness[i], L->cutoff[i], (const float*)at(ws, L->table[i]),
                    (const int16_t*)at(ws, L->bounds[i]), at(ws, backward ? L->gq[i] : L->chain[i]),
                    (float*)at(ws, backward ? L->g_chain[i] : L->spec[i]), backward ? nullptr : (float*)at(ws, L->wsum[i]));
    }
    return backward ? cm_launch_plan<true, true>(plan, ctas, st) : cm_launch_plan<false, true>(plan, ctas, st);
}

int gigs_light_build(const GigsLightLayout* L, const float* base, void* ws, const void* weights, void* stream)
{
    if (!L || !ws || !base || L->n_levels <= 0) { set_error("gigs_light_build: bad arguments"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(ST_CUBEMAP, st);
    const int n = L->n_levels;
    // mip chain, 4 halvings per launch
    for (int first = 0; first == 0 || first < n - 1; first += 4) {
        CmChainArgs a{};
        a.R = L->res[first];
        a.nh = std::min(4, n - 1 - first);
        a.src = first == 0 ? base : (const float*)at(ws, L->chain[first]);
        a.pack0 = first == 0 ? (float4*)at(ws, L->chain[0]) : nullptr;
        for (int h = 0; h < a.nh; ++h) a.out[h] = (float4*)at(ws, L->chain[first + 1 + h]);
        const dim3 grid(a.R / 16, a.R / 16, 6);
        if (first == 0) GIGS_CUDA(launch_k(cm_chain_kernel<false>, dim3(grid), dim3(256), (size_t)(0), st, a));
        else GIGS_CUDA(launch_k(cm_chain_kernel<true>, dim3(grid), dim3(256), (size_t)(0), st, a));
        GIGS_LAUNCH_CHECK("cm_chain_kernel");
    }
    return weights ? cm_sparse_launch(L, ws, const_cast<void*>(weights), false, st) : cm_compute_launch(L, ws, false, st);
}

int gigs_light_backward(const GigsLightLayout* L, void* ws, const void* weights, float* grad_base, int32_t accumulate,
                        int32_t clear_grads, void* stream)
{
    if (!L || !ws || !grad_base || L->n_levels <= 0) { set_error("gigs_light_backward: bad arguments"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps(ST_CUBEMAP_BWD, st);
    const int n = L->n_levels;
    // padded texels of the upstream gradients: q = g (stored operators: 1 / wsum is folded into the backward weights)
    // or q = g / wsum (compute variant); the cosine filter's g is only padded
    CmPrepArgs pa{};
    pa.n = n + 1;
    int tot = 0;
    for (int i = 0; i <= n; ++i) {
        pa.begin[i] = tot;
        const int r = i < n ? L->res[i] : L->res[n - 1];
        pa.g[i] = (const float*)at(ws, i < n ? L->g_spec[i] : L->g_diffuse);
        pa.wsum[i] = (i < n && !weights) ? (const float*)at(ws, L->wsum[i]) : nullptr;
        pa.q[i] = (float4*)at(ws, i < n ? L->gq[i] : L->gq_diffuse);
        tot += 6 * r * r;
    }
    pa.begin[n + 1] = tot;
    GIGS_CUDA(launch_k(cm_prep_kernel, dim3((tot + 255) / 256), dim3(256), (size_t)(0), st, pa));
    GIGS_LAUNCH_CHECK("cm_prep_kernel");
    const int rc = weights ? cm_sparse_launch(L, ws, const_cast<void*>(weights), true, st) : cm_compute_launch(L, ws, true, st);
    if (rc) return rc;
    // down the chain, coarse to fine; the last step lands in grad_base
    if (n == 1) {
        const int m = 18 * L->res[0] * L->res[0];
        GIGS_CUDA(launch_k(cm_add_kernel, dim3((m + 255) / 256), dim3(256), (size_t)(0), st, m, (const float*)at(ws, L->g_chain[0]), (const float*)at(ws, L->g_diffuse_in),
                                                       grad_base, accumulate != 0));
        GIGS_LAUNCH_CHECK("cm_add_kernel");
    }
    for (int i = n - 1; i >= 1; --i) {
        const int Nc = L->res[i];
        const int m = 6 * 4 * Nc * Nc;
        const float* ca = (const float*)at(ws, L->g_chain[i]);
        const float* cb = i == n - 1 ? (const float*)at(ws, L->g_diffuse_in) : nullptr;
        if (i > 1)
            GIGS_CUDA(launch_k(cm_mip_backward_kernel, dim3((m + 255) / 256), dim3(256), (size_t)(0), st, Nc, ca, cb, nullptr, (float*)at(ws, L->g_chain[i - 1]), true));
        else
            GIGS_CUDA(launch_k(cm_mip_backward_kernel, dim3((m + 255) / 256), dim3(256), (size_t)(0), st, Nc, ca, cb, (const float*)at(ws, L->g_chain[0]), grad_base,
                                                                    accumulate != 0));
        GIGS_LAUNCH_CHECK("cm_mip_backward_kernel");
    }
    if (clear_grads)
        GIGS_CUDA(cudaMemsetAsync(at(ws, L->grad_begin), 0, L->grad_bytes, st));
    return 0;
}

int gigs_env_tv(int32_t base_res, const float* base, const float* dirs, int32_t env_h, int32_t env_w, float scale,
                void* scratch, uint64_t* scratch_bytes, float* grad_base, float* loss_out, int32_t accumulate_loss,
                void* stream)
{
    if (base_res <= 0 || env_h < 2 || env_w < 2 || !scratch_bytes) { set_error("gigs_env_tv: bad arguments"); return -1; }
    const int n = env_h * env_w, nblk = (n + 255) / 256;
    const uint64_t need = ((uint64_t)n * 12 + 255) / 256 * 256 + (uint64_t)nblk * 4;
    if (!scratch) { *scratch_bytes = need; return 0; }
    if (*scratch_bytes < need || !base || !dirs) { set_error("gigs_env_tv: scratch too small or NULL input"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    float* env = (float*)scratch;
    float* partials = (float*)((char*)scratch + ((uint64_t)n * 12 + 255) / 256 * 256);
    GIGS_CUDA(launch_k(env_lookup_kernel, dim3(nblk), dim3(256), (size_t)(0), st, base_res, base, dirs, n, env));
    GIGS_LAUNCH_CHECK("env_lookup_kernel");
    GIGS_CUDA(launch_k(env_tv_kernel, dim3(nblk), dim3(256), (size_t)(0), st, base_res, env_h, env_w, env, dirs, scale, partials, grad_base));
    GIGS_LAUNCH_CHECK("env_tv_kernel");
    if (loss_out) {
        GIGS_CUDA(launch_k(env_tv_sum_kernel, dim3(1), dim3(256), (size_t)(0), st, nblk, partials, scale, loss_out, accumulate_loss != 0));
        GIGS_LAUNCH_CHECK("env_tv_sum_kernel");
    }
    return 0;
}

}  // extern "C"
