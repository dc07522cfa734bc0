// Fused PBR-stage frame (include/gigs_b200.h: gigs_frame_forward / gigs_frame_backward).
//
// What the reference does per view with ~200 framework launches (train.py:266-404 -> gaussian_renderer/
// __init__.py:157-199 -> pbr/shade.py:104-237 -> diff_gaussian_rasterization/__init__.py:541-743) is here
// three deferred kernels around the rasterizer and the SSR march:
//
//   deferred_shade_kernel    mask, normalise + 3x3 median of the two normal maps, rotation to view space, view
//                            directions from the canonical rays, roughness remap, split-sum shading (shade_core.cuh),
//                            background compositing, sRGB -> linear, F0 for the SSR, "lamb" prior partial sums
//   deferred_loss_kernel     linear -> sRGB of the SSR radiance, 3x3 median (+ which window element it picked),
//                            render_direct + IRR, L1 against the ground truth, dL/d render_rgb, deterministic
//                            two-level loss reduction (per-CTA partials, last CTA sums them in a fixed order)
//   deferred_backward_kernel median backward as a gather over the stored selections, sRGB and SSR (g * abd)
//                            backward, shading backward with warp-run-reduced texel gradients, lamb prior;
//                            writes dL/d{albedo, roughness, metallic} maps for the material-only blend backward
//   material_param_grad_kernel  per-Gaussian sigmoid backward, accumulated into the parameter gradients
//
// The arithmetic of each step follows the framework ops it replaces (same operation order where the op defines
// one); they are float32 elementwise chains, so agreement with the unfused path is to rounding, not bit-exact.
#include <cstdlib>
#include <cstring>
#include "filters.cuh"
#include "shade_core.cuh"
#include "gi_epilogue.cuh"

namespace gigs {

int launch_preprocess(const GigsRasterFwd* a, const Layout& L, cudaStream_t st, const float* sh_rest);
int forward_finish_impl(GigsRasterFwd* a, bool lite);
int read_back_num_rendered(GigsRasterFwd* a, const Layout& L, cudaStream_t st);
int gi_direction_count(float delta, int* n);
int launch_blend_backward(const GigsRasterBwd* a, const Layout& L, cudaStream_t st);

constexpr int TEX_PRIV_RES = 32;                                   // textures up to this face size are privatised
constexpr int TEX_PRIV_FLOATS = 6 * TEX_PRIV_RES * TEX_PRIV_RES * 3;  // floats of one copy slot
constexpr int TEX_PRIV_MAX = 4;                                     // diffuse + up to 3 specular levels

constexpr int DF_TW = 32, DF_TH = 8;              // output tile of one CTA pass (256 threads, a warp = one row)
constexpr int DF_HW1 = DF_TW + 2, DF_HH1 = DF_TH + 2;  // + 1-pixel halo

struct DeferParams {
    ShadeParams sh;          // textures + flags; sh.occlusion / sh.metallic double as "has" flags
    int W, H;
    int use_metallic;
    const float* viewmatrix; // device [16], transposed world-view (the tensor the reference indexes [:3,:3])
    const float *normal_map, *normal_view_raw, *albedo, *roughness, *metallic, *occlusion, *rays, *gt, *bg;
    float *shade_normal, *ssr_normal, *render_direct, *linear_rgb, *F0, *rough_remap, *metal_used;
    const float *ssr_color, *ssr_abd;
    float *render_rgb, *g_rgb;
    float *g_albedo, *g_roughness, *g_metallic;
    uint8_t *mask, *median_sel;
    float* partials;         // [0 .. 4*nblk): shade partials {cnt, s1, s2, SSR radiance not +0 on the tile}; [4*nblk .. 5*nblk): L1 partials;
                             // [5*nblk .. 7*nblk): BRDF TV partials {vertical, horizontal}
    uint32_t* counter;       // CTAs of the loss kernel that have finished
    float* stats;
    float* tv_edge;          // [2,H,W]: exp(-mean_c |d gt|) * mask * mask of the edge below / right of each pixel
    float loss_scale, lamb_weight, brdf_tv_weight;
    int nblk;
    int ssr_const;           // shade kernel only: SSR does not march either: its per-pixel epilogue (gi_epilogue.cuh) over a
    float ssr_dirs;          // zero gathered radiance and ssr_dirs directions is evaluated here; depth_pos = positions
    const float* depth_pos;
    int geom_skipped;        // the geometry chain did not run (GigsFrame.skip_geometry): depth_pos holds nothing; a pixel
    const float* depth;      // whose epilogue needs its position evaluates it from the depth map (filters.cuh)
    float fx, fy;
    int occl_const;          // shade kernel only: SSAO does not march (start >= step): occlusion is the constant 1, which
                             // this kernel writes to the map itself instead of reading a map a fill kernel wrote
    uint4* clear_ptr;        // backward kernel only: a 16-B aligned region it zeroes on entry (the blend backward's
    uint32_t clear_n16;      // per-Gaussian accumulator rows), in place of a memset node between the two kernels
};

__device__ __forceinline__ float srgb_to_linear_px(float s)
{
    // train.py:70-75
    const float l0 = (25.f / 323.f) * s;
    const float l1 = fast_pow((s + 0.055f) / 1.055f, 2.4f);
    return (s <= 0.04045f) ? l0 : l1;
}

// F.normalize(x, dim=0) where ||x|| > 0, x otherwise (gaussian_renderer/__init__.py:176-184)
__device__ __forceinline__ float3 normalize_where_positive(float3 v)
{
    const float n = torch_norm_outer3(v.x, v.y, v.z);
    if (n > 0.f) {
        const float d = fmaxf(n, 1e-12f);
        return make_float3(v.x / d, v.y / d, v.z / d);
    }
    return v;
}

// rotation part of inverse(world_view_transform.T) (train.py: c2w), by the adjugate; V is the transposed matrix
__device__ __forceinline__ void c2w_rotation(const float* __restrict__ V, float* __restrict__ C)
{
    // W2C[i][j] = V[4*j+i]
    const float a = V[0], b = V[4], c = V[8];
    const float d = V[1], e = V[5], f = V[9];
    const float g = V[2], h = V[6], i = V[10];
    const float A = e * i - f * h, B = -(d * i - f * g), Cc = d * h - e * g;
    const float det = a * A + b * B + c * Cc;
    const float id = 1.0f / det;
    C[0] = A * id;            C[1] = -(b * i - c * h) * id; C[2] = (b * f - c * e) * id;
    C[3] = B * id;            C[4] = (a * i - c * g) * id;  C[5] = -(a * f - c * d) * id;
    C[6] = Cc * id;           C[7] = -(a * h - b * g) * id; C[8] = (a * e - b * d) * id;
}

__device__ __forceinline__ float3 view_dir_of(const DeferParams& p, const float* __restrict__ C, size_t id)
{
    // view_dirs = -(normalize(ray)[None,:] * c2w[:3,:3]).sum(-1)   (train.py:329-337)
    float3 r = make_float3(p.rays[3 * id], p.rays[3 * id + 1], p.rays[3 * id + 2]);
    const float n = fmaxf(torch_norm_inner3(r.x, r.y, r.z), 1e-12f);
    r = make_float3(r.x / n, r.y / n, r.z / n);
    return make_float3(-(C[0] * r.x + C[1] * r.y + C[2] * r.z), -(C[3] * r.x + C[4] * r.y + C[5] * r.z),
                       -(C[6] * r.x + C[7] * r.y + C[8] * r.z));
}

__device__ __forceinline__ float block_sum_256(float v, float* s_red /*[8]*/, int tid)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((tid & 31) == 0) s_red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w];
    return t;
}

// One pixel of the 3x3 median of the normalised view-space normal map (what the shade kernel stages and filters for
// SSR's normal input), straight from the map: for the rare pixel that needs it when the lean frame skipped that filter.
// (arguments by value: a reference to the kernel's parameter struct would move the whole struct to local memory)
static __device__ __noinline__ float3 median_view_normal_pixel(const float* __restrict__ normal_view_raw, const int W,
                                                              const int H, const int x, const int y)
{
    const size_t HW = (size_t)W * H;
    float a[9], b[9], c[9];
#pragma unroll 1
    for (int k = 0; k < 9; ++k) {
        const int gx = x + (k % 3) - 1, gy = y + (k / 3) - 1;
        float3 v = make_float3(0.f, 0.f, 0.f);
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t id = (size_t)gy * W + gx;
            v = normalize_where_positive(
                make_float3(normal_view_raw[id], normal_view_raw[HW + id], normal_view_raw[2 * HW + id]));
        }
        a[k] = v.x; b[k] = v.y; c[k] = v.z;
    }
    return make_float3(median9(a), median9(b), median9(c));
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deferred_shade_kernel(const DeferParams p)
{
    pdl_enter();
    __shared__ float s_n[3][DF_HH1][DF_HW1];   // normalised normal_map (world)
    __shared__ float s_v[3][DF_HH1][DF_HW1];   // normalised out_normal_view
    __shared__ float s_C[9], s_R[9];
    __shared__ float s_red[8];
    const int tid = threadIdx.y * DF_TW + threadIdx.x;
    const int W = p.W, H = p.H;
    const size_t HW = (size_t)W * H;
    const int x0 = blockIdx.x * DF_TW, y0 = blockIdx.y * DF_TH;
    if (tid == 0) {
        c2w_rotation(p.viewmatrix, s_C);
        if (blockIdx.x == 0 && blockIdx.y == 0) *p.counter = 0u;
    }
    if (tid >= 32 && tid < 41) {
        const int k = tid - 32;
        s_R[k] = p.viewmatrix[4 * (k / 3) + (k % 3)];   // world_view_transform[:3,:3][i][j]
    }
    for (int i = tid; i < DF_HH1 * DF_HW1; i += 256) {
        const int lx = i % DF_HW1, ly = i / DF_HW1;
        const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
        float3 n = make_float3(0.f, 0.f, 0.f), v = n;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t id = (size_t)gy * W + gx;
            n = normalize_where_positive(make_float3(p.normal_map[id], p.normal_map[HW + id], p.normal_map[2 * HW + id]));
            // SSR's normal input: only a march (or an epilogue that is not provably zero) reads it; the lean frame
            // without a march (geom_skipped) leaves the filter out and evaluates single pixels on demand
            if (!p.geom_skipped)
                v = normalize_where_positive(
                    make_float3(p.normal_view_raw[id], p.normal_view_raw[HW + id], p.normal_view_raw[2 * HW + id]));
        }
        s_n[0][ly][lx] = n.x; s_n[1][ly][lx] = n.y; s_n[2][ly][lx] = n.z;
        if (!p.geom_skipped) { s_v[0][ly][lx] = v.x; s_v[1][ly][lx] = v.y; s_v[2][ly][lx] = v.z; }
    }
    __syncthreads();

    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    float cnt = 0.f, s1 = 0.f, s2 = 0.f;
    // does this tile hold an SSR radiance that is not +0.0? (unknown = yes when SSR runs as its own kernel afterwards)
    bool ssr_any = !p.ssr_const;
    if (x < W && y < H) {
        const size_t id = (size_t)y * W + x;
        float mn[3], mv[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float a[9];
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) a[dy * 3 + dx] = s_n[c][threadIdx.y + dy][threadIdx.x + dx];
            mn[c] = median9(a);
        }
        if (!p.geom_skipped) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float b[9];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) b[dy * 3 + dx] = s_v[c][threadIdx.y + dy][threadIdx.x + dx];
                mv[c] = median9(b);
                p.ssr_normal[c * HW + id] = mv[c];
            }
        }
        // normals_view = -(normal_map^T @ R)   (gaussian_renderer/__init__.py:188-190)
        ShadeIn in;
        in.n.x = -(mn[0] * s_R[0] + mn[1] * s_R[3] + mn[2] * s_R[6]);
        in.n.y = -(mn[0] * s_R[1] + mn[1] * s_R[4] + mn[2] * s_R[7]);
        in.n.z = -(mn[0] * s_R[2] + mn[1] * s_R[5] + mn[2] * s_R[8]);
        p.shade_normal[id] = in.n.x; p.shade_normal[HW + id] = in.n.y; p.shade_normal[2 * HW + id] = in.n.z;
        in.v = view_dir_of(p, s_C, id);
        in.alb = make_float3(p.albedo[id], p.albedo[HW + id], p.albedo[2 * HW + id]);
        in.rough = p.roughness[id] * (1.0f - 0.04f) + 0.04f;   // train.py:297-299
        const float metal_map = p.metallic[id];
        in.metal = p.use_metallic ? metal_map : 0.f;
        in.occ = (p.occlusion && !p.occl_const) ? p.occlusion[id] : 1.f;
        if (p.occl_const) const_cast<float*>(p.occlusion)[id] = 1.f;
        // normal_mask is taken on the RAW rasterizer normal (gaussian_renderer/__init__.py:158)
        const bool m = (p.normal_map[id] != 0.f) && (p.normal_map[HW + id] != 0.f) && (p.normal_map[2 * HW + id] != 0.f);
        PixelShade S;
        shade_eval(p.sh, in, S);
        const float lin[3] = {S.lin.x, S.lin.y, S.lin.z};
        const float alb[3] = {in.alb.x, in.alb.y, in.alb.z};
        float f0[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float xk = lin[k];
            xk = p.sh.tone ? fminf(fmaxf(aces_raw(xk), 0.f), 1.f) : fminf(fmaxf(xk, 0.f), 1.f);
            if (p.sh.gamma) xk = srgb_fwd(xk);
            const float direct = m ? xk : p.bg[k];   // train.py:367-372
            p.render_direct[k * HW + id] = direct;
            // train.py:374-378
            f0[k] = p.use_metallic ? ((1.0f - metal_map) * 0.04f + alb[k] * metal_map) : 0.04f;
            if (!p.geom_skipped) {   // inputs of the SSR march only (the lean frame without a march leaves them out)
                p.linear_rgb[k * HW + id] = srgb_to_linear_px(direct);
                p.F0[k * HW + id] = f0[k];
            }
        }
        p.rough_remap[id] = in.rough;
        p.metal_used[id] = in.metal;
        p.mask[id] = m ? 1 : 0;
        if (p.ssr_const) {
            // what ssr_nomarch_kernel (gi_march.cu) computes from the maps written above, for this pixel
            const float3 F0v = make_float3(f0[0], f0[1], f0[2]);
            float3 posv = make_float3(0.f, 0.f, 0.f);
            if (!p.geom_skipped)
                posv = make_float3(p.depth_pos[id], p.depth_pos[HW + id], p.depth_pos[2 * HW + id]);
            else if (!ssr_epilogue_is_zero(F0v, in.metal, make_float3(0.f, 0.f, 0.f), p.ssr_dirs)) {
                posv = depth_pos_pixel(x, y, W, H, p.fx, p.fy, p.depth);
                const float3 m3 = median_view_normal_pixel(p.normal_view_raw, W, H, x, y);
                mv[0] = m3.x; mv[1] = m3.y; mv[2] = m3.z;
            }
            float3 col, abd;
            ssr_epilogue_px(normalize3(make_float3(mv[0], mv[1], mv[2])), posv, in.alb, F0v, in.metal,
                            make_float3(0.f, 0.f, 0.f), p.ssr_dirs, col, abd);
            float* oc = const_cast<float*>(p.ssr_color);
            float* oa = const_cast<float*>(p.ssr_abd);
            oc[id] = col.x; oc[HW + id] = col.y; oc[2 * HW + id] = col.z;
            oa[id] = abd.x; oa[HW + id] = abd.y; oa[2 * HW + id] = abd.z;
            ssr_any = (__float_as_uint(col.x) | __float_as_uint(col.y) | __float_as_uint(col.z)) != 0u;
        }
        if (m) {
            cnt = 1.f;
            s1 = 1.0f - in.rough;
            s2 = in.metal;
        }
    }
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    const int tile_ssr_any = __syncthreads_or(ssr_any ? 1 : 0);
    const float t0 = block_sum_256(cnt, s_red, tid);
    const float t1 = block_sum_256(s1, s_red, tid);
    const float t2 = block_sum_256(s2, s_red, tid);
    if (tid == 0) {
        p.partials[4 * blk + 0] = t0;
        p.partials[4 * blk + 1] = t1;
        p.partials[4 * blk + 2] = t2;
        p.partials[4 * blk + 3] = tile_ssr_any ? 1.f : 0.f;   // read by the loss kernel of this tile and its neighbours
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) deferred_loss_kernel(const DeferParams p)
{
    pdl_enter();
    __shared__ float s_i[3][DF_HH1][DF_HW1];   // linear_to_srgb(SSR radiance), zero padded
    __shared__ float s_red[8];
    __shared__ bool s_last;
    const int tid = threadIdx.y * DF_TW + threadIdx.x;
    const int W = p.W, H = p.H;
    const size_t HW = (size_t)W * H;
    const int x0 = blockIdx.x * DF_TW, y0 = blockIdx.y * DF_TH;
    // The SSR radiance of this tile and of its eight neighbours is +0.0 everywhere (always so when nothing marches and
    // the materials are in range; the shade kernel, which then evaluates SSR's epilogue itself, left one flag per tile):
    // linear_to_srgb(+0) = +0, the 3x3 median of zeros is +0 and the first window element equal to it is element 0 -
    // the staging and the two median networks per channel are skipped with exactly that result.
    bool nb_any = false;
    if (tid < 9) {
        const int nbx = (int)blockIdx.x + tid % 3 - 1, nby = (int)blockIdx.y + tid / 3 - 1;
        if (nbx >= 0 && nbx < (int)gridDim.x && nby >= 0 && nby < (int)gridDim.y)
            nb_any = p.partials[4 * (nby * (int)gridDim.x + nbx) + 3] != 0.f;
    }
    const bool ssr_live = __syncthreads_or(nb_any ? 1 : 0) != 0;
    if (ssr_live) {
        for (int i = tid; i < DF_HH1 * DF_HW1; i += 256) {
            const int lx = i % DF_HW1, ly = i / DF_HW1;
            const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
            float v[3] = {0.f, 0.f, 0.f};
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
                const size_t id = (size_t)gy * W + gx;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = srgb_fwd(p.ssr_color[c * HW + id]);   // train.py:381 (linear_to_srgb)
            }
            s_i[0][ly][lx] = v[0]; s_i[1][ly][lx] = v[1]; s_i[2][ly][lx] = v[2];
        }
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    float l1 = 0.f;
    if (x < W && y < H) {
        const size_t id = (size_t)y * W + x;
        const float gscale = p.loss_scale / (float)(3 * HW);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float irr = 0.f;
            int sel = 0;
            if (ssr_live) {
                float a[9], b[9];
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) a[dy * 3 + dx] = b[dy * 3 + dx] = s_i[c][threadIdx.y + dy][threadIdx.x + dx];
                irr = median9(b);
                sel = median9_select(a, irr);
            }
            const float rgb = p.render_direct[c * HW + id] + irr;   // train.py:383
            p.render_rgb[c * HW + id] = rgb;
            if (p.gt) {
                const float diff = rgb - p.gt[c * HW + id];
                l1 += fabsf(diff);
                p.g_rgb[c * HW + id] = (diff > 0.f) ? gscale : ((diff < 0.f) ? -gscale : 0.f);
                p.median_sel[c * HW + id] = (sel >= 0) ? (uint8_t)sel : (uint8_t)255;
            }
        }
    }
    if (!p.gt) return;
    // BRDF smoothness prior (train.py:388-402, get_masked_tv_loss :118-142; with a full mask it equals get_tv_loss):
    // squared forward differences of [albedo, roughness (remapped), metallic] weighted by exp(-mean_c |d gt|) and by the
    // mask of both pixels of the edge. The edge weights are kept for the backward kernel.
    float tvh = 0.f, tvw = 0.f;
    if (p.brdf_tv_weight != 0.f && x < W && y < H) {
        const size_t id = (size_t)y * W + x;
        const float pr[5] = {p.albedo[id], p.albedo[HW + id], p.albedo[2 * HW + id], p.rough_remap[id], p.metal_used[id]};
        const float g0[3] = {p.gt[id], p.gt[HW + id], p.gt[2 * HW + id]};
        const float m0 = p.mask[id] ? 1.f : 0.f;
        float eh = 0.f, ew = 0.f;
        if (y + 1 < H) {
            const size_t jd = id + W;
            const float dg = (fabsf(p.gt[jd] - g0[0]) + fabsf(p.gt[HW + jd] - g0[1]) + fabsf(p.gt[2 * HW + jd] - g0[2])) / 3.0f;
            eh = expf(-dg) * (m0 * (p.mask[jd] ? 1.f : 0.f));
            const float q[5] = {p.albedo[jd], p.albedo[HW + jd], p.albedo[2 * HW + jd], p.rough_remap[jd], p.metal_used[jd]};
#pragma unroll
            for (int c = 0; c < 5; ++c) tvh += (q[c] - pr[c]) * (q[c] - pr[c]) * eh;
        }
        if (x + 1 < W) {
            const size_t jd = id + 1;
            const float dg = (fabsf(p.gt[jd] - g0[0]) + fabsf(p.gt[HW + jd] - g0[1]) + fabsf(p.gt[2 * HW + jd] - g0[2])) / 3.0f;
            ew = expf(-dg) * (m0 * (p.mask[jd] ? 1.f : 0.f));
            const float q[5] = {p.albedo[jd], p.albedo[HW + jd], p.albedo[2 * HW + jd], p.rough_remap[jd], p.metal_used[jd]};
#pragma unroll
            for (int c = 0; c < 5; ++c) tvw += (q[c] - pr[c]) * (q[c] - pr[c]) * ew;
        }
        p.tv_edge[id] = eh;
        p.tv_edge[HW + id] = ew;
    }
    const int blk = blockIdx.y * gridDim.x + blockIdx.x;
    const float t = block_sum_256(l1, s_red, tid);
    const float th = block_sum_256(tvh, s_red, tid);
    const float tw = block_sum_256(tvw, s_red, tid);
    if (tid == 0) {
        p.partials[4 * p.nblk + blk] = t;
        p.partials[5 * p.nblk + blk] = th;
        p.partials[6 * p.nblk + blk] = tw;
        __threadfence();
        const uint32_t done = atomicAdd(p.counter, 1u);
        s_last = (done == (uint32_t)p.nblk - 1u);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: fixed-order sum of the per-CTA partials (deterministic, unlike a float atomic)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f;
    for (int i = tid; i < p.nblk; i += 256) {
        a0 += __ldcg(p.partials + 4 * i + 0);
        a1 += __ldcg(p.partials + 4 * i + 1);
        a2 += __ldcg(p.partials + 4 * i + 2);
        a3 += __ldcg(p.partials + 4 * p.nblk + i);
        a4 += __ldcg(p.partials + 5 * p.nblk + i);
        a5 += __ldcg(p.partials + 6 * p.nblk + i);
    }
    const float cnt = block_sum_256(a0, s_red, tid);
    const float s1 = block_sum_256(a1, s_red, tid);
    const float s2 = block_sum_256(a2, s_red, tid);
    const float l1s = block_sum_256(a3, s_red, tid);
    const float tvhs = block_sum_256(a4, s_red, tid);
    const float tvws = block_sum_256(a5, s_red, tid);
    if (tid == 0) {
        const float l1_mean = l1s / (float)(3 * HW);
        const float c = fmaxf(cnt, 1.0f);
        // .mean() over [5,H-1,W] and [5,H,W-1] (a 1-pixel-high / -wide image has an empty mean: NaN in the reference)
        const float tv = tvhs / (5.0f * (float)(H - 1) * (float)W) + tvws / (5.0f * (float)H * (float)(W - 1));
        float loss = l1_mean + p.lamb_weight * (s1 / c + s2 / c);                    // train.py:384-386,402-404
        if (p.brdf_tv_weight != 0.f) loss += p.brdf_tv_weight * tv;                 // train.py:402
        p.stats[0] = loss * p.loss_scale;
        p.stats[1] = l1_mean;
        p.stats[2] = cnt;
        p.stats[3] = s1;
        p.stats[4] = s2;
        p.stats[5] = tv;
    }
}

// ------------------------------------------------------------------------------------------------
template <int MINB>
__global__ void __launch_bounds__(256, MINB) deferred_backward_kernel(const DeferParams p, const int tiles_x, const int ntiles)
{
    pdl_enter();
    extern __shared__ __align__(16) unsigned char dfb_raw[];
    float* s_dtex = reinterpret_cast<float*>(dfb_raw);                              // [SHB_MAX_DIFFUSE]
    float (*s_g)[DF_HH1][DF_HW1] = reinterpret_cast<float (*)[DF_HH1][DF_HW1]>(s_dtex + SHB_MAX_DIFFUSE);
    uint8_t (*s_sel)[DF_HH1][DF_HW1] = reinterpret_cast<uint8_t (*)[DF_HH1][DF_HW1]>(&s_g[3][0][0]);
    __shared__ float s_C[9];
    const int tid = threadIdx.y * DF_TW + threadIdx.x;
    const int lane = tid & 31;
    for (uint32_t i = blockIdx.x * 256u + tid; i < p.clear_n16; i += gridDim.x * 256u) p.clear_ptr[i] = make_uint4(0u, 0u, 0u, 0u);
    const int W = p.W, H = p.H;
    const size_t HW = (size_t)W * H;
    const int ndt = 6 * p.sh.diffuse_res * p.sh.diffuse_res * 3;
    const bool use_smem = (p.sh.g_diffuse_tex != nullptr) && (ndt <= SHB_MAX_DIFFUSE) && (p.sh.g_diffuse_stride == 0);
    if (use_smem)
        for (int i = tid; i < ndt; i += 256) s_dtex[i] = 0.f;
    if (tid == 0) c2w_rotation(p.viewmatrix, s_C);
    const float cnt = fmaxf(p.stats[2], 1.0f);
    const float lamb_g = p.lamb_weight * p.loss_scale / cnt;
    const bool tv_on = p.brdf_tv_weight != 0.f;
    const float tv_kh = 2.0f * p.brdf_tv_weight * p.loss_scale / (5.0f * (float)(H - 1) * (float)W);
    const float tv_kw = 2.0f * p.brdf_tv_weight * p.loss_scale / (5.0f * (float)H * (float)(W - 1));

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int x0 = (tile % tiles_x) * DF_TW, y0 = (tile / tiles_x) * DF_TH;
        __syncthreads();   // previous tile's readers are done (also orders the s_dtex / s_C initialisation)
        // The shade kernel's flag of this tile (deferred_loss_kernel): 0 = SSR's radiance, and with it the factor the
        // colour gradient is multiplied by (ssr_abd), is +0.0 on every pixel of the tile, so the gradient routed back
        // through the median and linear_to_srgb contributes g * (+0) to the albedo gradient: nothing. The gather over
        // the neighbours' selections is then skipped (the albedo gradient map may differ in the sign of a zero).
        const bool ssr_live = p.partials[4 * tile + 3] != 0.f;
        if (ssr_live) {
            for (int i = tid; i < DF_HH1 * DF_HW1; i += 256) {
                const int lx = i % DF_HW1, ly = i / DF_HW1;
                const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
                const bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
                const size_t id = (size_t)gy * W + gx;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    s_g[c][ly][lx] = in ? p.g_rgb[c * HW + id] : 0.f;
                    s_sel[c][ly][lx] = in ? p.median_sel[c * HW + id] : (uint8_t)255;
                }
            }
        }
        __syncthreads();
        const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
        const bool live = x < W && y < H;
        PixelShade S;
        ShadeGrad G;
        if (live) {
            const size_t id = (size_t)y * W + x;
            const bool m = p.mask[id] != 0;
            // median backward as a gather: neighbour (dx,dy) routed its gradient here iff it selected window
            // element (-dx,-dy), i.e. index 8 - idx(dx,dy)   (median3x3_backward_kernel, screen.cu)
            float g_alb_ssr[3], g_ren[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (!ssr_live) {
                    g_alb_ssr[c] = 0.f;
                    g_ren[c] = m ? p.g_rgb[c * HW + id] : 0.f;
                    continue;
                }
                float gi = 0.f;
#pragma unroll
                for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx)
                        if (s_sel[c][threadIdx.y + dy][threadIdx.x + dx] == (uint8_t)(8 - (dy * 3 + dx)))
                            gi += s_g[c][threadIdx.y + dy][threadIdx.x + dx];
                // linear_to_srgb backward (train.py:54-61), then _SSR.backward: grad_albedo = g * abd
                const float g_ssr = gi * srgb_bwd(p.ssr_color[c * HW + id]);
                g_alb_ssr[c] = g_ssr * p.ssr_abd[c * HW + id];
                g_ren[c] = m ? s_g[c][threadIdx.y + 1][threadIdx.x + 1] : 0.f;   // where(normal_mask, ., bg)
            }
            ShadeIn in;
            in.n = make_float3(p.shade_normal[id], p.shade_normal[HW + id], p.shade_normal[2 * HW + id]);
            in.v = view_dir_of(p, s_C, id);
            in.alb = make_float3(p.albedo[id], p.albedo[HW + id], p.albedo[2 * HW + id]);
            in.rough = p.rough_remap[id];
            in.metal = p.metal_used[id];
            in.occ = p.occlusion ? p.occlusion[id] : 1.f;
            shade_eval(p.sh, in, S);
            const float lin[3] = {S.lin.x, S.lin.y, S.lin.z};
            float gd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) gd[k] = shade_tone_bwd(p.sh, lin[k], g_ren[k]);
            shade_material_bwd(p.sh, S, gd, gd, G);
            // BRDF TV backward: d/d pred(p) of the four squared differences that touch p
            float g_tv[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (tv_on) {
                const float pr[5] = {in.alb.x, in.alb.y, in.alb.z, in.rough, in.metal};
                const float e_dn = p.tv_edge[id], e_rt = p.tv_edge[HW + id];
                const float e_up = y > 0 ? p.tv_edge[id - W] : 0.f, e_lf = x > 0 ? p.tv_edge[HW + id - 1] : 0.f;
                const size_t nb[4] = {y > 0 ? id - W : id, y + 1 < H ? id + W : id, x > 0 ? id - 1 : id, x + 1 < W ? id + 1 : id};
                const float ke[4] = {tv_kh * e_up, tv_kh * e_dn, tv_kw * e_lf, tv_kw * e_rt};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (ke[k] == 0.f) continue;
                    const size_t jd = nb[k];
                    const float q[5] = {p.albedo[jd], p.albedo[HW + jd], p.albedo[2 * HW + jd], p.rough_remap[jd], p.metal_used[jd]};
#pragma unroll
                    for (int c = 0; c < 5; ++c) g_tv[c] += (pr[c] - q[c]) * ke[k];
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) p.g_albedo[k * HW + id] = G.g_alb[k] + g_alb_ssr[k] + g_tv[k];
            // roughness_map * 0.96 + 0.04 and the lamb prior (1 - rough).mean() + metal.mean() over the mask
            p.g_roughness[id] = (G.g_rough - (m ? lamb_g : 0.f) + g_tv[3]) * (1.0f - 0.04f);
            p.g_metallic[id] = p.use_metallic ? (G.g_metal + (m ? lamb_g : 0.f) + g_tv[4]) : 0.f;
        } else {
            shade_dead_lane(S, G);
        }
        shade_texel_scatter(p.sh, S, G, s_dtex, use_smem, lane);
    }
    __syncthreads();
    if (use_smem) {
        for (int i = tid; i < ndt; i += 256) {
            const float v = s_dtex[i];
            if (v != 0.f) red_add_f32(p.sh.g_diffuse_tex + i, v);
        }
    }
}

// grad[i] += sum over the TEX_COPIES private copies (fixed order: deterministic given the copies); every privatised
// level in ONE launch: block b belongs to the segment whose [first_block, first_block + blocks) holds it
struct TexelFoldArgs {
    const float* priv[TEX_PRIV_MAX];
    float* grad[TEX_PRIV_MAX];
    int n[TEX_PRIV_MAX];
    int first_block[TEX_PRIV_MAX + 1];
    int segments;
};
__global__ void __launch_bounds__(256) texel_fold_kernel(const TexelFoldArgs a)
{
    pdl_enter();
    const float* __restrict__ priv = a.priv[0];
    float* __restrict__ grad = a.grad[0];
    int n = a.n[0], first = 0;
#pragma unroll
    for (int j = 1; j < TEX_PRIV_MAX; ++j)
        if (j < a.segments && (int)blockIdx.x >= a.first_block[j]) {
            priv = a.priv[j];
            grad = a.grad[j];
            n = a.n[j];
            first = a.first_block[j];
        }
    const int i = ((int)blockIdx.x - first) * 256 + threadIdx.x;
    if (i >= n) return;
    // all TEX_COPIES loads in flight at once (the 108-CTA grid is latency bound: four rounds of eight took 7.8 us),
    // added in the same fixed order
    float v[TEX_COPIES];
#pragma unroll
    for (int c = 0; c < TEX_COPIES; ++c) v[c] = priv[(size_t)c * TEX_PRIV_FLOATS + i];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < TEX_COPIES; ++c) s += v[c];
    if (s != 0.f) grad[i] += s;
}

// ------------------------------------------------------------------------------------------------
template <bool RAW>
__global__ void __launch_bounds__(256)
material_param_grad_kernel(const int P, const float* __restrict__ accum, const float* __restrict__ albedo,
                           const float* __restrict__ roughness, const float* __restrict__ metallic,
                           float* __restrict__ g_albedo, float* __restrict__ g_roughness, float* __restrict__ g_metallic)
{
    pdl_enter();
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= P) return;
    const float* row = accum + (size_t)idx * ACC_FLOATS;
    float g[5] = {row[A_ALB], row[A_ALB + 1], row[A_ALB + 2], row[A_ROUGH], row[A_METAL]};
    if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f && g[3] == 0.f && g[4] == 0.f) return;
    if (RAW) {
        // sigmoid backward: grad * (1 - y) * y
        const float raw[5] = {albedo[3 * idx], albedo[3 * idx + 1], albedo[3 * idx + 2], roughness[idx], metallic[idx]};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float yk = 1.0f / (1.0f + expf(-raw[k]));
            g[k] = g[k] * ((1.0f - yk) * yk);
        }
    }
    g_albedo[3 * idx] += g[0]; g_albedo[3 * idx + 1] += g[1]; g_albedo[3 * idx + 2] += g[2];
    g_roughness[idx] += g[3];
    if (g_metallic) g_metallic[idx] += g[4];
}

// ------------------------------------------------------------------------------------------------
static GigsFrameLayout frame_layout(int W, int H)
{
    GigsFrameLayout L;
    const uint64_t N = (uint64_t)W * H;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) {
        o = align_up(o, 256);
        const uint64_t r = o;
        o += bytes;
        return r;
    };
    auto f = [&](int planes) { return take((uint64_t)planes * N * 4); };
    L.color = f(3); L.opacity = f(1); L.depth = f(1); L.normal = f(3); L.normal_view = f(3); L.pos = f(3);
    L.albedo = f(3); L.roughness = f(1); L.metallic = f(1);
    L.normal_from_depth = f(3); L.depth_pos = f(3); L.occlusion = f(1);
    L.shade_normal = f(3); L.ssr_normal = f(3);
    L.render_direct = f(3); L.linear_rgb = f(3); L.F0 = f(3); L.rough_remap = f(1); L.metal_used = f(1);
    L.ssr_color = f(3); L.ssr_abd = f(3); L.render_rgb = f(3);
    L.g_rgb = f(3);
    L.g_albedo = f(3); L.g_roughness = f(1); L.g_metallic = f(1);
    L.mask = take(N);
    L.median_sel = take(3 * N);
    const uint64_t nblk = (uint64_t)((W + DF_TW - 1) / DF_TW) * ((H + DF_TH - 1) / DF_TH);
    L.tex_scratch = take((uint64_t)TEX_PRIV_MAX * TEX_PRIV_FLOATS * TEX_COPIES * 4);
    L.partials = take(7 * nblk * 4);
    L.stats = take(64);
    L.tv_edge = f(2);
    L.total_bytes = align_up(o, 256) + 256;
    return L;
}

static int frame_check(const GigsFrame* f)
{
    if (!f) { set_error("frame: null args"); return -1; }
    if (f->P <= 0) { set_error("frame: P must be positive"); return -1; }
    const GigsCamera& c = f->cam;
    if (c.width <= 1 || c.height <= 1 || !c.viewmatrix || !c.projmatrix || !c.campos || !c.bg) { set_error("frame: bad camera"); return -1; }
    if (!f->means3D || !f->sh_dc || !f->opacities || !f->normal || !f->albedo || !f->roughness || !f->metallic ||
        !f->scales || !f->rotations) { set_error("frame: a parameter pointer is NULL"); return -1; }
    if (f->raw_params && !f->sh_rest && c.sh_coeffs > 1) { set_error("frame: raw_params needs sh_rest"); return -1; }
    if (f->n_spec_levels < 2 || f->n_spec_levels > 8 || !f->diffuse || !f->brdf_lut || !f->canonical_rays) { set_error("frame: bad light / lut / rays"); return -1; }
    for (int i = 0; i < f->n_spec_levels; ++i)
        if (!f->spec[i] || f->spec_res[i] <= 0) { set_error("frame: specular level %d missing", i); return -1; }
    if (!f->geom || !f->img || !f->maps || !f->radii) { set_error("frame: workspace pointer is NULL"); return -1; }
    const GigsFrameLayout FL = frame_layout(c.width, c.height);
    if (f->maps_bytes < FL.total_bytes) { set_error("frame: maps blob too small (%llu < %llu)", (unsigned long long)f->maps_bytes, (unsigned long long)FL.total_bytes); return -2; }
    return 0;
}

static void fill_raster_fwd(const GigsFrame* f, const GigsFrameLayout& FL, GigsRasterFwd& a)
{
    memset(&a, 0, sizeof(a));
    char* m = (char*)f->maps;
    a.P = f->P;
    a.material_only = f->material_only;
    a.cam = f->cam;
    a.cam.prefiltered = 0; a.cam.argmax_depth = 0;   // cam.inference is honoured (eval / relight sweeps, forward only)
    a.means3D = f->means3D; a.shs = f->sh_dc; a.opacities = f->opacities; a.normal = f->normal; a.albedo = f->albedo;
    a.roughness = f->roughness; a.metallic = f->metallic; a.scales = f->scales; a.rotations = f->rotations;
    a.out_color = (float*)(m + FL.color); a.out_opacity = (float*)(m + FL.opacity); a.out_depth = (float*)(m + FL.depth);
    a.out_normal = (float*)(m + FL.normal); a.out_normal_view = (float*)(m + FL.normal_view);
    a.out_pos = (float*)(m + FL.pos); a.out_albedo = (float*)(m + FL.albedo);
    a.out_roughness = (float*)(m + FL.roughness); a.out_metallic = (float*)(m + FL.metallic);
    a.radii = f->radii;
    a.geom = f->geom; a.geom_bytes = f->geom_bytes; a.img = f->img; a.img_bytes = f->img_bytes;
    a.binning = f->binning; a.binning_bytes = f->binning_bytes; a.sort = f->sort; a.sort_bytes = f->sort_bytes;
    a.pinned_num_rendered = f->pinned_num_rendered;
    a.num_rendered = f->num_rendered;
    a.stream = f->stream;
}

static void fill_defer(const GigsFrame* f, const GigsFrameLayout& FL, DeferParams& p, bool backward)
{
    memset(&p, 0, sizeof(p));
    char* m = (char*)f->maps;
    const GigsCamera& c = f->cam;
    ShadeParams& s = p.sh;
    s.W = c.width; s.H = c.height; s.n_lev = f->n_spec_levels; s.diffuse_res = f->diffuse_res; s.lut_res = f->lut_res;
    s.tone = f->tone; s.gamma = f->gamma;
    for (int i = 0; i < 8; ++i) {
        s.spec_res[i] = i < f->n_spec_levels ? f->spec_res[i] : 0;
        s.spec[i] = i < f->n_spec_levels ? f->spec[i] : nullptr;
        s.g_spec[i] = (backward && i < f->n_spec_levels) ? f->g_spec[i] : nullptr;
    }
    s.diffuse = f->diffuse; s.lut = f->brdf_lut; s.rmin = f->min_roughness; s.rmax = f->max_roughness;
    s.g_diffuse_tex = backward ? f->g_diffuse_tex : nullptr;
    if (backward) {
        // small textures accumulate into private copies in the maps blob (folded into the user's gradient
        // tensors by texel_fold_kernel after the backward kernel)
        float* scratch = (float*)(m + FL.tex_scratch);
        int slot = 0;
        if (s.g_diffuse_tex && f->diffuse_res <= TEX_PRIV_RES) {
            s.g_diffuse_tex = scratch + (size_t)(slot++) * TEX_PRIV_FLOATS * TEX_COPIES;
            s.g_diffuse_stride = TEX_PRIV_FLOATS;
        }
        for (int i = f->n_spec_levels - 1; i >= 0 && slot < TEX_PRIV_MAX; --i) {
            if (s.g_spec[i] && f->spec_res[i] <= TEX_PRIV_RES) {
                s.g_spec[i] = scratch + (size_t)(slot++) * TEX_PRIV_FLOATS * TEX_COPIES;
                s.g_spec_stride[i] = TEX_PRIV_FLOATS;
            }
        }
    }
    p.W = c.width; p.H = c.height; p.use_metallic = f->use_metallic;
    p.viewmatrix = c.viewmatrix;
    p.normal_map = (float*)(m + FL.normal); p.normal_view_raw = (float*)(m + FL.normal_view);
    p.albedo = (float*)(m + FL.albedo); p.roughness = (float*)(m + FL.roughness); p.metallic = (float*)(m + FL.metallic);
    p.occlusion = f->indirect ? (float*)(m + FL.occlusion) : nullptr;
    s.occlusion = p.occlusion;                               // "has occlusion" flag for the shading core
    s.metallic = f->use_metallic ? p.metallic : nullptr;     // "has metallic" flag for the shading core
    p.rays = f->canonical_rays; p.gt = f->gt_image; p.bg = c.bg;
    p.shade_normal = (float*)(m + FL.shade_normal); p.ssr_normal = (float*)(m + FL.ssr_normal);
    p.render_direct = (float*)(m + FL.render_direct); p.linear_rgb = (float*)(m + FL.linear_rgb);
    p.F0 = (float*)(m + FL.F0); p.rough_remap = (float*)(m + FL.rough_remap); p.metal_used = (float*)(m + FL.metal_used);
    p.ssr_color = (float*)(m + FL.ssr_color); p.ssr_abd = (float*)(m + FL.ssr_abd);
    p.render_rgb = (float*)(m + FL.render_rgb); p.g_rgb = (float*)(m + FL.g_rgb);
    p.g_albedo = (float*)(m + FL.g_albedo); p.g_roughness = (float*)(m + FL.g_roughness);
    p.g_metallic = (float*)(m + FL.g_metallic);
    p.mask = (uint8_t*)(m + FL.mask); p.median_sel = (uint8_t*)(m + FL.median_sel);
    p.partials = (float*)(m + FL.partials);
    p.stats = (float*)(m + FL.stats);
    p.counter = (uint32_t*)(m + FL.stats + 32);
    p.loss_scale = f->loss_scale; p.lamb_weight = f->lamb_weight; p.brdf_tv_weight = f->brdf_tv_weight;
    p.tv_edge = (float*)(m + FL.tv_edge);
    p.nblk = ((c.width + DF_TW - 1) / DF_TW) * ((c.height + DF_TH - 1) / DF_TH);
}

}  // namespace gigs

using namespace gigs;

extern "C" {

int gigs_frame_layout(int32_t W, int32_t H, GigsFrameLayout* out)
{
    if (!out || W <= 0 || H <= 0) { set_error("gigs_frame_layout: bad arguments"); return -1; }
    *out = frame_layout(W, H);
    return 0;
}

int gigs_frame_forward(GigsFrame* f)
{
    if (int e = frame_check(f)) return e;
    const GigsCamera& c = f->cam;
    cudaStream_t st = (cudaStream_t)f->stream;
    const GigsFrameLayout FL = frame_layout(c.width, c.height);
    GigsRasterFwd a;
    fill_raster_fwd(f, FL, a);
    if (!f->resume) {
        const Layout L0 = make_layout(f->P, c.width, c.height, 0);
        if (f->geom_bytes < L0.size.geom_bytes || f->img_bytes < L0.size.img_bytes) { set_error("frame: geom/img workspace too small"); return -2; }
        {
            ProfScope ps(ST_PREPROCESS, st);
            if (int e = launch_preprocess(&a, L0, st, f->raw_params ? (f->sh_rest ? f->sh_rest : f->sh_dc) : nullptr)) return e;
        }
        if (int e = read_back_num_rendered(&a, L0, st)) return e;
        f->num_rendered = a.num_rendered;
    }
    a.num_rendered = f->num_rendered;
    const Layout L = make_layout(f->P, c.width, c.height, (uint64_t)f->num_rendered);
    f->need_binning_bytes = L.size.binning_bytes;
    f->need_sort_bytes = L.size.sort_bytes;
    if (!f->binning || !f->sort || f->binning_bytes < L.size.binning_bytes || f->sort_bytes < L.size.sort_bytes) {
        set_error("frame: binning/sort workspace too small (need %llu / %llu bytes)", (unsigned long long)L.size.binning_bytes,
                  (unsigned long long)L.size.sort_bytes);
        return GIGS_E_GROW;
    }
    if (int e = forward_finish_impl(&a, false)) return e;

    char* m = (char*)f->maps;
    const int W = c.width, H = c.height;
    const float fx = W / (2.0f * c.tan_fovx), fy = H / (2.0f * c.tan_fovy);
    const bool ssao_marches = f->start < f->step;   // otherwise the constant 1 (gi_march.cu: gi_launch), written by the shade kernel
    // normal_from_depth is a first-stage loss term and depth_pos only feeds the march (train.py:290-381): without a
    // march nothing the loss or the gradients depend on reads the chain's outputs, and skip_geometry leaves it out
    const bool geom_skipped = f->skip_geometry && !ssao_marches;
    if (!geom_skipped)
        if (int e = gigs_geometry_chain(W, H, fx, fy, c.viewmatrix, (float*)(m + FL.depth), 1, (float*)(m + FL.normal_from_depth),
                                        (float*)(m + FL.depth_pos), f->stream)) return e;
    if (f->indirect && ssao_marches) {
        if (int e = gigs_ssao(W, H, fx, fy, f->radius, f->bias, f->thick, f->delta, f->step, f->start,
                              (float*)(m + FL.normal_view), (float*)(m + FL.depth_pos), (float*)(m + FL.occlusion),
                              m + FL.tex_scratch, gigs_gi_scratch_bytes(W, H), f->stream))
            return e;   // the march's scratch: the head of the texel-gradient scratch, which only the backward uses
    }
    DeferParams p;
    fill_defer(f, FL, p, false);
    p.occl_const = (f->indirect && !ssao_marches) ? 1 : 0;
    int ssr_dirs = 0;
    const bool ssr_fused = !ssao_marches && gi_direction_count(f->delta, &ssr_dirs) == 0;
    p.ssr_const = ssr_fused ? 1 : 0;
    p.ssr_dirs = (float)ssr_dirs;
    p.depth_pos = (const float*)(m + FL.depth_pos);
    p.geom_skipped = geom_skipped ? 1 : 0;
    p.depth = (const float*)(m + FL.depth);
    p.fx = fx; p.fy = fy;
    dim3 grid((W + DF_TW - 1) / DF_TW, (H + DF_TH - 1) / DF_TH), block(DF_TW, DF_TH);
    {
        ProfScope ps(ST_DEFER_SHADE, st);
        GIGS_CUDA(launch_k(deferred_shade_kernel, dim3(grid), dim3(block), (size_t)(0), st, p));
        GIGS_LAUNCH_CHECK("deferred_shade_kernel");
    }
    if (!ssr_fused)
    if (int e = gigs_ssr(W, H, fx, fy, f->radius, f->bias, f->thick, f->delta, f->step, f->start, p.ssr_normal,
                         (float*)(m + FL.depth_pos), p.linear_rgb, p.albedo, p.rough_remap, p.metal_used, p.F0,
                         (float*)(m + FL.ssr_color), (float*)(m + FL.ssr_abd), m + FL.tex_scratch,
                         gigs_gi_scratch_bytes(W, H), f->stream))
        return e;
    if (f->gt_ready_event) GIGS_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)f->gt_ready_event, 0));
    {
        ProfScope ps(ST_DEFER_LOSS, st);
        GIGS_CUDA(launch_k(deferred_loss_kernel, dim3(grid), dim3(block), (size_t)(0), st, p));
        GIGS_LAUNCH_CHECK("deferred_loss_kernel");
    }
    if (c.debug) GIGS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int gigs_frame_backward(GigsFrame* f)
{
    if (int e = frame_check(f)) return e;
    if (!f->gt_image) { set_error("frame_backward: the forward ran without a ground-truth image"); return -1; }
    if (!f->accum || !f->g_albedo || !f->g_roughness || (f->use_metallic && !f->g_metallic)) { set_error("frame_backward: gradient / accum pointer is NULL"); return -1; }
    if (!f->binning) { set_error("frame_backward: binning workspace is NULL"); return -1; }
    const GigsCamera& c = f->cam;
    cudaStream_t st = (cudaStream_t)f->stream;
    const GigsFrameLayout FL = frame_layout(c.width, c.height);
    const int W = c.width, H = c.height;
    DeferParams p;
    fill_defer(f, FL, p, true);
    const int tiles_x = (W + DF_TW - 1) / DF_TW, ntiles = p.nblk;
    const size_t smem = SHB_MAX_DIFFUSE * sizeof(float) + 3 * DF_HH1 * DF_HW1 * (sizeof(float) + 1) + 16;
    bool accum_cleared = false;
    {
        ProfScope ps(ST_DEFER_BWD, st);
        char* m = (char*)f->maps;
        float* scratch = (float*)(m + FL.tex_scratch);
        // the same slot assignment as fill_defer: diffuse first, then the specular levels from coarse to fine
        float* fold_dst[TEX_PRIV_MAX];
        int fold_n[TEX_PRIV_MAX], slots = 0;
        if (f->g_diffuse_tex && f->diffuse_res <= TEX_PRIV_RES) {
            fold_dst[slots] = f->g_diffuse_tex;
            fold_n[slots++] = 6 * f->diffuse_res * f->diffuse_res * 3;
        }
        for (int i = f->n_spec_levels - 1; i >= 0 && slots < TEX_PRIV_MAX; --i) {
            if (f->g_spec[i] && f->spec_res[i] <= TEX_PRIV_RES) {
                fold_dst[slots] = f->g_spec[i];
                fold_n[slots++] = 6 * f->spec_res[i] * f->spec_res[i] * 3;
            }
        }
        if (slots > 0)
            GIGS_CUDA(cudaMemsetAsync(scratch, 0, (size_t)slots * TEX_PRIV_FLOATS * TEX_COPIES * sizeof(float), st));
        // the blend backward's accumulator rows are cleared by this kernel's first instructions (no memset node)
        const size_t accum_bytes = (size_t)f->P * ACC_FLOATS * sizeof(float);
        accum_cleared = ((uintptr_t)f->accum % 16 == 0) && (accum_bytes % 16 == 0) && (accum_bytes / 16 < (1ull << 32));
        p.clear_ptr = accum_cleared ? (uint4*)f->accum : nullptr;
        p.clear_n16 = accum_cleared ? (uint32_t)(accum_bytes / 16) : 0u;
        static const int variant = getenv("GIGS_DFB") ? atoi(getenv("GIGS_DFB")) : 3;  // 3 CTAs/SM measured best (203 vs 214 us)
        const int per_sm = variant == 3 ? 3 : (variant == 4 ? 4 : 2);
        const int blocks = ntiles < 148 * per_sm * 2 ? ntiles : 148 * per_sm * 2;
        {
            ProfScope pk(ST_DEFER_BWD_KERNEL, st);   // the kernel alone (the stage around it adds the clear and the folds)
            if (variant == 3) GIGS_CUDA(launch_k(deferred_backward_kernel<3>, dim3(blocks), dim3(dim3(DF_TW, DF_TH)), (size_t)(smem), st, p, tiles_x, ntiles));
            else if (variant == 4) GIGS_CUDA(launch_k(deferred_backward_kernel<4>, dim3(blocks), dim3(dim3(DF_TW, DF_TH)), (size_t)(smem), st, p, tiles_x, ntiles));
            else GIGS_CUDA(launch_k(deferred_backward_kernel<2>, dim3(blocks), dim3(dim3(DF_TW, DF_TH)), (size_t)(smem), st, p, tiles_x, ntiles));
        }
        GIGS_LAUNCH_CHECK("deferred_backward_kernel");
        if (slots > 0) {
            TexelFoldArgs fa;
            fa.segments = slots;
            int nb = 0;
            for (int k = 0; k < slots; ++k) {
                fa.priv[k] = scratch + (size_t)k * TEX_PRIV_FLOATS * TEX_COPIES;
                fa.grad[k] = fold_dst[k];
                fa.n[k] = fold_n[k];
                fa.first_block[k] = nb;
                nb += (fold_n[k] + 255) / 256;
            }
            fa.first_block[slots] = nb;
            GIGS_CUDA(launch_k(texel_fold_kernel, dim3(nb), dim3(256), (size_t)0, st, fa));
            GIGS_LAUNCH_CHECK("texel_fold_kernel");
        }
    }
    if (f->light_ready_event) GIGS_CUDA(cudaEventRecord((cudaEvent_t)f->light_ready_event, st));
    GigsRasterBwd b;
    memset(&b, 0, sizeof(b));
    b.P = f->P; b.num_rendered = f->num_rendered; b.cam = c;
    b.geom = f->geom; b.binning = f->binning; b.img = f->img;
    b.dL_dpix_albedo = p.g_albedo; b.dL_dpix_roughness = p.g_roughness;
    b.dL_dpix_metallic = f->use_metallic ? p.g_metallic : nullptr;
    b.accum = f->accum;
    const Layout L = make_layout(f->P, W, H, (uint64_t)f->num_rendered);
    {
        ProfScope ps(ST_BLEND_BWD, st);
        if (!accum_cleared) GIGS_CUDA(cudaMemsetAsync(f->accum, 0, (size_t)f->P * ACC_FLOATS * sizeof(float), st));
        if (f->num_rendered > 0)
            if (int e = launch_blend_backward(&b, L, st)) return e;
    }
    {
        ProfScope ps(ST_PARAM_GRAD, st);
        const int blocks = (f->P + 255) / 256;
        if (f->raw_params)
            GIGS_CUDA(launch_k(material_param_grad_kernel<true>, dim3(blocks), dim3(256), (size_t)(0), st, f->P, f->accum, f->albedo, f->roughness, f->metallic,
                                                                     f->g_albedo, f->g_roughness,
                                                                     f->use_metallic ? f->g_metallic : nullptr));
        else
            GIGS_CUDA(launch_k(material_param_grad_kernel<false>, dim3(blocks), dim3(256), (size_t)(0), st, f->P, f->accum, f->albedo, f->roughness, f->metallic,
                                                                      f->g_albedo, f->g_roughness,
                                                                      f->use_metallic ? f->g_metallic : nullptr));
        GIGS_LAUNCH_CHECK("material_param_grad_kernel");
    }
    if (c.debug) GIGS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
