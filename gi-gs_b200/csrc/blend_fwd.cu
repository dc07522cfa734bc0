// G-buffer alpha blend, forward. Replaces renderCUDA / liteRenderCUDA
// (reference cuda_rasterizer/forward.cu:423-633, :279-418).
//
// One CTA per 16x16 tile, one thread per pixel, like the reference; what differs is the data path:
//   * each tile batch (256 Gaussians) is staged into shared memory as packed 96-B records by
//     per-thread TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier, double
//     buffered so batch b+1 streams in while batch b is blended. The reference stages 28 B and
//     re-gathers 60 B of features from global memory per contributing (pixel, Gaussian) pair.
//   * warps (16x2 pixel strips) vote: a strip whose pixels are all saturated stops blending, and a
//     Gaussian whose 1/255 iso-ellipse cannot reach the strip is rejected once per warp instead of
//     once per pixel (conservative bound; the exact per-pixel test still runs for the survivors, so
//     results are unchanged).
// The per-pair arithmetic keeps the reference's expression order so n_contrib / final_T match.
#include "common.cuh"

namespace gigs {

constexpr int BL_THREADS = 256;
#ifndef GIGS_BL_BATCH
#define GIGS_BL_BATCH 256
#endif
constexpr int BL_BATCH = GIGS_BL_BATCH;
constexpr uint32_t REC_BYTES = REC_FLOATS * 4;

template <int BATCH>
struct BlendSmemT {
    float rec[2][BATCH][REC_FLOATS];  // 2 x 24 KB at 256 records
    uint64_t bar[2];
};
using BlendSmem = BlendSmemT<BL_BATCH>;
// The MATERIAL forward (PBR-stage frame) runs 128-record batches: half the shared memory lifts it from 4 to 5 resident
// CTAs per SM (then register bound), 0.179 -> 0.170 ms; the 17-channel forward and both backwards measured 1-3 % slower
// with 128 and keep 256.
constexpr int BL_BATCH_MATERIAL = 128;

// ARGMAX: track the heaviest contributor's depth / position (settings.argmax_depth); the training and evaluation
// drivers leave it off, and then the three selects + two compares per contributing pair are dead weight.
// MATERIAL (GigsRasterFwd.material_only): the radiance image and the blended position are not wanted — 5 packed FMAs
// and 3 record loads per contributing pair instead of 8 and 4; every other output is bit-identical.
// GEOM (material_only == 2, the first-stage frame): radiance, normal, depth and opacity only — the material channels,
// the blended position and the view-space normal are neither accumulated nor written (5 packed FMAs per pair).
template <bool LITE, bool ARGMAX, bool MATERIAL = false, int BATCH = BL_BATCH, bool GEOM = false>
__global__ void __launch_bounds__(BL_THREADS)
blend_forward_kernel(const int W, const int H, const uint2* __restrict__ ranges,
                     const uint32_t* __restrict__ point_list, const float* __restrict__ records,
                     const float* __restrict__ viewmatrix, const float* __restrict__ bg_color,
                     uint32_t* __restrict__ n_contrib, float* __restrict__ final_T, float* __restrict__ out_color,
                     float* __restrict__ out_opacity, float* __restrict__ out_depth, float* __restrict__ out_normal,
                     float* __restrict__ out_normal_view, float* __restrict__ out_pos,
                     float* __restrict__ out_albedo, float* __restrict__ out_roughness,
                     float* __restrict__ out_metallic, const bool inference, uint32_t* __restrict__ warp_masks)
{
    pdl_enter();
    extern __shared__ __align__(128) unsigned char bl_smem_raw[];
    using Smem = BlendSmemT<BATCH>;
    Smem& S = *reinterpret_cast<Smem*>(bl_smem_raw);

    const int tid = threadIdx.y * TILE_X + threadIdx.x;
    const int lane = tid & 31;
    const uint32_t horizontal_blocks = (W + TILE_X - 1) / TILE_X;
    int lx_, ly_;
    warp_block_pixel(tid, lx_, ly_);
    const uint2 pix = {blockIdx.x * TILE_X + lx_, blockIdx.y * TILE_Y + ly_};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    bool done = !inside;

    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const int n = (int)(range.y - range.x);
    const int rounds = (n + BATCH - 1) / BATCH;
    const uint32_t mask_chunk0 = (range.x >> 5) + (blockIdx.y * horizontal_blocks + blockIdx.x);

    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // strip bounds of this warp (2 rows x 16 columns) for the conservative rejection test
    int ox_, oy_;
    warp_block_origin(tid, ox_, oy_);
    const float strip_x0 = (float)(blockIdx.x * TILE_X + ox_);   // d.x over the block: [hx - (COLS-1), hx], hx = mean.x - first column
    const float strip_y0 = (float)(blockIdx.y * TILE_Y + oy_);

    auto issue = [&](int b) {
        const int s = b & 1;
        const int cnt = min(BATCH, n - b * BATCH);
        if (tid == 0) mbar_arrive_expect_tx(&S.bar[s], (uint32_t)cnt * REC_BYTES);
        if (tid < cnt) {
            const uint32_t id = point_list[range.x + b * BATCH + tid];
            bulk_g2s(&S.rec[s][tid][0], records + (size_t)id * REC_FLOATS, REC_BYTES, &S.bar[s]);
        }
    };

    float T = 1.0f;
    uint32_t last_contributor = 0;
    // accumulators in FFMA2 pairs: {C0,C1} {C2,roughness} {A0,A1} {A2,metallic} {N0,N1} {N2,pos.x} {pos.y,pos.z} {D,O}
    const float2 z2 = make_float2(0.f, 0.f);
    float2 C01 = z2, C2R = z2, A01 = z2, A2M = z2, N01 = z2, N2P = z2, PYZ = z2, DO = z2;
    float max_weight = 0.f, except_depth = 0.f;
    float3 except_pos = {0.f, 0.f, 0.f};

    if (rounds > 0) issue(0);
    for (int b = 0; b < rounds; ++b) {
        const int s = b & 1;
        if (b + 1 < rounds) issue(b + 1);  // stage s^1 was released by the barrier ending round b-1
        mbar_wait(&S.bar[s], (uint32_t)((b >> 1) & 1));

        const int cnt = min(BATCH, n - b * BATCH);
        bool warp_done = __all_sync(0xffffffffu, done);
        for (int jb = 0; jb < cnt && !warp_done; jb += 32) {
            // --- lane l tests Gaussian jb+l against this warp's 16x2 strip (conservative bound) ---
            bool keep = false;
            const int jl = jb + lane;
            if (jl < cnt) {
                const float4 t0 = *reinterpret_cast<const float4*>(&S.rec[s][jl][0]);
                const float4 t1 = *reinterpret_cast<const float4*>(&S.rec[s][jl][4]);
                const float cA = t0.z, cB = t0.w, cC = t1.x;
                const float hx = t0.x - strip_x0;  // d.x over the strip: [hx-15, hx]
                const float hy = t0.y - strip_y0;  // d.y over the strip: {hy, hy-1}
                const float qmin = warp_block_qmin(cA, cB, cC, hx, hy);
                // reject only when provably alpha < 1/255 on all 32 pixels; NaNs / non-convex -> keep
                keep = !(cA > 0.f) || !(qmin > t1.w + 0.05f);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, keep);
            // kept for the backward kernels, which walk the same list against the same pixel block (common.cuh: Layout)
            if (warp_masks != nullptr && lane == 0)
                warp_masks[((size_t)mask_chunk0 + ((b * BATCH + jb) >> 5)) * WARP_MASK_WORDS + (tid >> 5)] = mask;
            while (mask) {
                const int j = jb + __ffs(mask) - 1;
                mask &= mask - 1;
                if (done) continue;
                const float4 q0 = *reinterpret_cast<const float4*>(&S.rec[s][j][0]);
                const float4 q1 = *reinterpret_cast<const float4*>(&S.rec[s][j][4]);
                const float2 d = {q0.x - pixf.x, q0.y - pixf.y};
                const float power = -0.5f * (q0.z * d.x * d.x + q1.x * d.y * d.y) - q0.w * d.x * d.y;
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, q1.y * expf(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) {
                    done = true;
                    continue;
                }
                const float weight = alpha * T;
                const float depth = q1.z;
                const float2 w2 = make_float2(weight, weight);
                if (MATERIAL) {
                    const float rough = S.rec[s][j][11];
                    const float4 q3 = *reinterpret_cast<const float4*>(&S.rec[s][j][12]);
                    const float4 q4 = *reinterpret_cast<const float4*>(&S.rec[s][j][16]);
                    A01 = ffma2(make_float2(q3.x, q3.y), w2, A01);
                    A2M = ffma2(make_float2(q3.z, q3.w), w2, A2M);
                    N01 = ffma2(make_float2(q4.x, q4.y), w2, N01);
                    C2R = ffma2(make_float2(q4.z, rough), w2, C2R);   // {N2, roughness} in this variant
                    DO = ffma2(make_float2(depth, 1.0f), w2, DO);
                    T = test_T;
                    last_contributor = (uint32_t)(b * BATCH + j + 1);
                    continue;
                }
                const float4 q2 = *reinterpret_cast<const float4*>(&S.rec[s][j][8]);
                // the 17 accumulations as 8 packed FMAs (FFMA2); halves round exactly like the scalar fma
                C01 = ffma2(make_float2(q2.x, q2.y), w2, C01);
                C2R = ffma2(make_float2(q2.z, q2.w), w2, C2R);
                if (GEOM) {
                    const float4 q4 = *reinterpret_cast<const float4*>(&S.rec[s][j][16]);
                    N01 = ffma2(make_float2(q4.x, q4.y), w2, N01);
                    N2P = ffma2(make_float2(q4.z, q4.w), w2, N2P);
                } else if (!LITE) {
                    const float4 q3 = *reinterpret_cast<const float4*>(&S.rec[s][j][12]);
                    const float4 q4 = *reinterpret_cast<const float4*>(&S.rec[s][j][16]);
                    const float2 q5 = *reinterpret_cast<const float2*>(&S.rec[s][j][20]);
                    A01 = ffma2(make_float2(q3.x, q3.y), w2, A01);
                    A2M = ffma2(make_float2(q3.z, q3.w), w2, A2M);
                    N01 = ffma2(make_float2(q4.x, q4.y), w2, N01);
                    N2P = ffma2(make_float2(q4.z, q4.w), w2, N2P);
                    PYZ = ffma2(q5, w2, PYZ);
                    if (ARGMAX && weight > max_weight) except_pos = make_float3(q4.w, q5.x, q5.y);
                }
                DO = ffma2(make_float2(depth, 1.0f), w2, DO);
                if (ARGMAX && weight > max_weight) {
                    except_depth = depth;
                    max_weight = weight;
                }
                T = test_T;
                last_contributor = (uint32_t)(b * BATCH + j + 1);
            }
            warp_done = __all_sync(0xffffffffu, done);
        }
        // block vote doubles as the release barrier of stage s
        if (__syncthreads_and(done)) {
            if (b + 1 < rounds) mbar_wait(&S.bar[s ^ 1], (uint32_t)(((b + 1) >> 1) & 1));  // drain in-flight copy
            break;
        }
    }

    if (inside) {
        const float C[3] = {C01.x, C01.y, C2R.x}, N[3] = {N01.x, N01.y, MATERIAL ? C2R.x : N2P.x}, A[3] = {A01.x, A01.y, A2M.x};
        const float Rg = C2R.y, Mt = A2M.y, D = DO.x, O = DO.y;
        const float3 POS = make_float3(N2P.y, PYZ.x, PYZ.y);
        const int HW = H * W;
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        if (!MATERIAL)
            for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = C[ch] + T * bg_color[ch];
        if (GEOM) {
            for (int ch = 0; ch < 3; ch++) out_normal[ch * HW + pix_id] = N[ch];
        } else if (!LITE) {
            const float* V = viewmatrix;
            float3 Nv;
            Nv.x = V[0] * N[0] + V[4] * N[1] + V[8] * N[2];
            Nv.y = V[1] * N[0] + V[5] * N[1] + V[9] * N[2];
            Nv.z = V[2] * N[0] + V[6] * N[1] + V[10] * N[2];
            Nv = normalize3(Nv);  // NaN where N == 0, as in the reference
            out_normal_view[pix_id] = Nv.x;
            out_normal_view[HW + pix_id] = Nv.y;
            out_normal_view[2 * HW + pix_id] = Nv.z;
            for (int ch = 0; ch < 3; ch++) {
                out_normal[ch * HW + pix_id] = N[ch];
                out_albedo[ch * HW + pix_id] = A[ch];
            }
            out_roughness[pix_id] = inference ? (Rg + T) : Rg;
            out_metallic[pix_id] = Mt;
        }
        if (O > 1e-6) {
            out_depth[pix_id] = ARGMAX ? except_depth : D / O;
            if (!LITE && !MATERIAL && !GEOM) {
                out_pos[pix_id] = ARGMAX ? except_pos.x : POS.x / O;
                out_pos[HW + pix_id] = ARGMAX ? except_pos.y : POS.y / O;
                out_pos[2 * HW + pix_id] = ARGMAX ? except_pos.z : POS.z / O;
            }
        } else {
            out_depth[pix_id] = 0.0f;
            if (!LITE && !MATERIAL && !GEOM) {
                out_pos[pix_id] = 0.0f;
                out_pos[HW + pix_id] = 0.0f;
                out_pos[2 * HW + pix_id] = 0.0f;
            }
        }
        out_opacity[pix_id] = O;
    }
}

int launch_blend_forward(const GigsRasterFwd* a, const Layout& L, bool lite, cudaStream_t st)
{
    const GigsCamera& c = a->cam;
    const char* g = (const char*)a->geom;
    char* im = (char*)a->img;
    const char* bn = (const char*)a->binning;
    dim3 grid(L.tiles_x, L.tiles_y, 1), block(TILE_X, TILE_Y, 1);
    {
        const int sm = (int)sizeof(BlendSmem), smm = (int)sizeof(BlendSmemT<BL_BATCH_MATERIAL>);
        GIGS_SMEM_ATTR((blend_forward_kernel<false, false>), sm);
        GIGS_SMEM_ATTR((blend_forward_kernel<false, true>), sm);
        GIGS_SMEM_ATTR((blend_forward_kernel<true, false>), sm);
        GIGS_SMEM_ATTR((blend_forward_kernel<true, true>), sm);
        GIGS_SMEM_ATTR((blend_forward_kernel<false, false, false, BL_BATCH_MATERIAL, true>), smm);
        GIGS_SMEM_ATTR((blend_forward_kernel<false, false, true, BL_BATCH_MATERIAL>), smm);
    }
    const uint2* ranges = (const uint2*)(im + L.off.i_ranges);
    const uint32_t* plist = (const uint32_t*)(bn + L.off.b_point_list);
    const float* recs = (const float*)(g + L.off.g_record);
    uint32_t* ncontrib = (uint32_t*)(im + L.off.i_n_contrib);
    float* finalT = (float*)(im + L.off.i_final_T);
    const bool am = c.argmax_depth != 0;
    uint32_t* masks = (uint32_t*)(const_cast<char*>(bn) + L.b_warp_masks);
#define BL_LITE_ARGS c.width, c.height, ranges, plist, recs, c.viewmatrix, c.bg, ncontrib, finalT, a->out_color, a->out_opacity, \
                     a->out_depth, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, false, (uint32_t*)nullptr
#define BL_FULL_ARGS c.width, c.height, ranges, plist, recs, c.viewmatrix, c.bg, ncontrib, finalT, a->out_color, a->out_opacity, \
                     a->out_depth, a->out_normal, a->out_normal_view, a->out_pos, a->out_albedo, a->out_roughness,             \
                     a->out_metallic, c.inference != 0, masks
    if (lite && am) GIGS_CUDA(launch_k(blend_forward_kernel<true, true>, dim3(grid), dim3(block), (size_t)(sizeof(BlendSmem)), st, BL_LITE_ARGS));
    else if (lite) GIGS_CUDA(launch_k(blend_forward_kernel<true, false>, dim3(grid), dim3(block), (size_t)(sizeof(BlendSmem)), st, BL_LITE_ARGS));
    else if (am) GIGS_CUDA(launch_k(blend_forward_kernel<false, true>, dim3(grid), dim3(block), (size_t)(sizeof(BlendSmem)), st, BL_FULL_ARGS));
    else if (a->material_only == 2)
        GIGS_CUDA(launch_k(blend_forward_kernel<false, false, false, BL_BATCH_MATERIAL, true>, dim3(grid), dim3(block), (size_t)(sizeof(BlendSmemT<BL_BATCH_MATERIAL>)), st, BL_FULL_ARGS));
    else if (a->material_only)
        GIGS_CUDA(launch_k(blend_forward_kernel<false, false, true, BL_BATCH_MATERIAL>, dim3(grid), dim3(block), (size_t)(sizeof(BlendSmemT<BL_BATCH_MATERIAL>)), st, BL_FULL_ARGS));
    else GIGS_CUDA(launch_k(blend_forward_kernel<false, false>, dim3(grid), dim3(block), (size_t)(sizeof(BlendSmem)), st, BL_FULL_ARGS));
#undef BL_LITE_ARGS
#undef BL_FULL_ARGS
    GIGS_LAUNCH_CHECK("blend_forward_kernel");
    return 0;
}

}  // namespace gigs
