// G-buffer alpha blend, backward. Replaces the backward renderCUDA
// (reference cuda_rasterizer/backward.cu:404-630), which issues 19 scalar atomicAdds per
// contributing (pixel, Gaussian) pair.
//
// Here each warp (a 16x2 pixel strip) reduces its 32 pixels' contributions with a transposing
// butterfly (16+4 values in 22 shuffles instead of 19 x 5) and then issues ONE warp-wide
// red.global.add over the Gaussian's packed 80-B accumulator row, so the L2 sees <=19 lane-atomics
// per (warp, Gaussian) on 3 sectors instead of 19 x 32 on 19 arrays. Batches are staged with TMA
// bulk copies like the forward. The tile walks only the prefix of its list that the forward
// actually reached (max n_contrib over the tile), and a Gaussian whose 1/255 iso-ellipse cannot
// reach a strip is rejected once per warp.
//
// MODE_MATERIAL is the PBR-stage fast path: when neither the colour nor the opacity map has an
// upstream gradient, dL/dalpha is identically zero in the reference as well, so only the
// w * dL/dpixel sums of albedo / roughness / metallic are needed (5 values, 9 shuffles).
#include <cstdlib>
#include "common.cuh"

namespace gigs {

constexpr int BB_THREADS = 256;
#ifndef GIGS_BB_BATCH
#define GIGS_BB_BATCH 256
#endif
constexpr int BB_BATCH = GIGS_BB_BATCH;

template <int RECF>
struct BwdSmem {
    float rec[2][BB_BATCH][RECF];
    uint32_t ids[2][BB_BATCH];
    uint64_t bar[2];
    int red[BB_THREADS / 32];
};

__device__ __forceinline__ float xor_shfl(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

// sum over the 32 lanes of v[i]; afterwards lane l (l < 16) returns the total of v[l]
__device__ __forceinline__ float warp_transpose_reduce16(float (&v)[16], int lane)
{
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = hi ? v[i] : v[i + off];
            const float keep = hi ? v[i + off] : v[i];
            v[i] = keep + xor_shfl(send, off);
        }
    }
    return v[0] + xor_shfl(v[0], 16);
}
// lane l returns the total of v[l & 3]
__device__ __forceinline__ float warp_transpose_reduce4(float (&v)[4], int lane)
{
#pragma unroll
    for (int off = 2; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = hi ? v[i] : v[i + off];
            const float keep = hi ? v[i + off] : v[i];
            v[i] = keep + xor_shfl(send, off);
        }
    }
    float r = v[0];
    r += xor_shfl(r, 4);
    r += xor_shfl(r, 8);
    r += xor_shfl(r, 16);
    return r;
}
// lane l returns the total of v[l & 7]
__device__ __forceinline__ float warp_transpose_reduce8(float (&v)[8], int lane)
{
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = hi ? v[i] : v[i + off];
            const float keep = hi ? v[i + off] : v[i];
            v[i] = keep + xor_shfl(send, off);
        }
    }
    float r = v[0];
    r += xor_shfl(r, 8);
    r += xor_shfl(r, 16);
    return r;
}

constexpr int MODE_FULL = 0, MODE_MATERIAL = 1;

template <int MODE>
__global__ void __launch_bounds__(BB_THREADS)
blend_backward_kernel(const int W, const int H, const uint2* __restrict__ ranges,
                      const uint32_t* __restrict__ point_list, const float* __restrict__ records,
                      const float* __restrict__ bg_color, const float* __restrict__ final_Ts,
                      const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix_depth,
                      const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_opacity,
                      const float* __restrict__ dL_dpix_normal, const float* __restrict__ dL_dpix_albedo,
                      const float* __restrict__ dL_dpix_roughness, const float* __restrict__ dL_dpix_metallic,
                      float* __restrict__ accum)
{
    pdl_enter();
    constexpr int RECF = (MODE == MODE_FULL) ? 12 : 8;
    constexpr uint32_t RECB = RECF * 4;
    using Smem = BwdSmem<RECF>;
    extern __shared__ __align__(128) unsigned char bb_smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(bb_smem_raw);

    const int tid = threadIdx.y * TILE_X + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t horizontal_blocks = (W + TILE_X - 1) / TILE_X;
    const uint2 pix = {blockIdx.x * TILE_X + threadIdx.x, blockIdx.y * TILE_Y + threadIdx.y};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    const int HW = H * W;

    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    // the forward never went past max(n_contrib) in this tile: walk only that prefix, back to front
    int wmax = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) S.red[warp] = wmax;
    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    int n = 0;
#pragma unroll
    for (int w = 0; w < BB_THREADS / 32; ++w) n = max(n, S.red[w]);
    n = min(n, (int)(range.y - range.x));
    const int rounds = (n + BB_BATCH - 1) / BB_BATCH;

    const float strip_x0 = (float)(blockIdx.x * TILE_X);
    const float strip_y0 = (float)(blockIdx.y * TILE_Y + (threadIdx.y & ~1));

    auto issue = [&](int b) {
        const int s = b & 1;
        const int cnt = min(BB_BATCH, n - b * BB_BATCH);
        if (tid == 0) mbar_arrive_expect_tx(&S.bar[s], (uint32_t)cnt * RECB);
        if (tid < cnt) {
            const uint32_t id = point_list[range.x + (n - 1 - (b * BB_BATCH + tid))];
            S.ids[s][tid] = id;
            bulk_g2s(&S.rec[s][tid][0], records + (size_t)id * REC_FLOATS, RECB, &S.bar[s]);
        }
    };

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    float T = T_final;
    float last_alpha = 0.f, accum_opacity = 0.f;
    float accum_rec[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
    float g_col[3] = {0.f, 0.f, 0.f}, g_nrm[3] = {0.f, 0.f, 0.f}, g_alb[3] = {0.f, 0.f, 0.f};
    float g_op = 0.f, g_rough = 0.f, g_metal = 0.f, g_depth = 0.f;
    if (inside) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (MODE == MODE_FULL) {
                g_col[i] = dL_dpix ? dL_dpix[i * HW + pix_id] : 0.f;
                g_nrm[i] = dL_dpix_normal ? dL_dpix_normal[i * HW + pix_id] : 0.f;
            }
            g_alb[i] = dL_dpix_albedo ? dL_dpix_albedo[i * HW + pix_id] : 0.f;
        }
        if (MODE == MODE_FULL) {
            g_op = dL_dpix_opacity ? dL_dpix_opacity[pix_id] : 0.f;
            g_depth = dL_dpix_depth ? dL_dpix_depth[pix_id] : 0.f;
        }
        g_rough = dL_dpix_roughness ? dL_dpix_roughness[pix_id] : 0.f;
        g_metal = dL_dpix_metallic ? dL_dpix_metallic[pix_id] : 0.f;
    }
    // the reference zeroes the normal gradient of image-border pixels (backward.cu:497-501)
    if (pix.x == 0 || pix.x == (uint32_t)(W - 1) || pix.y == 0 || pix.y == (uint32_t)(H - 1)) {
        g_nrm[0] = g_nrm[1] = g_nrm[2] = 0.f;
    }
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;
    float bg_dot_dpixel = 0.f;
    if (MODE == MODE_FULL) {
#pragma unroll
        for (int i = 0; i < 3; ++i) bg_dot_dpixel += bg_color[i] * g_col[i];
    }

    if (rounds > 0) issue(0);
    __syncthreads();  // ids of batch 0 are plain shared stores: make them visible
    for (int b = 0; b < rounds; ++b) {
        const int s = b & 1;
        if (b + 1 < rounds) issue(b + 1);
        mbar_wait(&S.bar[s], (uint32_t)((b >> 1) & 1));
        const int cnt = min(BB_BATCH, n - b * BB_BATCH);
        // forward index of entry j of this batch: n-1-(b*BATCH+j); entries at or beyond the warp's
        // deepest contributor are skipped warp-wide
        for (int jb = 0; jb < cnt; jb += 32) {
            const int fwd_hi = n - 1 - (b * BB_BATCH + jb);  // largest forward index in this chunk
            if (fwd_hi - 31 >= wmax) continue;               // whole chunk beyond this warp's reach
            bool keep = false;
            const int jl = jb + lane;
            if (jl < cnt && (fwd_hi - lane) < wmax) {
                const float4 t0 = *reinterpret_cast<const float4*>(&S.rec[s][jl][0]);
                const float4 t1 = *reinterpret_cast<const float4*>(&S.rec[s][jl][4]);
                const float cA = t0.z, cB = t0.w, cC = t1.x;
                const float hx = t0.x - strip_x0;
                const float hy = t0.y - strip_y0;
                float qmin;
                {
                    const float dy = hy;
                    const float dxs = fminf(hx, fmaxf(hx - 15.f, __fdividef(-cB * dy, cA)));
                    const float ta = 0.5f * cA * dxs * dxs, tb = cB * dxs * dy, tc = 0.5f * cC * dy * dy;
                    qmin = (ta + tb + tc) - 4e-6f * (fabsf(ta) + fabsf(tb) + fabsf(tc));
                }
                {
                    const float dy = hy - 1.f;
                    const float dxs = fminf(hx, fmaxf(hx - 15.f, __fdividef(-cB * dy, cA)));
                    const float ta = 0.5f * cA * dxs * dxs, tb = cB * dxs * dy, tc = 0.5f * cC * dy * dy;
                    qmin = fminf(qmin, (ta + tb + tc) - 4e-6f * (fabsf(ta) + fabsf(tb) + fabsf(tc)));
                }
                keep = !(cA > 0.f) || !(qmin > t1.w + 0.05f);
            }
            uint32_t mask = __ballot_sync(0xffffffffu, keep);
            while (mask) {
                const int jo = __ffs(mask) - 1;
                mask &= mask - 1;
                const int j = jb + jo;
                const int fwd_idx = fwd_hi - jo;

                bool contrib = false;
                float v16[16];
                float v4[4];
                float v8[8];
                if (MODE == MODE_FULL) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v16[i] = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i) v4[i] = 0.f;
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v8[i] = 0.f;
                }
                if (fwd_idx < last_contributor) {
                    const float4 q0 = *reinterpret_cast<const float4*>(&S.rec[s][j][0]);
                    const float4 q1 = *reinterpret_cast<const float4*>(&S.rec[s][j][4]);
                    const float2 d = {q0.x - pixf.x, q0.y - pixf.y};
                    const float power = -0.5f * (q0.z * d.x * d.x + q1.x * d.y * d.y) - q0.w * d.x * d.y;
                    if (!(power > 0.0f)) {
                        const float G = expf(power);
                        const float alpha = fminf(0.99f, q1.y * G);
                        if (!(alpha < 1.0f / 255.0f)) {
                            contrib = true;
                            T = T / (1.f - alpha);
                            const float w = alpha * T;
                            if (MODE == MODE_FULL) {
                                const float4 q2 = *reinterpret_cast<const float4*>(&S.rec[s][j][8]);
                                const float col[3] = {q2.x, q2.y, q2.z};
                                float dL_dalpha = 0.0f;
#pragma unroll
                                for (int ch = 0; ch < 3; ++ch) {
                                    const float c = col[ch];
                                    accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                                    last_color[ch] = c;
                                    dL_dalpha += (c - accum_rec[ch]) * g_col[ch];
                                    v16[A_COL + ch] = w * g_col[ch];
                                    v16[A_NRM + ch] = w * g_nrm[ch];
                                    v16[A_ALB + ch] = w * g_alb[ch];
                                }
                                v16[A_ROUGH] = w * g_rough;
                                v16[A_METAL] = w * g_metal;
                                v16[A_DEPTH] = w * g_depth;
                                accum_opacity = last_alpha + (1.f - last_alpha) * accum_opacity;
                                dL_dalpha += (1.0f - accum_opacity) * g_op;
                                dL_dalpha *= T;
                                last_alpha = alpha;
                                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

                                const float dL_dG = q1.y * dL_dalpha;
                                const float gdx = G * d.x;
                                const float gdy = G * d.y;
                                const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
                                const float dG_ddely = -gdy * q1.x - gdx * q0.w;
                                const float m2x = dL_dG * dG_ddelx * ddelx_dx;
                                const float m2y = dL_dG * dG_ddely * ddely_dy;
                                v16[A_M2X] = m2x;
                                v16[A_M2Y] = m2y;
                                v16[A_M2Z] = fabsf(m2x) + fabsf(m2y);
                                v16[A_OPAC] = G * dL_dalpha;
                                v4[0] = -0.5f * gdx * d.x * dL_dG;
                                v4[1] = -0.5f * gdx * d.y * dL_dG;
                                v4[2] = -0.5f * gdy * d.y * dL_dG;
                            } else {
                                v8[0] = w * g_rough;
                                v8[1] = w * g_alb[0];
                                v8[2] = w * g_alb[1];
                                v8[3] = w * g_alb[2];
                                v8[4] = w * g_metal;
                            }
                        }
                    }
                }
                if (__any_sync(0xffffffffu, contrib)) {
                    float* row = accum + (size_t)S.ids[s][j] * ACC_FLOATS;
                    if (MODE == MODE_FULL) {
                        const float r16 = warp_transpose_reduce16(v16, lane);
                        const float r4 = warp_transpose_reduce4(v4, lane);
                        const float r = (lane < 16) ? r16 : r4;
                        if (lane < 19) red_add_f32(row + lane, r);
                    } else {
                        const float r = warp_transpose_reduce8(v8, lane);
                        if (lane < 5) red_add_f32(row + A_ROUGH + lane, r);
                    }
                }
            }
        }
        __syncthreads();  // release stage s
    }
}

// ---------------------------------------------------------------------------------------------
// Material-only backward, the PBR-stage path (dL/dcolor = dL/dopacity = 0 => dL/dalpha = 0 in the reference too):
//     dL/d{roughness, albedo, metallic}_g = sum over pixels of w(p,g) * dL/dpixel_p,   w = alpha * T.
// Differences from the general kernel above, all aimed at the instruction count (the general kernel is issue bound,
// 85 % of issue slots busy, ~50 warp instructions per visited (warp, Gaussian) pair of which ~23 are the cross-lane
// reduction):
//   * the tile list is walked FRONT TO BACK with the forward's own recurrence (w = alpha*T; T *= 1-alpha), so the
//     weights are bit-identical to what the forward blended and the per-pair IEEE division T/(1-alpha) is gone;
//   * the cross-lane reduction is a small dense contraction through shared memory instead of a shuffle butterfly per
//     Gaussian: each surviving Gaussian's 32 per-pixel weights go into a per-warp 32 x 16 queue (one STS per lane);
//     when 16 Gaussians are queued, lane l sums w[p][l & 15] * dL/dpixel_p over 16 of the 32 pixels (half l >> 4),
//     one shuffle joins the halves, and lanes 0..15 issue the 5 reductions of "their" Gaussian. Amortised cost
//     ~9 instructions per Gaussian instead of ~23, and 5 red instructions per 16 Gaussians instead of 16.
// ---------------------------------------------------------------------------------------------
constexpr int MQ = 16;  // queue slots per warp

struct MatSmem {
    float rec[2][BB_BATCH][8];
    uint32_t ids[2][BB_BATCH];
    float4 g4[BB_THREADS];                 // per pixel: dL/d{roughness, albedo.xyz}
    float g1[BB_THREADS];                  // per pixel: dL/dmetallic
    float w[BB_THREADS / 32][32][MQ + 1];  // per warp: weights of the queued Gaussians, [pixel][slot]
    uint32_t qid[BB_THREADS / 32][MQ];     // per warp: Gaussian id of each slot
    uint64_t bar[2];
    int red[BB_THREADS / 32];
};

__global__ void __launch_bounds__(BB_THREADS)
blend_backward_material_kernel(const int W, const int H, const uint2* __restrict__ ranges,
                               const uint32_t* __restrict__ point_list, const float* __restrict__ records,
                               const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix_albedo,
                               const float* __restrict__ dL_dpix_roughness, const float* __restrict__ dL_dpix_metallic,
                               float* __restrict__ accum, const uint32_t* __restrict__ warp_masks)
{
    pdl_enter();
    constexpr uint32_t RECB = 32;
    extern __shared__ __align__(128) unsigned char bb_smem_raw[];
    MatSmem& S = *reinterpret_cast<MatSmem*>(bb_smem_raw);

    const int tid = threadIdx.y * TILE_X + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t horizontal_blocks = (W + TILE_X - 1) / TILE_X;
    int lx_, ly_;
    warp_block_pixel(tid, lx_, ly_);
    const uint2 pix = {blockIdx.x * TILE_X + lx_, blockIdx.y * TILE_Y + ly_};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    const int HW = H * W;

    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const uint32_t mask_chunk0 = (range.x >> 5) + (blockIdx.y * horizontal_blocks + blockIdx.x);
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;
    {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        float gm = 0.f;
        if (inside) {
            g.x = dL_dpix_roughness ? dL_dpix_roughness[pix_id] : 0.f;
            if (dL_dpix_albedo) {
                g.y = dL_dpix_albedo[pix_id];
                g.z = dL_dpix_albedo[HW + pix_id];
                g.w = dL_dpix_albedo[2 * HW + pix_id];
            }
            gm = dL_dpix_metallic ? dL_dpix_metallic[pix_id] : 0.f;
        }
        S.g4[tid] = g;
        S.g1[tid] = gm;
    }
    // the forward never went past max(n_contrib) in this tile: walk only that prefix
    int wmax = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) S.red[warp] = wmax;
    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    int n = 0;
#pragma unroll
    for (int w = 0; w < BB_THREADS / 32; ++w) n = max(n, S.red[w]);
    n = min(n, (int)(range.y - range.x));
    const int rounds = (n + BB_BATCH - 1) / BB_BATCH;

    int ox_, oy_;
    warp_block_origin(tid, ox_, oy_);
    const float strip_x0 = (float)(blockIdx.x * TILE_X + ox_);   // d.x over the block: [hx - (COLS-1), hx], hx = mean.x - first column
    const float strip_y0 = (float)(blockIdx.y * TILE_Y + oy_);

    auto issue = [&](int b) {
        const int s = b & 1;
        const int cnt = min(BB_BATCH, n - b * BB_BATCH);
        if (tid == 0) mbar_arrive_expect_tx(&S.bar[s], (uint32_t)cnt * RECB);
        if (tid < cnt) {
            const uint32_t id = point_list[range.x + b * BB_BATCH + tid];
            S.ids[s][tid] = id;
            bulk_g2s(&S.rec[s][tid][0], records + (size_t)id * REC_FLOATS, RECB, &S.bar[s]);
        }
    };

    // drain the warp's queue: q Gaussians (q <= MQ) x 32 pixels -> 5 sums per Gaussian
    auto flush = [&](int q) {
        __syncwarp();
        const int slot = lane & (MQ - 1), half = lane >> 4;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int p = half * 16 + k;
            const float wv = S.w[warp][p][slot];
            const float4 g = S.g4[warp * 32 + p];
            const float gm = S.g1[warp * 32 + p];
            a0 = fmaf(wv, g.x, a0);
            a1 = fmaf(wv, g.y, a1);
            a2 = fmaf(wv, g.z, a2);
            a3 = fmaf(wv, g.w, a3);
            a4 = fmaf(wv, gm, a4);
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
        a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
        a2 += __shfl_xor_sync(0xffffffffu, a2, 16);
        a3 += __shfl_xor_sync(0xffffffffu, a3, 16);
        a4 += __shfl_xor_sync(0xffffffffu, a4, 16);
        if (half == 0 && slot < q) {
            float* row = accum + (size_t)S.qid[warp][slot] * ACC_FLOATS + A_ROUGH;  // {rough, albedo.xyz, metallic}
            red_add_f32(row + 0, a0);
            red_add_f32(row + 1, a1);
            red_add_f32(row + 2, a2);
            red_add_f32(row + 3, a3);
            red_add_f32(row + 4, a4);
        }
        __syncwarp();
    };

    float T = 1.0f;
    int q = 0;  // queued Gaussians of this warp (warp-uniform)
    if (rounds > 0) issue(0);
    __syncthreads();  // ids of batch 0 and the g4/g1 staging are plain shared stores: make them visible
    for (int b = 0; b < rounds; ++b) {
        const int s = b & 1;
        if (b + 1 < rounds) issue(b + 1);
        // lane l fetches the forward's mask of chunk l of this batch (issued before the wait on the records)
        uint32_t batch_masks = 0u;
        if (warp_masks != nullptr && lane < BB_BATCH / 32 && b * BB_BATCH + lane * 32 < wmax)
            batch_masks = warp_masks[((size_t)mask_chunk0 + (b * (BB_BATCH / 32) + lane)) * WARP_MASK_WORDS + warp];
        mbar_wait(&S.bar[s], (uint32_t)((b >> 1) & 1));
        const int cnt = min(BB_BATCH, n - b * BB_BATCH);
        for (int jb = 0; jb < cnt; jb += 32) {
            const int fwd_lo = b * BB_BATCH + jb;  // forward index of the first entry of this chunk
            if (fwd_lo >= wmax) break;             // this warp's pixels were all finished before this chunk
            uint32_t mask;
            if (warp_masks != nullptr) {
                // the forward ran this very test for this warp's block on this chunk: take its result
                mask = __shfl_sync(0xffffffffu, batch_masks, jb >> 5);
                const int reach = wmax - fwd_lo;   // entries of the chunk this warp's pixels got to (> 0)
                if (reach < 32) mask &= (1u << reach) - 1u;
            } else {
                bool keep = false;
                const int jl = jb + lane;
                if (jl < cnt && (fwd_lo + lane) < wmax) {
                    const float4 t0 = *reinterpret_cast<const float4*>(&S.rec[s][jl][0]);
                    const float4 t1 = *reinterpret_cast<const float4*>(&S.rec[s][jl][4]);
                    const float cA = t0.z, cB = t0.w, cC = t1.x;
                    const float hx = t0.x - strip_x0;
                    const float hy = t0.y - strip_y0;
                    const float qmin = warp_block_qmin(cA, cB, cC, hx, hy);
                    keep = !(cA > 0.f) || !(qmin > t1.w + 0.05f);
                }
                mask = __ballot_sync(0xffffffffu, keep);
            }
            while (mask) {
                const int jo = __ffs(mask) - 1;
                mask &= mask - 1;
                const int j = jb + jo;
                float wgt = 0.f;
                if (fwd_lo + jo < last_contributor) {
                    // forward.cu:540-560 recurrence, same expressions as blend_fwd.cu
                    const float4 q0 = *reinterpret_cast<const float4*>(&S.rec[s][j][0]);
                    const float4 q1 = *reinterpret_cast<const float4*>(&S.rec[s][j][4]);
                    const float2 d = {q0.x - pixf.x, q0.y - pixf.y};
                    const float power = -0.5f * (q0.z * d.x * d.x + q1.x * d.y * d.y) - q0.w * d.x * d.y;
                    if (!(power > 0.0f)) {
                        const float alpha = fminf(0.99f, q1.y * expf(power));
                        if (!(alpha < 1.0f / 255.0f)) {
                            wgt = alpha * T;
                            T = T * (1.f - alpha);
                        }
                    }
                }
                if (__any_sync(0xffffffffu, wgt != 0.f)) {
                    S.w[warp][lane][q] = wgt;
                    if (lane == 0) S.qid[warp][q] = S.ids[s][j];
                    if (++q == MQ) {
                        flush(MQ);
                        q = 0;
                    }
                }
            }
        }
        __syncthreads();  // release stage s
    }
    if (q > 0) flush(q);
}

// ---------------------------------------------------------------------------------------------
// General backward, hybrid reduction (default for MODE_FULL work). The 19 per-Gaussian sums split into
//   * 12 SEPARABLE ones — colour3, depth, normal3, roughness, albedo3, metallic: sum_p w(p,g) * dL/dpixel_p[ch] with
//     w = alpha*T — which are a small dense contraction: the pair's weight goes into the per-warp 32 x 16 queue of
//     the material kernel above and is contracted against the per-pixel upstream gradients staged once in shared
//     memory (accumulator slots 0..11 are laid out in exactly this order);
//   * 7 NON-separable ones — mean2D.xy, |mean2D|, opacity, conic3 (slots 12..18), products of per-pair quantities —
//     which keep the transposing butterfly, now over 8 values (9 shuffles) instead of 16 + 4 (22 shuffles).
// The per-lane arithmetic (back-to-front recurrence, dL/dalpha) is the general kernel's, expression for expression.
// GROUPS = how many float4 groups of upstream maps are present: 2 = {colour, depth} + {normal, roughness} (first
// training stage: no material gradients), 3 = all.
// ---------------------------------------------------------------------------------------------
struct HybSmem {
    float rec[2][BB_BATCH][12];
    uint32_t ids[2][BB_BATCH];
    float4 g4[3][BB_THREADS];              // per pixel: {col.xyz, depth}, {nrm.xyz, rough}, {alb.xyz, metal}
    float w[BB_THREADS / 32][32][MQ + 1];
    uint32_t qid[BB_THREADS / 32][MQ];
    uint64_t bar[2];
    int red[BB_THREADS / 32];
};

template <int GROUPS>
__global__ void __launch_bounds__(BB_THREADS)
blend_backward_hybrid_kernel(const int W, const int H, const uint2* __restrict__ ranges,
                             const uint32_t* __restrict__ point_list, const float* __restrict__ records,
                             const float* __restrict__ bg_color, const float* __restrict__ final_Ts,
                             const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpix_depth,
                             const float* __restrict__ dL_dpix, const float* __restrict__ dL_dpix_opacity,
                             const float* __restrict__ dL_dpix_normal, const float* __restrict__ dL_dpix_albedo,
                             const float* __restrict__ dL_dpix_roughness, const float* __restrict__ dL_dpix_metallic,
                             float* __restrict__ accum, const uint32_t* __restrict__ warp_masks)
{
    pdl_enter();
    constexpr uint32_t RECB = 12 * 4;
    extern __shared__ __align__(128) unsigned char bb_smem_raw[];
    HybSmem& S = *reinterpret_cast<HybSmem*>(bb_smem_raw);

    const int tid = threadIdx.y * TILE_X + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const uint32_t horizontal_blocks = (W + TILE_X - 1) / TILE_X;
    int lx_, ly_;
    warp_block_pixel(tid, lx_, ly_);
    const uint2 pix = {blockIdx.x * TILE_X + lx_, blockIdx.y * TILE_Y + ly_};
    const uint32_t pix_id = W * pix.y + pix.x;
    const float2 pixf = {(float)pix.x, (float)pix.y};
    const bool inside = pix.x < (uint32_t)W && pix.y < (uint32_t)H;
    const int HW = H * W;

    const uint2 range = ranges[blockIdx.y * horizontal_blocks + blockIdx.x];
    const uint32_t mask_chunk0 = (range.x >> 5) + (blockIdx.y * horizontal_blocks + blockIdx.x);
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    float g_col[3] = {0.f, 0.f, 0.f};
    float g_op = 0.f;
    {
        float4 ga = make_float4(0.f, 0.f, 0.f, 0.f), gb = ga, gc = ga;
        if (inside) {
            if (dL_dpix) { ga.x = dL_dpix[pix_id]; ga.y = dL_dpix[HW + pix_id]; ga.z = dL_dpix[2 * HW + pix_id]; }
            ga.w = dL_dpix_depth ? dL_dpix_depth[pix_id] : 0.f;
            // the reference zeroes the normal gradient of image-border pixels (backward.cu:497-501)
            const bool border = pix.x == 0 || pix.x == (uint32_t)(W - 1) || pix.y == 0 || pix.y == (uint32_t)(H - 1);
            if (dL_dpix_normal && !border) {
                gb.x = dL_dpix_normal[pix_id]; gb.y = dL_dpix_normal[HW + pix_id]; gb.z = dL_dpix_normal[2 * HW + pix_id];
            }
            gb.w = dL_dpix_roughness ? dL_dpix_roughness[pix_id] : 0.f;
            if (GROUPS > 2) {
                if (dL_dpix_albedo) {
                    gc.x = dL_dpix_albedo[pix_id]; gc.y = dL_dpix_albedo[HW + pix_id]; gc.z = dL_dpix_albedo[2 * HW + pix_id];
                }
                gc.w = dL_dpix_metallic ? dL_dpix_metallic[pix_id] : 0.f;
            }
            g_op = dL_dpix_opacity ? dL_dpix_opacity[pix_id] : 0.f;
        }
        g_col[0] = ga.x; g_col[1] = ga.y; g_col[2] = ga.z;
        S.g4[0][tid] = ga;
        S.g4[1][tid] = gb;
        if (GROUPS > 2) S.g4[2][tid] = gc;
    }

    // the forward never went past max(n_contrib) in this tile: walk only that prefix, back to front
    int wmax = last_contributor;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane == 0) S.red[warp] = wmax;
    if (tid == 0) {
        mbar_init(&S.bar[0], 1);
        mbar_init(&S.bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    int n = 0;
#pragma unroll
    for (int w = 0; w < BB_THREADS / 32; ++w) n = max(n, S.red[w]);
    n = min(n, (int)(range.y - range.x));
    // The list is walked back to front in chunks of 32 that coincide with the FORWARD's chunks (so that its
    // footprint-test masks apply): slot k of the walk is forward index n_pad - 1 - k, n_pad = n rounded up to 32; the
    // first n_pad - n slots are padding (no record is loaded for them and no lane ever selects them: they are >= wmax).
    const int n_pad = (n + 31) & ~31;
    const int rounds = (n_pad + BB_BATCH - 1) / BB_BATCH;

    int ox_, oy_;
    warp_block_origin(tid, ox_, oy_);
    const float strip_x0 = (float)(blockIdx.x * TILE_X + ox_);   // d.x over the block: [hx - (COLS-1), hx], hx = mean.x - first column
    const float strip_y0 = (float)(blockIdx.y * TILE_Y + oy_);

    auto issue = [&](int b) {
        const int s = b & 1;
        const int cnt = min(BB_BATCH, n_pad - b * BB_BATCH);
        const int pad = (b == 0) ? (n_pad - n) : 0;          // padding slots sit at the head of batch 0
        if (tid == 0) mbar_arrive_expect_tx(&S.bar[s], (uint32_t)(cnt - pad) * RECB);
        if (tid >= pad && tid < cnt) {
            const uint32_t id = point_list[range.x + (n_pad - 1 - (b * BB_BATCH + tid))];
            S.ids[s][tid] = id;
            bulk_g2s(&S.rec[s][tid][0], records + (size_t)id * REC_FLOATS, RECB, &S.bar[s]);
        }
    };

    // drain the warp's queue: q Gaussians (q <= MQ) x 32 pixels -> 4*GROUPS sums per Gaussian (accumulator slots 0..)
    auto flush = [&](int q) {
        __syncwarp();
        const int slot = lane & (MQ - 1), half = lane >> 4;
        float4 a[GROUPS];
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) a[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int p = half * 16 + k;
            const float wv = S.w[warp][p][slot];
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                const float4 u = S.g4[g][warp * 32 + p];
                a[g].x = fmaf(wv, u.x, a[g].x);
                a[g].y = fmaf(wv, u.y, a[g].y);
                a[g].z = fmaf(wv, u.z, a[g].z);
                a[g].w = fmaf(wv, u.w, a[g].w);
            }
        }
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
            a[g].x += __shfl_xor_sync(0xffffffffu, a[g].x, 16);
            a[g].y += __shfl_xor_sync(0xffffffffu, a[g].y, 16);
            a[g].z += __shfl_xor_sync(0xffffffffu, a[g].z, 16);
            a[g].w += __shfl_xor_sync(0xffffffffu, a[g].w, 16);
        }
        if (slot < q) {
            // the two half-warps hold the same totals: half 0 issues group 0 (and 2), half 1 group 1
            float* row = accum + (size_t)S.qid[warp][slot] * ACC_FLOATS;
#pragma unroll
            for (int g = 0; g < GROUPS; ++g) {
                if ((g & 1) == half) red_add_v4(row + 4 * g, a[g].x, a[g].y, a[g].z, a[g].w);
            }
        }
        __syncwarp();
    };

    const float T_final = inside ? final_Ts[pix_id] : 0.f;
    float T = T_final;
    float last_alpha = 0.f, accum_opacity = 0.f;
    float accum_rec[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;
    float bg_dot_dpixel = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) bg_dot_dpixel += bg_color[i] * g_col[i];

    int q = 0;
    if (rounds > 0) issue(0);
    __syncthreads();  // ids of batch 0 and the g4 staging are plain shared stores: make them visible
    for (int b = 0; b < rounds; ++b) {
        const int s = b & 1;
        if (b + 1 < rounds) issue(b + 1);
        // lane l fetches the forward's mask of this batch's chunk l (forward chunk n_pad/32 - 1 - (b*BATCH/32 + l))
        uint32_t batch_masks = 0u;
        {
            const int fc = (n_pad >> 5) - 1 - (b * (BB_BATCH / 32) + lane);
            if (warp_masks != nullptr && lane < BB_BATCH / 32 && fc >= 0 && fc * 32 < wmax)
                batch_masks = warp_masks[((size_t)mask_chunk0 + fc) * WARP_MASK_WORDS + warp];
        }
        mbar_wait(&S.bar[s], (uint32_t)((b >> 1) & 1));
        const int cnt = min(BB_BATCH, n_pad - b * BB_BATCH);
        for (int jb = 0; jb < cnt; jb += 32) {
            const int fwd_hi = n_pad - 1 - (b * BB_BATCH + jb);  // largest forward index in this chunk (= 31 mod 32)
            if (fwd_hi - 31 >= wmax) continue;                   // whole chunk beyond this warp's reach
            uint32_t mask;
            if (warp_masks != nullptr) {
                // lane jo of this walk is the forward's lane 31 - jo of the same chunk
                mask = __brev(__shfl_sync(0xffffffffu, batch_masks, jb >> 5));
                const int skip = fwd_hi - wmax + 1;              // lanes below this are past the warp's reach
                if (skip > 0) mask &= ~((1u << skip) - 1u);
            } else {
                bool keep = false;
                const int jl = jb + lane;
                if (jl < cnt && (fwd_hi - lane) < wmax) {
                    const float4 t0 = *reinterpret_cast<const float4*>(&S.rec[s][jl][0]);
                    const float4 t1 = *reinterpret_cast<const float4*>(&S.rec[s][jl][4]);
                    const float cA = t0.z, cB = t0.w, cC = t1.x;
                    const float hx = t0.x - strip_x0;
                    const float hy = t0.y - strip_y0;
                    const float qmin = warp_block_qmin(cA, cB, cC, hx, hy);
                    keep = !(cA > 0.f) || !(qmin > t1.w + 0.05f);
                }
                mask = __ballot_sync(0xffffffffu, keep);
            }
            while (mask) {
                const int jo = __ffs(mask) - 1;
                mask &= mask - 1;
                const int j = jb + jo;
                const int fwd_idx = fwd_hi - jo;

                bool contrib = false;
                float wgt = 0.f;
                float v8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v8[i] = 0.f;
                if (fwd_idx < last_contributor) {
                    const float4 q0 = *reinterpret_cast<const float4*>(&S.rec[s][j][0]);
                    const float4 q1 = *reinterpret_cast<const float4*>(&S.rec[s][j][4]);
                    const float2 d = {q0.x - pixf.x, q0.y - pixf.y};
                    const float power = -0.5f * (q0.z * d.x * d.x + q1.x * d.y * d.y) - q0.w * d.x * d.y;
                    if (!(power > 0.0f)) {
                        const float G = expf(power);
                        const float alpha = fminf(0.99f, q1.y * G);
                        if (!(alpha < 1.0f / 255.0f)) {
                            contrib = true;
                            // one reciprocal serves both divisions by (1 - alpha) of backward.cu:545,584 (the general
                            // kernel above keeps the two IEEE divisions: 54 M of its 230 M warp instructions)
                            // 1 - alpha lies in [0.01, 1): the special-case guard of the IEEE division (FCHK + slow
                            // path, ~7 instructions, 7 % of this kernel) is not needed; reciprocal + one Newton step
                            // gives the same correctly rounded quotient on that range
                            const float oma = 1.f - alpha;
                            float inv1a;
                            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv1a) : "f"(oma));
                            inv1a = fmaf(inv1a, fmaf(-oma, inv1a, 1.f), inv1a);
                            T = T * inv1a;
                            wgt = alpha * T;
                            const float4 q2 = *reinterpret_cast<const float4*>(&S.rec[s][j][8]);
                            const float col[3] = {q2.x, q2.y, q2.z};
                            float dL_dalpha = 0.0f;
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) {
                                const float c = col[ch];
                                accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                                last_color[ch] = c;
                                dL_dalpha += (c - accum_rec[ch]) * g_col[ch];
                            }
                            accum_opacity = last_alpha + (1.f - last_alpha) * accum_opacity;
                            dL_dalpha += (1.0f - accum_opacity) * g_op;
                            dL_dalpha *= T;
                            last_alpha = alpha;
                            dL_dalpha += (-T_final * inv1a) * bg_dot_dpixel;

                            const float dL_dG = q1.y * dL_dalpha;
                            const float gdx = G * d.x;
                            const float gdy = G * d.y;
                            const float dG_ddelx = -gdx * q0.z - gdy * q0.w;
                            const float dG_ddely = -gdy * q1.x - gdx * q0.w;
                            const float m2x = dL_dG * dG_ddelx * ddelx_dx;
                            const float m2y = dL_dG * dG_ddely * ddely_dy;
                            v8[0] = m2x;
                            v8[1] = m2y;
                            v8[2] = fabsf(m2x) + fabsf(m2y);
                            v8[3] = G * dL_dalpha;
                            v8[4] = -0.5f * gdx * d.x * dL_dG;
                            v8[5] = -0.5f * gdx * d.y * dL_dG;
                            v8[6] = -0.5f * gdy * d.y * dL_dG;
                        }
                    }
                }
                if (__any_sync(0xffffffffu, contrib)) {
                    const uint32_t gid = S.ids[s][j];
                    const float r = warp_transpose_reduce8(v8, lane);
                    if (lane < 7) red_add_f32(accum + (size_t)gid * ACC_FLOATS + A_M2X + lane, r);
                    S.w[warp][lane][q] = wgt;
                    if (lane == 0) S.qid[warp][q] = gid;
                    if (++q == MQ) {
                        flush(MQ);
                        q = 0;
                    }
                }
            }
        }
        __syncthreads();  // release stage s
    }
    if (q > 0) flush(q);
}

int launch_blend_backward(const GigsRasterBwd* a, const Layout& L, cudaStream_t st)
{
    const GigsCamera& c = a->cam;
    const char* g = (const char*)a->geom;
    const char* im = (const char*)a->img;
    const char* bn = (const char*)a->binning;
    dim3 grid(L.tiles_x, L.tiles_y, 1), block(TILE_X, TILE_Y, 1);
    GIGS_SMEM_ATTR(blend_backward_kernel<MODE_FULL>, sizeof(BwdSmem<12>));
    GIGS_SMEM_ATTR(blend_backward_kernel<MODE_MATERIAL>, sizeof(BwdSmem<8>));
    const uint2* ranges = (const uint2*)(im + L.off.i_ranges);
    const uint32_t* plist = (const uint32_t*)(bn + L.off.b_point_list);
    const float* recs = (const float*)(g + L.off.g_record);
    const uint32_t* ncontrib = (const uint32_t*)(im + L.off.i_n_contrib);
    const float* finalT = (const float*)(im + L.off.i_final_T);
    // the forward's footprint-test masks (binning blob); GIGS_BB_NOMASKS=1: every kernel runs the test itself
    static const bool no_masks = getenv("GIGS_BB_NOMASKS") != nullptr;
    const uint32_t* masks = no_masks ? nullptr : (const uint32_t*)(bn + L.b_warp_masks);
    const bool material_only = (a->dL_dpix == nullptr) && (a->dL_dpix_opacity == nullptr) &&
                               (a->dL_dpix_normal == nullptr) && (a->dL_dpix_depth == nullptr);
    static const bool legacy_material = getenv("GIGS_BB_LEGACY") != nullptr;
    if (material_only && !legacy_material) {
        GIGS_SMEM_ATTR(blend_backward_material_kernel, sizeof(MatSmem));
        GIGS_CUDA(launch_k(blend_backward_material_kernel, dim3(grid), dim3(block), (size_t)(sizeof(MatSmem)), st, 
            c.width, c.height, ranges, plist, recs, ncontrib, a->dL_dpix_albedo, a->dL_dpix_roughness,
            a->dL_dpix_metallic, a->accum, masks));
    } else if (material_only)
        GIGS_CUDA(launch_k(blend_backward_kernel<MODE_MATERIAL>, dim3(grid), dim3(block), (size_t)(sizeof(BwdSmem<8>)), st, 
            c.width, c.height, ranges, plist, recs, c.bg, finalT, ncontrib, a->dL_dpix_depth, a->dL_dpix,
            a->dL_dpix_opacity, a->dL_dpix_normal, a->dL_dpix_albedo, a->dL_dpix_roughness, a->dL_dpix_metallic,
            a->accum));
    else if (!getenv("GIGS_BB_FULL_LEGACY")) {
        GIGS_SMEM_ATTR(blend_backward_hybrid_kernel<2>, sizeof(HybSmem));
        GIGS_SMEM_ATTR(blend_backward_hybrid_kernel<3>, sizeof(HybSmem));
        if (!a->dL_dpix_albedo && !a->dL_dpix_metallic)
            GIGS_CUDA(launch_k(blend_backward_hybrid_kernel<2>, dim3(grid), dim3(block), (size_t)(sizeof(HybSmem)), st, 
                c.width, c.height, ranges, plist, recs, c.bg, finalT, ncontrib, a->dL_dpix_depth, a->dL_dpix,
                a->dL_dpix_opacity, a->dL_dpix_normal, a->dL_dpix_albedo, a->dL_dpix_roughness, a->dL_dpix_metallic,
                a->accum, masks));
        else
            GIGS_CUDA(launch_k(blend_backward_hybrid_kernel<3>, dim3(grid), dim3(block), (size_t)(sizeof(HybSmem)), st, 
                c.width, c.height, ranges, plist, recs, c.bg, finalT, ncontrib, a->dL_dpix_depth, a->dL_dpix,
                a->dL_dpix_opacity, a->dL_dpix_normal, a->dL_dpix_albedo, a->dL_dpix_roughness, a->dL_dpix_metallic,
                a->accum, masks));
    } else
        GIGS_CUDA(launch_k(blend_backward_kernel<MODE_FULL>, dim3(grid), dim3(block), (size_t)(sizeof(BwdSmem<12>)), st, 
            c.width, c.height, ranges, plist, recs, c.bg, finalT, ncontrib, a->dL_dpix_depth, a->dL_dpix,
            a->dL_dpix_opacity, a->dL_dpix_normal, a->dL_dpix_albedo, a->dL_dpix_roughness, a->dL_dpix_metallic,
            a->accum));
    GIGS_LAUNCH_CHECK("blend_backward_kernel");
    return 0;
}

}  // namespace gigs
