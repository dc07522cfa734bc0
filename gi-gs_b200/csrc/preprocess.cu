// Per-Gaussian forward stage: projection, EWA covariance, SH->RGB, tile rect, packed blend record,
// fused block-sum for the tile-count scan; then key emission and tile ranges.
//
// Follows the arithmetic of (reference, read-only):
//   cuda_rasterizer/forward.cu:22-80    SH -> RGB
//   cuda_rasterizer/forward.cu:83-122   2D covariance (EWA)
//   cuda_rasterizer/forward.cu:127-161  3D covariance from scale/quaternion
//   cuda_rasterizer/forward.cu:164-276  preprocess
//   cuda_rasterizer/rasterizer_impl.cu:70-138  duplicateWithKeys / identifyTileRanges (here: depth argsort of the
//                                              Gaussians + (tile, id) pair emission in depth order, see radix_sort.cu)
//   cuda_rasterizer/auxiliary.h:41-66,150-176  ndc2Pix, getRect, transforms, in_frustum
// The *expression order* is kept identical on purpose: tile keys (depth bits, rects) must be
// bit-exact, which requires nvcc to contract the same FMAs. The data flow is ours: one packed 96-B
// record per visible Gaussian instead of seven scattered arrays, block sums fused into this
// kernel instead of a device-wide scan pass, warp-cooperative key emission.
#include <cstdlib>
#include "common.cuh"

namespace gigs {

__device__ const float kSH_C0 = 0.28209479177387814f;
__device__ const float kSH_C1 = 0.4886025119029199f;
__device__ const float kSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float kSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 operator*(float s, const V3& v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }

// SH -> RGB for one Gaussian. sh points at this Gaussian's M x 3 coefficients.
// sh0 points at coefficient 0, shr at "coefficient 0" of the array that holds coefficients >= 1 (so coefficient i
// lives at shr + 3*i); the two coincide for the API's [P,M,3] tensor and differ for raw f_dc / f_rest leaves.
__device__ __forceinline__ V3 sh_to_rgb(int deg, const V3 pos, const V3 campos, const float* __restrict__ sh0,
                                        const float* __restrict__ shr, bool clamped[3])
{
    V3 dir = pos - campos;
    float len = sqrtf(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
    dir.x = dir.x / len;
    dir.y = dir.y / len;
    dir.z = dir.z / len;
    auto SH = [&](int i) -> V3 { return {shr[3 * i + 0], shr[3 * i + 1], shr[3 * i + 2]}; };

    V3 result = kSH_C0 * V3{sh0[0], sh0[1], sh0[2]};
    if (deg > 0) {
        float x = dir.x, y = dir.y, z = dir.z;
        result = result - kSH_C1 * y * SH(1) + kSH_C1 * z * SH(2) - kSH_C1 * x * SH(3);
        if (deg > 1) {
            float xx = x * x, yy = y * y, zz = z * z;
            float xy = x * y, yz = y * z, xz = x * z;
            result = result + kSH_C2[0] * xy * SH(4) + kSH_C2[1] * yz * SH(5) +
                     kSH_C2[2] * (2.0f * zz - xx - yy) * SH(6) + kSH_C2[3] * xz * SH(7) +
                     kSH_C2[4] * (xx - yy) * SH(8);
            if (deg > 2) {
                result = result + kSH_C3[0] * y * (3.0f * xx - yy) * SH(9) + kSH_C3[1] * xy * z * SH(10) +
                         kSH_C3[2] * y * (4.0f * zz - xx - yy) * SH(11) +
                         kSH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * SH(12) +
                         kSH_C3[4] * x * (4.0f * zz - xx - yy) * SH(13) + kSH_C3[5] * z * (xx - yy) * SH(14) +
                         kSH_C3[6] * x * (xx - 3.0f * yy) * SH(15);
            }
        }
    }
    result.x += 0.5f;
    result.y += 0.5f;
    result.z += 0.5f;
    clamped[0] = (result.x < 0);
    clamped[1] = (result.y < 0);
    clamped[2] = (result.z < 0);
    // max(result, 0) with the (a < b) ? b : a shape (NaN stays NaN)
    result.x = (result.x < 0.0f) ? 0.0f : result.x;
    result.y = (result.y < 0.0f) ? 0.0f : result.y;
    result.z = (result.z < 0.0f) ? 0.0f : result.z;
    return result;
}

__device__ __forceinline__ void cov3d_from_scale_rot(const float3 scale, float mod, const float4 rot, float* cov3D)
{
    Mat3 S = mat3_cols(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
    S.m[0][0] = mod * scale.x;
    S.m[1][1] = mod * scale.y;
    S.m[2][2] = mod * scale.z;
    // quaternion used as given (the reference does not renormalise, forward.cu:136)
    float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
    // R = (1 - 2(yy+zz), 2(xy-rz), 2(xz+ry) | 2(xy+rz), 1 - 2(xx+zz), 2(yz-rx) | 2(xz-ry), 2(yz+rx), 1 - 2(xx+yy))
    // (forward.cu:140-144). The sums of two products are where the compiler has a choice (which product is rounded,
    // which is fused) and it takes it per context: the two instantiations of this kernel came out different. Pinned
    // to what the reference's build does (read off the SASS; 6M Gaussians bit-exact against it): the shared products
    // yy, zz, xz, rz, rx are rounded, the other one of each sum goes into the fma.
    const float yy = __fmul_rn(y, y), zz = __fmul_rn(z, z), xz = __fmul_rn(x, z), rz = __fmul_rn(r, z), rx = __fmul_rn(r, x);
    const float s_yz = __fadd_rn(yy, zz), s_xz = __fmaf_rn(x, x, zz), s_xy = __fmaf_rn(x, x, yy);
    const float d_xy_rz = __fmaf_rn(x, y, -rz), a_xy_rz = __fmaf_rn(x, y, rz);
    const float a_xz_ry = __fmaf_rn(r, y, xz), d_xz_ry = __fmaf_rn(-r, y, xz);
    const float d_yz_rx = __fmaf_rn(y, z, -rx), a_yz_rx = __fmaf_rn(y, z, rx);
    Mat3 R = mat3_cols(__fsub_rn(1.f, __fadd_rn(s_yz, s_yz)), __fadd_rn(d_xy_rz, d_xy_rz), __fadd_rn(a_xz_ry, a_xz_ry),
                       __fadd_rn(a_xy_rz, a_xy_rz), __fsub_rn(1.f, __fadd_rn(s_xz, s_xz)), __fadd_rn(d_yz_rx, d_yz_rx),
                       __fadd_rn(d_xz_ry, d_xz_ry), __fadd_rn(a_yz_rx, a_yz_rx), __fsub_rn(1.f, __fadd_rn(s_xy, s_xy)));
    Mat3 M = mat3_mul(S, R);
    Mat3 Sigma = mat3_mul(mat3_transpose(M), M);
    cov3D[0] = Sigma.m[0][0];
    cov3D[1] = Sigma.m[0][1];
    cov3D[2] = Sigma.m[0][2];
    cov3D[3] = Sigma.m[1][1];
    cov3D[4] = Sigma.m[1][2];
    cov3D[5] = Sigma.m[2][2];
}

__device__ __forceinline__ float3 cov2d_ewa(const float3& mean, float focal_x, float focal_y, float tan_fovx,
                                            float tan_fovy, const float* cov3D, const float* __restrict__ V)
{
    float3 t = xform_point_4x3(mean, V);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;

    Mat3 J = mat3_cols(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z), 0.0f, focal_y / t.z,
                       -(focal_y * t.y) / (t.z * t.z), 0, 0, 0);
    Mat3 W = mat3_cols(V[0], V[4], V[8], V[1], V[5], V[9], V[2], V[6], V[10]);
    Mat3 T = mat3_mul(W, J);
    Mat3 Vrk = mat3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    Mat3 cov = mat3_mul(mat3_mul(mat3_transpose(T), mat3_transpose(Vrk)), T);
    cov.m[0][0] += 0.3f;
    cov.m[1][1] += 0.3f;
    return make_float3(cov.m[0][0], cov.m[0][1], cov.m[1][1]);
}

constexpr int PRE_THREADS = 256;

// the GaussianModel getters (scene/gaussian_model.py:178-266) for the RAW variant of the kernel
__device__ __forceinline__ float act_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// RAW: the parameter pointers are the trainer's pre-activation leaves (logits, log-scales, un-normalised
// quaternion / normal, f_dc and f_rest as two tensors) and the activations are applied on load, which removes
// ~10 elementwise launches and the 192 B/Gaussian torch.cat per frame.
template <bool RAW>
__global__ void __launch_bounds__(PRE_THREADS)
preprocess_kernel(const int P, const int D, const int M, const float* __restrict__ means3D,
                  const float* __restrict__ scales, const float scale_modifier, const float* __restrict__ rotations,
                  const float* __restrict__ opacities, const float* __restrict__ shs,
                  const float* __restrict__ sh_rest, const float* __restrict__ cov3D_precomp, const float* __restrict__ colors_precomp,
                  const float* __restrict__ normal, const float* __restrict__ albedo,
                  const float* __restrict__ roughness, const float* __restrict__ metallic,
                  const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix,
                  const float* __restrict__ cam_pos, const int W, const int H, const float tan_fovx,
                  const float tan_fovy, const float focal_x, const float focal_y, const uint32_t grid_x,
                  const uint32_t grid_y, const bool prefiltered, const bool stage_sh, const bool no_color,
                  // outputs
                  int* __restrict__ radii, float* __restrict__ records, float* __restrict__ cov3Ds,
                  uint8_t* __restrict__ clamped, uint32_t* __restrict__ tiles_touched,
                  uint32_t* __restrict__ depth_keys, uint32_t* __restrict__ block_sums,
                  uint4* __restrict__ clear_ptr, const uint32_t clear_n16)
{
    pdl_enter();
    // the depth argsort's histogram / look-back words, cleared here instead of by a memset node in front of the sort
    for (uint32_t i = blockIdx.x * PRE_THREADS + threadIdx.x; i < clear_n16; i += gridDim.x * PRE_THREADS)
        clear_ptr[i] = make_uint4(0u, 0u, 0u, 0u);
    __shared__ float sV[16], sPM[16], sCam[3];
    __shared__ uint32_t s_warp_sum[PRE_THREADS / 32];
    if (threadIdx.x < 16) {
        sV[threadIdx.x] = viewmatrix[threadIdx.x];
        sPM[threadIdx.x] = projmatrix[threadIdx.x];
    }
    if (threadIdx.x < 3) sCam[threadIdx.x] = cam_pos[threadIdx.x];

    // The SH coefficients are the bulk of this kernel's traffic (180 B of its 244 B per Gaussian at degree 3) and an
    // array of structures: one thread reading its own 45 floats touches a new 128-B line with every load (measured:
    // long-scoreboard + lg_throttle stalls, 35 % of HBM peak). The 256 Gaussians of a CTA own one CONTIGUOUS 46 KB
    // slab of it, so a single TMA bulk copy stages the slab while the threads do the projection / covariance maths;
    // each thread then reads its coefficients from shared memory (stride 45 words: conflict free).
    extern __shared__ __align__(128) float s_sh[];
    __shared__ uint64_t s_bar;
    const int sh_stride = RAW ? (M - 1) * 3 : M * 3;                       // floats per Gaussian in the staged array
    const float* sh_src = RAW ? sh_rest : shs;
    const int cta_cnt = min(PRE_THREADS, P - (int)blockIdx.x * PRE_THREADS);
    const uint32_t sh_bytes = (uint32_t)cta_cnt * sh_stride * 4u;
    // staged only when the slab is a whole number of 16-B units (always true for full CTAs at the usual degrees)
    const bool staged = (colors_precomp == nullptr) && sh_src != nullptr && sh_stride > 0 && (sh_bytes % 16u == 0u) &&
                        (((size_t)blockIdx.x * PRE_THREADS * sh_stride * 4u) % 16u == 0u) && stage_sh;
    if (threadIdx.x == 0 && staged) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&s_bar, sh_bytes);
        bulk_g2s(s_sh, sh_src + (size_t)blockIdx.x * PRE_THREADS * sh_stride, sh_bytes, &s_bar);
    }
    __syncthreads();

    const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
    uint32_t touched = 0;
    if (idx < P) {
        int my_radius_i = 0;
        uint32_t depth_key = 0xffffffffu;  // Gaussians that emit nothing sort last
        do {
            const float3 p_orig = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
            // near culling (auxiliary.h:150-176); no x/y frustum test in this fork
            const float3 p_view = xform_point_4x3(p_orig, sV);
            if (p_view.z <= 0.2f) {
                if (prefiltered) {
                    printf("Point is filtered although prefiltered is set. This shouldn't happen!");
                    __trap();
                }
                break;
            }
            const float4 p_hom = xform_point_4x4(p_orig, sPM);
            const float p_w = 1.0f / (p_hom.w + 0.0000001f);
            const float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

            const float* cov3D;
            float cov_local[6];
            if (cov3D_precomp != nullptr) {
                cov3D = cov3D_precomp + idx * 6;
            } else {
                float3 sc = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
                float4 rot = *reinterpret_cast<const float4*>(rotations + 4 * idx);
                if (RAW) {
                    sc = make_float3(expf(sc.x), expf(sc.y), expf(sc.z));
                    const float qn = fmaxf(torch_norm_inner4(rot.x, rot.y, rot.z, rot.w), 1e-12f);
                    rot = make_float4(rot.x / qn, rot.y / qn, rot.z / qn, rot.w / qn);
                }
                cov3d_from_scale_rot(sc, scale_modifier, rot, cov_local);
#pragma unroll
                for (int k = 0; k < 6; ++k) cov3Ds[idx * 6 + k] = cov_local[k];
                cov3D = cov_local;
            }
            const float3 cov = cov2d_ewa(p_orig, focal_x, focal_y, tan_fovx, tan_fovy, cov3D, sV);

            const float det = (cov.x * cov.z - cov.y * cov.y);
            if (det == 0.0f) break;
            const float det_inv = 1.f / det;
            const float3 conic = {cov.z * det_inv, -cov.y * det_inv, cov.x * det_inv};

            const float mid = 0.5f * (cov.x + cov.z);
            const float lambda1 = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
            const float lambda2 = mid - sqrtf(fmaxf(0.1f, mid * mid - det));
            const float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
            const float2 point_image = {ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H)};
            uint2 rect_min, rect_max;
            tile_rect(point_image.x, point_image.y, (int)my_radius, grid_x, grid_y, rect_min, rect_max);
            if ((rect_max.x - rect_min.x) * (rect_max.y - rect_min.y) == 0) break;

            float3 rgb;
            if (no_color) {
                rgb = make_float3(0.f, 0.f, 0.f);   // material_only: nobody reads the radiance image
            } else if (colors_precomp == nullptr) {
                bool cl[3];
                const float* sh0 = RAW ? shs + (size_t)idx * 3 : shs + (size_t)idx * M * 3;
                const float* shr = RAW ? sh_rest + ((size_t)idx * (M - 1) - 1) * 3 : sh0;
                if (staged) {
                    mbar_wait(&s_bar, 0);
                    if (RAW) {
                        shr = s_sh + (int)threadIdx.x * sh_stride - 3;
                    } else {
                        sh0 = s_sh + (int)threadIdx.x * sh_stride;
                        shr = sh0;
                    }
                }
                V3 c = sh_to_rgb(D, V3{p_orig.x, p_orig.y, p_orig.z}, V3{sCam[0], sCam[1], sCam[2]}, sh0, shr, cl);
                rgb = make_float3(c.x, c.y, c.z);
                *reinterpret_cast<uchar4*>(clamped + 4 * (size_t)idx) =
                    make_uchar4(cl[0] ? 1 : 0, cl[1] ? 1 : 0, cl[2] ? 1 : 0, 0);
            } else {
                rgb = make_float3(colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2]);
            }

            my_radius_i = (int)my_radius;
            touched = (rect_max.y - rect_min.y) * (rect_max.x - rect_min.x);
            depth_key = __float_as_uint(p_view.z);  // the low 32 bits of the reference's key (rasterizer_impl.cu:104)

            const float op = RAW ? act_sigmoid(opacities[idx]) : opacities[idx];
            float4* rec = reinterpret_cast<float4*>(records + (size_t)idx * REC_FLOATS);
            rec[0] = make_float4(point_image.x, point_image.y, conic.x, conic.y);
            rec[1] = make_float4(conic.z, op, p_view.z, logf(255.0f * op));
            if (normal != nullptr) {
                float3 nr = {normal[3 * idx], normal[3 * idx + 1], normal[3 * idx + 2]};
                float3 al = {albedo[3 * idx], albedo[3 * idx + 1], albedo[3 * idx + 2]};
                float ro = roughness[idx], me = metallic[idx];
                if (RAW) {
                    const float nn = fmaxf(torch_norm_inner3(nr.x, nr.y, nr.z), 1e-12f);
                    nr = make_float3(nr.x / nn, nr.y / nn, nr.z / nn);
                    al = make_float3(act_sigmoid(al.x), act_sigmoid(al.y), act_sigmoid(al.z));
                    ro = act_sigmoid(ro);
                    me = act_sigmoid(me);
                }
                rec[2] = make_float4(rgb.x, rgb.y, rgb.z, ro);
                rec[3] = make_float4(al.x, al.y, al.z, me);
                rec[4] = make_float4(nr.x, nr.y, nr.z, p_view.x);
            } else {  // lite path: colour/opacity/depth only
                rec[2] = make_float4(rgb.x, rgb.y, rgb.z, 0.f);
                rec[3] = make_float4(0.f, 0.f, 0.f, 0.f);
                rec[4] = make_float4(0.f, 0.f, 0.f, p_view.x);
            }
            rec[5] = make_float4(p_view.y, p_view.z, 0.f, 0.f);
        } while (false);
        radii[idx] = my_radius_i;
        tiles_touched[idx] = touched;
        depth_keys[idx] = depth_key;
    }

    // never leave the CTA with the bulk copy into its shared memory still in flight (culled threads did not wait)
    if (staged) mbar_wait(&s_bar, 0);

    // fused block sum of tiles_touched (feeds the single-block scan of block sums)
    uint32_t v = touched;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < PRE_THREADS / 32; ++w) s += s_warp_sum[w];
        block_sums[blockIdx.x] = s;
    }
}

// Exclusive scan of the per-block sums (in place) + total. One block; nb = ceil(P/256) is small
// (23k at 6M Gaussians).
// total_host (may be NULL): device alias of a page-locked host word that also receives the total, so that the host
// can read num_rendered without a copy node in the stream.
__global__ void __launch_bounds__(1024) scan_block_sums_kernel(uint32_t* __restrict__ block_sums, int nb,
                                                               uint32_t* __restrict__ total,
                                                               uint32_t* __restrict__ total_host, const bool total_only)
{
    pdl_enter();
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (total_only) {
        // only the sum is wanted (preprocess: num_rendered; the per-block offsets of the emission come from the
        // depth-ordered block sums later): a plain reduction, no scan rounds
        uint32_t v = 0;
        for (int i = threadIdx.x; i < nb; i += 1024) v += block_sums[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int w = 0; w < 32; ++w) t += s_warp[w];
            *total = t;
            if (total_host) {
                *reinterpret_cast<volatile uint32_t*>(total_host) = t;
                __threadfence_system();
            }
        }
        return;
    }
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        uint32_t v = (i < nb) ? block_sums[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
            uint32_t winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp[lane] = winc - w;  // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[warp] + (inc - v);
        if (i < nb) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total = s_carry;
        if (total_host) {
            *reinterpret_cast<volatile uint32_t*>(total_host) = s_carry;
            __threadfence_system();
        }
    }
}

// Block sums of tiles_touched taken in DEPTH order (order[] = argsort of the depth keys): feeds the same
// single-block scan as the index-order sums of preprocess.
__global__ void __launch_bounds__(PRE_THREADS)
ordered_block_sums_kernel(const int P, const uint32_t* __restrict__ order, const uint32_t* __restrict__ tiles_touched,
                          uint32_t* __restrict__ block_sums)
{
    pdl_enter();
    __shared__ uint32_t s_warp_sum[PRE_THREADS / 32];
    const int i = blockIdx.x * PRE_THREADS + threadIdx.x;
    uint32_t v = (i < P) ? tiles_touched[order[i]] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp_sum[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < PRE_THREADS / 32; ++w) t += s_warp_sum[w];
        block_sums[blockIdx.x] = t;
    }
}

// (tile id, Gaussian id) pair emission in depth order. Thread i of the grid owns the i-th Gaussian of the depth
// argsort; offsets are block_offset + in-block exclusive scan. Each warp flattens its 32 Gaussians' tile lists and
// the lanes emit consecutive entries (coalesced 4-B stores), instead of one thread looping over its whole rect.
// Because the pairs leave here ordered by (depth bits, Gaussian id), the stable sort by tile id that follows yields
// the reference's (tile, depth) order without ever materialising or moving the 64-bit keys.
__global__ void __launch_bounds__(PRE_THREADS)
emit_keys_kernel(const int P, const uint32_t* __restrict__ order, const int* __restrict__ radii,
                 const float* __restrict__ records, const uint32_t* __restrict__ tiles_touched,
                 const uint32_t* __restrict__ block_offsets, const uint32_t grid_x, const uint32_t grid_y,
                 uint32_t* __restrict__ keys, uint32_t* __restrict__ vals,
                 uint4* __restrict__ clear_ptr, const uint32_t clear_n16, const bool self_prefix)
{
    pdl_enter();
    // self_prefix: block_offsets holds the UNSCANNED block sums and this CTA adds up its predecessors' itself (a few
    // loads per thread for up to EMIT_SELF_PREFIX_BLOCKS blocks, issued first so that they overlap the gathers below)
    // instead of a one-CTA scan kernel sitting between the block sums and this kernel
    uint32_t part = 0;
    if (self_prefix)
        for (uint32_t j = threadIdx.x; j < blockIdx.x; j += PRE_THREADS) part += block_offsets[j];
    // the instance sort's histogram / look-back words (see preprocess_kernel)
    for (uint32_t i = blockIdx.x * PRE_THREADS + threadIdx.x; i < clear_n16; i += gridDim.x * PRE_THREADS)
        clear_ptr[i] = make_uint4(0u, 0u, 0u, 0u);
    __shared__ uint32_t s_warp_tot[PRE_THREADS / 32];
    __shared__ uint32_t s_part[PRE_THREADS / 32];
    __shared__ uint32_t s_pref[PRE_THREADS / 32][32];
    __shared__ uint4 s_info[PRE_THREADS / 32][32];  // rect_min.x | rect_min.y << 16, 1/width (float bits), rect width, Gaussian id

    const int i = blockIdx.x * PRE_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t cnt = 0;
    uint4 info = make_uint4(0, 0, 0, 0);
    if (i < P) {
        const uint32_t idx = order[i];
        cnt = tiles_touched[idx];
        if (cnt > 0) {
            const float2 q0 = *reinterpret_cast<const float2*>(records + (size_t)idx * REC_FLOATS);
            uint2 rmin, rmax;
            tile_rect(q0.x, q0.y, radii[idx], grid_x, grid_y, rmin, rmax);
            // {rect_min.x | rect_min.y << 16, bits of 1 / width, width, Gaussian id}
            info = make_uint4(rmin.x | (rmin.y << 16), __float_as_uint(1.0f / (float)(rmax.x - rmin.x)), rmax.x - rmin.x, idx);
        }
    }
    // warp inclusive scan
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp_tot[warp] = inc;
    s_pref[warp][lane] = inc - cnt;
    s_info[warp][lane] = info;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_part[warp] = part;
    __syncthreads();
    uint32_t warp_off = 0;
    if (self_prefix) {
#pragma unroll
        for (int w = 0; w < PRE_THREADS / 32; ++w) warp_off += s_part[w];
    } else {
        warp_off = block_offsets[blockIdx.x];
    }
    for (int w = 0; w < warp; ++w) warp_off += s_warp_tot[w];

    const uint32_t E = s_warp_tot[warp];
    for (uint32_t e = lane; e < E; e += 32) {
        // last g in [0,32) with pref[g] <= e
        int g = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1)
            if (s_pref[warp][g + s] <= e) g += s;
        const uint4 inf = s_info[warp][g];
        const uint32_t k = e - s_pref[warp][g];
        // row of entry k inside the rect = k / width. No integer divider on the SM (~20 instructions for / and %):
        // floor((k + 0.5) * (1 / width)) in float is exact for width <= 512 and k < 600 * width (checked
        // exhaustively); anything larger takes the integer division.
        uint32_t row;
        if (inf.z <= 512u && k < 600u * inf.z) row = (uint32_t)(((float)k + 0.5f) * __uint_as_float(inf.y));
        else row = k / inf.z;
        const uint32_t ty = (inf.x >> 16) + row;
        const uint32_t tx = (inf.x & 0xffffu) + (k - row * inf.z);
        keys[(size_t)warp_off + e] = ty * grid_x + tx;
        vals[(size_t)warp_off + e] = inf.w;
    }
}

// Tile ranges from the sorted keys (rasterizer_impl.cu:117-138). ranges must be zeroed first.
__global__ void __launch_bounds__(256)
tile_ranges_kernel(const uint32_t L, const uint32_t* __restrict__ tiles, uint2* __restrict__ ranges)
{
    const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= L) return;
    const uint32_t cur = tiles[idx];
    if (idx == 0)
        ranges[cur].x = 0;
    else {
        const uint32_t prev = tiles[idx - 1];
        if (cur != prev) {
            ranges[prev].y = idx;
            ranges[cur].x = idx;
        }
    }
    if (idx == L - 1) ranges[cur].y = L;
}

__global__ void __launch_bounds__(256)
mark_visible_kernel(const int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                    uint8_t* __restrict__ present)
{
    const int idx = blockIdx.x * 256 + threadIdx.x;
    if (idx >= P) return;
    const float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    const float3 pv = xform_point_4x3(p, viewmatrix);
    present[idx] = (pv.z <= 0.2f) ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
// host launchers
// The depth argsort's zeroed words live in the geometry blob; preprocess_kernel clears them when they can be written
// as whole 16-byte units (then launch_depth_argsort passes zero_bytes = 0: nothing left to clear).
static bool depth_scratch_cleared_by_preprocess(const char* g, const Layout& L)
{
    return ((uintptr_t)(g + L.p_hist) % 16 == 0) && (L.p_zero_bytes % 16 == 0) && (L.p_zero_bytes / 16 < (1ull << 32));
}

// ------------------------------------------------------------------------------------------------
// sh_rest != nullptr selects the RAW variant: a->shs is then f_dc [P,1,3], sh_rest is f_rest [P,M-1,3], and
// opacities / normal / albedo / roughness / metallic / scales / rotations are pre-activation leaves.
int launch_preprocess(const GigsRasterFwd* a, const Layout& L, cudaStream_t st, const float* sh_rest)
{
    const int P = a->P;
    const GigsCamera& c = a->cam;
    const float focal_y = c.height / (2.0f * c.tan_fovy);
    const float focal_x = c.width / (2.0f * c.tan_fovx);
    char* g = (char*)a->geom;
#define PRE_ARGS                                                                                                      \
    P, c.sh_degree, c.sh_coeffs, a->means3D, a->scales, c.scale_modifier, a->rotations, a->opacities, a->shs, sh_rest, \
        a->cov3D_precomp, a->colors_precomp, a->normal, a->albedo, a->roughness, a->metallic, c.viewmatrix,           \
        c.projmatrix, c.campos, c.width, c.height, c.tan_fovx, c.tan_fovy, focal_x, focal_y, L.tiles_x, L.tiles_y,    \
        c.prefiltered != 0, stage_sh, a->material_only == 1, a->radii, (float*)(g + L.off.g_record), (float*)(g + L.off.g_cov3D),            \
        (uint8_t*)(g + L.off.g_clamped), (uint32_t*)(g + L.off.g_tiles_touched), (uint32_t*)(g + L.off.g_depth_keys), \
        (uint32_t*)(g + L.off.g_block_sums), clear_ptr, clear_n16
    // dynamic shared memory: the CTA's slab of SH coefficients (see the kernel); not staged if it would not fit
    const size_t sh_floats = (size_t)PRE_THREADS * (sh_rest ? (c.sh_coeffs - 1) * 3 : c.sh_coeffs * 3);
    static const bool no_stage = getenv("GIGS_PRE_NOSTAGE") != nullptr;
    const bool stage_sh = !no_stage && a->material_only != 1 && a->colors_precomp == nullptr && sh_floats > 0 && sh_floats * 4 <= 96 * 1024;
    const size_t smem = stage_sh ? sh_floats * 4 : 0;
    const bool clears = depth_scratch_cleared_by_preprocess(g, L);
    uint4* clear_ptr = clears ? (uint4*)(g + L.p_hist) : nullptr;
    const uint32_t clear_n16 = clears ? (uint32_t)(L.p_zero_bytes / 16) : 0u;
    GIGS_SMEM_ATTR(preprocess_kernel<true>, 96 * 1024);
    GIGS_SMEM_ATTR(preprocess_kernel<false>, 96 * 1024);
    if (sh_rest != nullptr)
        GIGS_CUDA(launch_k(preprocess_kernel<true>, dim3(L.num_blocks), dim3(PRE_THREADS), (size_t)(smem), st, PRE_ARGS));
    else
        GIGS_CUDA(launch_k(preprocess_kernel<false>, dim3(L.num_blocks), dim3(PRE_THREADS), (size_t)(smem), st, PRE_ARGS));
#undef PRE_ARGS
    GIGS_LAUNCH_CHECK("preprocess_kernel");
    HostSlot slot;
    if (int e = host_total_slot(a, &slot)) return e;
    if (slot.dev) *reinterpret_cast<volatile uint32_t*>(slot.host) = NUM_RENDERED_PENDING;   // read_back_num_rendered polls it
    GIGS_CUDA(launch_k(scan_block_sums_kernel, dim3(1), dim3(1024), (size_t)(0), st, (uint32_t*)(g + L.off.g_block_sums), (int)L.num_blocks,
                       (uint32_t*)(g + L.off.g_num_rendered), slot.dev, true));
    GIGS_LAUNCH_CHECK("scan_block_sums_kernel");
    return 0;
}

int launch_radix_sort32(uint64_t R, int end_bit, const uint32_t* keys_u, const uint32_t* vals_u, uint32_t* keys_a,
                        uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, uint32_t* hist, uint32_t* status,
                        uint32_t* tickets, uint64_t zero_bytes, int pass_stage, cudaStream_t st);

// argsort of the Gaussians by depth bits (stable: ties keep ascending Gaussian index) -> g_order
int launch_depth_argsort(const GigsRasterFwd* a, const Layout& L, cudaStream_t st)
{
    char* g = (char*)a->geom;
    return launch_radix_sort32((uint64_t)a->P, 32, (const uint32_t*)(g + L.off.g_depth_keys), nullptr,
                               (uint32_t*)(g + L.p_keys_a), (uint32_t*)(g + L.off.g_order), (uint32_t*)(g + L.p_keys_b),
                               (uint32_t*)(g + L.p_vals_b), (uint32_t*)(g + L.p_hist), (uint32_t*)(g + L.p_status),
                               (uint32_t*)(g + L.p_ticket),
                               depth_scratch_cleared_by_preprocess(g, L) ? 0 : L.p_zero_bytes, -1, st);
}

constexpr uint32_t EMIT_SELF_PREFIX_BLOCKS = 4096;   // up to 1M Gaussians (<= 16 loads per thread)
// clear / clear_bytes: the instance sort's zeroed words; *cleared = 1 when the emit kernel cleared them
int launch_emit_keys(const GigsRasterFwd* a, const Layout& L, uint32_t* keys, uint32_t* vals, void* clear,
                     uint64_t clear_bytes, int* cleared, cudaStream_t st)
{
    if (L.tiles_x > 65535u || L.tiles_y > 65535u) { set_error("emit_keys: more than 65535 tiles along an image axis"); return -1; }
    *cleared = (clear && (uintptr_t)clear % 16 == 0 && clear_bytes % 16 == 0 && clear_bytes / 16 < (1ull << 32)) ? 1 : 0;
    char* g = (char*)a->geom;
    const uint32_t* order = (const uint32_t*)(g + L.off.g_order);
    const uint32_t* touched = (const uint32_t*)(g + L.off.g_tiles_touched);
    uint32_t* sums2 = (uint32_t*)(g + L.g_block_sums2);
    GIGS_CUDA(launch_k(ordered_block_sums_kernel, dim3(L.num_blocks), dim3(PRE_THREADS), (size_t)(0), st, a->P, order, touched, sums2));
    GIGS_LAUNCH_CHECK("ordered_block_sums_kernel");
    const bool self_prefix = L.num_blocks <= EMIT_SELF_PREFIX_BLOCKS;
    if (!self_prefix) {
        GIGS_CUDA(launch_k(scan_block_sums_kernel, dim3(1), dim3(1024), (size_t)(0), st, sums2, (int)L.num_blocks, sums2 + L.num_blocks,
                           (uint32_t*)nullptr, false));
        GIGS_LAUNCH_CHECK("scan_block_sums_kernel");
    }
    GIGS_CUDA(launch_k(emit_keys_kernel, dim3(L.num_blocks), dim3(PRE_THREADS), (size_t)(0), st, a->P, order, a->radii, (const float*)(g + L.off.g_record),
                                                          touched, sums2, L.tiles_x, L.tiles_y, keys, vals,
                                                          *cleared ? (uint4*)clear : (uint4*)nullptr,
                                                          *cleared ? (uint32_t)(clear_bytes / 16) : 0u, self_prefix));
    GIGS_LAUNCH_CHECK("emit_keys_kernel");
    return 0;
}

int launch_tile_ranges(uint64_t R, const uint32_t* tiles_sorted, uint2* ranges, uint32_t num_tiles, cudaStream_t st)
{
    GIGS_CUDA(cudaMemsetAsync(ranges, 0, (size_t)num_tiles * sizeof(uint2), st));
    if (R > 0) {
        tile_ranges_kernel<<<(unsigned)((R + 255) / 256), 256, 0, st>>>((uint32_t)R, tiles_sorted, ranges);
        GIGS_LAUNCH_CHECK("tile_ranges_kernel");
    }
    return 0;
}

int launch_mark_visible(int P, const float* means3D, const float* viewmatrix, uint8_t* present, cudaStream_t st)
{
    if (P <= 0) return 0;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, means3D, viewmatrix, present);
    GIGS_LAUNCH_CHECK("mark_visible_kernel");
    return 0;
}

}  // namespace gigs
