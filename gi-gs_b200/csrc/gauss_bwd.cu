// Per-Gaussian backward: one fused kernel replacing computeCov2DCUDA + backward preprocessCUDA
// (reference cuda_rasterizer/backward.cu:145-279, :351-401, SH/cov3D helpers :21-140, :283-346) and
// the 14 torch::zeros fills that precede them (rasterize_points.cu:299-312).
//
// Reads the packed 80-B accumulator row the blend backward reduced into, chains through conic ->
// cov2D -> cov3D -> (scale, quaternion), mean2D -> mean3D, colour -> SH (+ view-direction term),
// and writes EVERY output element exactly once (zeros for Gaussians with radius <= 0), so the
// caller can hand in uninitialised tensors.
#include <cstdlib>
#include "common.cuh"

namespace gigs {

__device__ const float bSH_C0 = 0.28209479177387814f;
__device__ const float bSH_C1 = 0.4886025119029199f;
__device__ const float bSH_C2[] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                   -1.0925484305920792f, 0.5462742152960396f};
__device__ const float bSH_C3[] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                   -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

struct B3 {
    float x, y, z;
};
__device__ __forceinline__ B3 operator*(float s, const B3& v) { return {s * v.x, s * v.y, s * v.z}; }
__device__ __forceinline__ B3 operator*(const B3& v, float s) { return {v.x * s, v.y * s, v.z * s}; }
__device__ __forceinline__ B3 operator+(const B3& a, const B3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ B3 operator-(const B3& a) { return {-a.x, -a.y, -a.z}; }
__device__ __forceinline__ void operator+=(B3& a, const B3& b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ float dotB(const B3& a, const B3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

constexpr int GB_THREADS = 128;

// RAW variant (first-stage frame, csrc/stage1.cu): the parameter pointers are the trainer's pre-activation leaves
// (means3D = xyz, shs = f_dc [P,1,3] + raw.f_rest [P,M-1,3], scales = log-scales, rotations = un-normalised quaternions)
// — the same convention as preprocess_kernel<RAW> — and every result is chained through the getter
// (scene/gaussian_model.py:178-266: exp, sigmoid, F.normalize) and ACCUMULATED (+=) into the leaves' gradient tensors,
// which removes the activated copies, the 14 intermediate gradient tensors and the ~20 autograd launches of the getters.

__device__ __forceinline__ float gb_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }
// backward of F.normalize(v, eps=1e-12): y = v / max(|v|, eps)
__device__ __forceinline__ void gb_normalize_bwd(const float* v, const float* g, int n, float* out)
{
    float nn = 0.f;
    for (int i = 0; i < n; ++i) nn += v[i] * v[i];
    nn = sqrtf(nn);
    if (nn > 1e-12f) {
        float d = 0.f;
        for (int i = 0; i < n; ++i) d += (v[i] / nn) * g[i];
        for (int i = 0; i < n; ++i) out[i] = (g[i] - (v[i] / nn) * d) / nn;
    } else {
        for (int i = 0; i < n; ++i) out[i] = g[i] / 1e-12f;
    }
}

#ifdef GIGS_GB_MINB
#define GB_BOUNDS __launch_bounds__(GB_THREADS, GIGS_GB_MINB)
#else
#define GB_BOUNDS __launch_bounds__(GB_THREADS)
#endif
// s_row: RAW fast path only — this Gaussian's 45 f_rest floats staged in shared memory by the kernel below; the SH
// gradient is written back into the same row (in place: every read of the row precedes the first write)
template <bool RAW>
__device__ __forceinline__ void
gaussian_backward_body(const int idx, float* __restrict__ s_row,
                       const int P, const int D, const int M, const float* __restrict__ means3D,
                         const int* __restrict__ radii, const float* __restrict__ shs,
                         const uint8_t* __restrict__ clamped, const float* __restrict__ scales,
                         const float* __restrict__ rotations, const float scale_modifier,
                         const float* __restrict__ cov3Ds, const float* __restrict__ view_matrix,
                         const float* __restrict__ proj, const float* __restrict__ campos, const float h_x,
                         const float h_y, const float tan_fovx, const float tan_fovy,
                         const float* __restrict__ accum,
                         // outputs
                         float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic_out,
                         float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor,
                         float* __restrict__ dL_dnormal, float* __restrict__ dL_dalbedo,
                         float* __restrict__ dL_droughness, float* __restrict__ dL_dmetallic,
                         float* __restrict__ dL_dmean3D, float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh,
                         float* __restrict__ dL_dscale, float* __restrict__ dL_drot, const RawGrads& raw)
{
    const bool visible = radii[idx] > 0;

    float acc[ACC_FLOATS];
    if (visible) {
        const float4* row = reinterpret_cast<const float4*>(accum + (size_t)idx * ACC_FLOATS);
#pragma unroll
        for (int k = 0; k < ACC_FLOATS / 4; ++k) {
            const float4 t = row[k];
            acc[4 * k + 0] = t.x; acc[4 * k + 1] = t.y; acc[4 * k + 2] = t.z; acc[4 * k + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < ACC_FLOATS; ++k) acc[k] = 0.f;
    }

    // direct per-Gaussian outputs of the blend backward
    if (!RAW || dL_dmean2D) {
        dL_dmean2D[3 * idx + 0] = acc[A_M2X];
        dL_dmean2D[3 * idx + 1] = acc[A_M2Y];
        dL_dmean2D[3 * idx + 2] = acc[A_M2Z];
    }
    if (RAW) {
        if (!visible) {
            if (s_row)      // the staged row goes back as this Gaussian's (zero) SH gradient
                for (int c = 0; c < 3 * (M - 1); ++c) s_row[c] = 0.f;
            return;
        }
        const float op = gb_sigmoid(raw.opacity[idx]);
        raw.g_opacity[idx] += acc[A_OPAC] * ((1.0f - op) * op);
        if (acc[A_ROUGH] != 0.f) {
            const float y = gb_sigmoid(raw.roughness[idx]);
            raw.g_roughness[idx] += acc[A_ROUGH] * ((1.0f - y) * y);
        }
        if (acc[A_METAL] != 0.f) {
            const float y = gb_sigmoid(raw.metallic[idx]);
            raw.g_metallic[idx] += acc[A_METAL] * ((1.0f - y) * y);
        }
        if (acc[A_ALB] != 0.f || acc[A_ALB + 1] != 0.f || acc[A_ALB + 2] != 0.f) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float y = gb_sigmoid(raw.albedo[3 * idx + c]);
                raw.g_albedo[3 * idx + c] += acc[A_ALB + c] * ((1.0f - y) * y);
            }
        }
        if (acc[A_NRM] != 0.f || acc[A_NRM + 1] != 0.f || acc[A_NRM + 2] != 0.f) {
            const float v[3] = {raw.normal[3 * idx], raw.normal[3 * idx + 1], raw.normal[3 * idx + 2]};
            const float g[3] = {acc[A_NRM], acc[A_NRM + 1], acc[A_NRM + 2]};
            float o[3];
            gb_normalize_bwd(v, g, 3, o);
#pragma unroll
            for (int c = 0; c < 3; ++c) raw.g_normal[3 * idx + c] += o[c];
        }
    } else {
    if (dL_dconic_out) {
        *reinterpret_cast<float4*>(dL_dconic_out + 4 * (size_t)idx) = make_float4(acc[A_CX], acc[A_CY], 0.f, acc[A_CW]);
    }
    dL_dopacity[idx] = acc[A_OPAC];
    dL_droughness[idx] = acc[A_ROUGH];
    dL_dmetallic[idx] = acc[A_METAL];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        dL_dcolor[3 * idx + c] = acc[A_COL + c];
        dL_dnormal[3 * idx + c] = acc[A_NRM + c];
        dL_dalbedo[3 * idx + c] = acc[A_ALB + c];
    }
    }

    if (!visible) {
#pragma unroll
        for (int c = 0; c < 3; ++c) dL_dmean3D[3 * idx + c] = 0.f;
#pragma unroll
        for (int c = 0; c < 6; ++c) dL_dcov3D[6 * idx + c] = 0.f;
        if (dL_dsh)
            for (int c = 0; c < 3 * M; ++c) dL_dsh[(size_t)idx * 3 * M + c] = 0.f;
        if (dL_dscale)
            for (int c = 0; c < 3; ++c) dL_dscale[3 * idx + c] = 0.f;
        if (dL_drot)
            *reinterpret_cast<float4*>(dL_drot + 4 * (size_t)idx) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }

    // ------------------------------------------------------------------ cov2D backward
    const float* cov3D = cov3Ds + 6 * (size_t)idx;
    const float3 mean = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    const float3 dL_dconic = {acc[A_CX], acc[A_CY], acc[A_CW]};
    const float dL_ddepth = acc[A_DEPTH];
    float3 t = xform_point_4x3(mean, view_matrix);

    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = fminf(limx, fmaxf(-limx, txtz)) * t.z;
    t.y = fminf(limy, fmaxf(-limy, tytz)) * t.z;
    const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
    const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

    Mat3 J = mat3_cols(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z), 0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z), 0, 0,
                       0);
    Mat3 Wm = mat3_cols(view_matrix[0], view_matrix[4], view_matrix[8], view_matrix[1], view_matrix[5],
                        view_matrix[9], view_matrix[2], view_matrix[6], view_matrix[10]);
    Mat3 Vrk = mat3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    Mat3 T = mat3_mul(Wm, J);
    Mat3 cov2D = mat3_mul(mat3_mul(mat3_transpose(T), mat3_transpose(Vrk)), T);

    const float a = cov2D.m[0][0] += 0.3f;
    const float b = cov2D.m[0][1];
    const float c = cov2D.m[1][1] += 0.3f;

    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

    float dcov[6];
    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
        dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
        dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

        dcov[0] = (T.m[0][0] * T.m[0][0] * dL_da + T.m[0][0] * T.m[1][0] * dL_db + T.m[1][0] * T.m[1][0] * dL_dc);
        dcov[3] = (T.m[0][1] * T.m[0][1] * dL_da + T.m[0][1] * T.m[1][1] * dL_db + T.m[1][1] * T.m[1][1] * dL_dc);
        dcov[5] = (T.m[0][2] * T.m[0][2] * dL_da + T.m[0][2] * T.m[1][2] * dL_db + T.m[1][2] * T.m[1][2] * dL_dc);
        dcov[1] = 2 * T.m[0][0] * T.m[0][1] * dL_da + (T.m[0][0] * T.m[1][1] + T.m[0][1] * T.m[1][0]) * dL_db +
                  2 * T.m[1][0] * T.m[1][1] * dL_dc;
        dcov[2] = 2 * T.m[0][0] * T.m[0][2] * dL_da + (T.m[0][0] * T.m[1][2] + T.m[0][2] * T.m[1][0]) * dL_db +
                  2 * T.m[1][0] * T.m[1][2] * dL_dc;
        dcov[4] = 2 * T.m[0][2] * T.m[0][1] * dL_da + (T.m[0][1] * T.m[1][2] + T.m[0][2] * T.m[1][1]) * dL_db +
                  2 * T.m[1][1] * T.m[1][2] * dL_dc;
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) dcov[i] = 0;
    }
    if (!RAW) {
#pragma unroll
        for (int i = 0; i < 6; i++) dL_dcov3D[6 * (size_t)idx + i] = dcov[i];
    }

    const float dL_dT00 = 2 * (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_da +
                          (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_db;
    const float dL_dT01 = 2 * (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_da +
                          (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_db;
    const float dL_dT02 = 2 * (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_da +
                          (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_db;
    const float dL_dT10 = 2 * (T.m[1][0] * Vrk.m[0][0] + T.m[1][1] * Vrk.m[0][1] + T.m[1][2] * Vrk.m[0][2]) * dL_dc +
                          (T.m[0][0] * Vrk.m[0][0] + T.m[0][1] * Vrk.m[0][1] + T.m[0][2] * Vrk.m[0][2]) * dL_db;
    const float dL_dT11 = 2 * (T.m[1][0] * Vrk.m[1][0] + T.m[1][1] * Vrk.m[1][1] + T.m[1][2] * Vrk.m[1][2]) * dL_dc +
                          (T.m[0][0] * Vrk.m[1][0] + T.m[0][1] * Vrk.m[1][1] + T.m[0][2] * Vrk.m[1][2]) * dL_db;
    const float dL_dT12 = 2 * (T.m[1][0] * Vrk.m[2][0] + T.m[1][1] * Vrk.m[2][1] + T.m[1][2] * Vrk.m[2][2]) * dL_dc +
                          (T.m[0][0] * Vrk.m[2][0] + T.m[0][1] * Vrk.m[2][1] + T.m[0][2] * Vrk.m[2][2]) * dL_db;

    const float dL_dJ00 = Wm.m[0][0] * dL_dT00 + Wm.m[0][1] * dL_dT01 + Wm.m[0][2] * dL_dT02;
    const float dL_dJ02 = Wm.m[2][0] * dL_dT00 + Wm.m[2][1] * dL_dT01 + Wm.m[2][2] * dL_dT02;
    const float dL_dJ11 = Wm.m[1][0] * dL_dT10 + Wm.m[1][1] * dL_dT11 + Wm.m[1][2] * dL_dT12;
    const float dL_dJ12 = Wm.m[2][0] * dL_dT10 + Wm.m[2][1] * dL_dT11 + Wm.m[2][2] * dL_dT12;

    const float tz = 1.f / t.z;
    const float tz2 = tz * tz;
    const float tz3 = tz2 * tz;

    const float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
    const float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
    const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
                         (2 * h_y * t.y) * tz3 * dL_dJ12;

    // transformVec4x3Transpose + depth term (backward.cu:270-273)
    B3 dmean;
    dmean.x = view_matrix[0] * dL_dtx + view_matrix[1] * dL_dty + view_matrix[2] * dL_dtz;
    dmean.y = view_matrix[4] * dL_dtx + view_matrix[5] * dL_dty + view_matrix[6] * dL_dtz;
    dmean.z = view_matrix[8] * dL_dtx + view_matrix[9] * dL_dty + view_matrix[10] * dL_dtz;
    dmean.x += view_matrix[2] * dL_ddepth;
    dmean.y += view_matrix[6] * dL_ddepth;
    dmean.z += view_matrix[10] * dL_ddepth;

    // ------------------------------------------------------------------ mean2D -> mean3D
    {
        const float3 m = mean;
        const float4 m_hom = xform_point_4x4(m, proj);
        const float m_w = 1.0f / (m_hom.w + 0.0000001f);
        const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
        const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
        const float gx = acc[A_M2X], gy = acc[A_M2Y];
        B3 d2;
        d2.x = (proj[0] * m_w - proj[3] * mul1) * gx + (proj[1] * m_w - proj[3] * mul2) * gy;
        d2.y = (proj[4] * m_w - proj[7] * mul1) * gx + (proj[5] * m_w - proj[7] * mul2) * gy;
        d2.z = (proj[8] * m_w - proj[11] * mul1) * gx + (proj[9] * m_w - proj[11] * mul2) * gy;
        dmean += d2;
    }

    // ------------------------------------------------------------------ colour -> SH (+ dir term)
    if (shs) {
        const B3 pos = {mean.x, mean.y, mean.z};
        const B3 cam = {campos[0], campos[1], campos[2]};
        const B3 dir_orig = {pos.x - cam.x, pos.y - cam.y, pos.z - cam.z};
        const float len = sqrtf(dotB(dir_orig, dir_orig));
        const B3 dir = {dir_orig.x / len, dir_orig.y / len, dir_orig.z / len};
        // RAW: coefficient 0 lives in f_dc, coefficient i >= 1 at f_rest + (idx*(M-1) + i-1)*3 (or in the staged row)
        const bool staged = RAW && s_row != nullptr;
        const float* shp = staged ? s_row - 3 : (RAW ? raw.f_rest + ((size_t)idx * (M - 1) - 1) * 3 : shs + (size_t)idx * M * 3);
        auto SH = [&](int i) -> B3 { return {shp[3 * i + 0], shp[3 * i + 1], shp[3 * i + 2]}; };
        float* dshp = staged ? s_row - 3 : (RAW ? raw.g_f_rest + ((size_t)idx * (M - 1) - 1) * 3 : dL_dsh + (size_t)idx * M * 3);
        auto DSH = [&](int i, const B3& v) {
            if (RAW && i == 0) {
                float* d = raw.g_f_dc + (size_t)idx * 3;
                d[0] += v.x;
                d[1] += v.y;
                d[2] += v.z;
            } else if (RAW && !staged) {
                dshp[3 * i + 0] += v.x;
                dshp[3 * i + 1] += v.y;
                dshp[3 * i + 2] += v.z;
            } else {
                dshp[3 * i + 0] = v.x;
                dshp[3 * i + 1] = v.y;
                dshp[3 * i + 2] = v.z;
            }
        };
        const uchar4 cl = *reinterpret_cast<const uchar4*>(clamped + 4 * (size_t)idx);
        B3 dL_dRGB = {acc[A_COL + 0], acc[A_COL + 1], acc[A_COL + 2]};
        dL_dRGB.x *= cl.x ? 0 : 1;
        dL_dRGB.y *= cl.y ? 0 : 1;
        dL_dRGB.z *= cl.z ? 0 : 1;

        // pass 1: everything that READS the coefficients (the view-direction term), ...
        B3 dRGBdx = {0, 0, 0}, dRGBdy = {0, 0, 0}, dRGBdz = {0, 0, 0};
        const float x = dir.x, y = dir.y, z = dir.z;
        const float xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, yz = y * z, xz = x * z;
        if (D > 0) {
            dRGBdx = -bSH_C1 * SH(3);
            dRGBdy = -bSH_C1 * SH(1);
            dRGBdz = bSH_C1 * SH(2);
            if (D > 1) {
                dRGBdx += bSH_C2[0] * y * SH(4) + bSH_C2[2] * 2.f * -x * SH(6) + bSH_C2[3] * z * SH(7) +
                          bSH_C2[4] * 2.f * x * SH(8);
                dRGBdy += bSH_C2[0] * x * SH(4) + bSH_C2[1] * z * SH(5) + bSH_C2[2] * 2.f * -y * SH(6) +
                          bSH_C2[4] * 2.f * -y * SH(8);
                dRGBdz += bSH_C2[1] * y * SH(5) + bSH_C2[2] * 2.f * 2.f * z * SH(6) + bSH_C2[3] * x * SH(7);
                if (D > 2) {
                    dRGBdx += (bSH_C3[0] * SH(9) * 3.f * 2.f * xy + bSH_C3[1] * SH(10) * yz +
                               bSH_C3[2] * SH(11) * -2.f * xy + bSH_C3[3] * SH(12) * -3.f * 2.f * xz +
                               bSH_C3[4] * SH(13) * (-3.f * xx + 4.f * zz - yy) + bSH_C3[5] * SH(14) * 2.f * xz +
                               bSH_C3[6] * SH(15) * 3.f * (xx - yy));
                    dRGBdy += (bSH_C3[0] * SH(9) * 3.f * (xx - yy) + bSH_C3[1] * SH(10) * xz +
                               bSH_C3[2] * SH(11) * (-3.f * yy + 4.f * zz - xx) +
                               bSH_C3[3] * SH(12) * -3.f * 2.f * yz + bSH_C3[4] * SH(13) * -2.f * xy +
                               bSH_C3[5] * SH(14) * -2.f * yz + bSH_C3[6] * SH(15) * -3.f * 2.f * xy);
                    dRGBdz += (bSH_C3[1] * SH(10) * xy + bSH_C3[2] * SH(11) * 4.f * 2.f * yz +
                               bSH_C3[3] * SH(12) * 3.f * (2.f * zz - xx - yy) +
                               bSH_C3[4] * SH(13) * 4.f * 2.f * xz + bSH_C3[5] * SH(14) * (xx - yy));
                }
            }
        }
        // ... pass 2: the coefficient gradients (the staged row is overwritten in place from here on)
        DSH(0, bSH_C0 * dL_dRGB);
        int written = 1;
        if (D > 0) {
            DSH(1, (-bSH_C1 * y) * dL_dRGB);
            DSH(2, (bSH_C1 * z) * dL_dRGB);
            DSH(3, (-bSH_C1 * x) * dL_dRGB);
            written = 4;
            if (D > 1) {
                DSH(4, (bSH_C2[0] * xy) * dL_dRGB);
                DSH(5, (bSH_C2[1] * yz) * dL_dRGB);
                DSH(6, (bSH_C2[2] * (2.f * zz - xx - yy)) * dL_dRGB);
                DSH(7, (bSH_C2[3] * xz) * dL_dRGB);
                DSH(8, (bSH_C2[4] * (xx - yy)) * dL_dRGB);
                written = 9;
                if (D > 2) {
                    DSH(9, (bSH_C3[0] * y * (3.f * xx - yy)) * dL_dRGB);
                    DSH(10, (bSH_C3[1] * xy * z) * dL_dRGB);
                    DSH(11, (bSH_C3[2] * y * (4.f * zz - xx - yy)) * dL_dRGB);
                    DSH(12, (bSH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL_dRGB);
                    DSH(13, (bSH_C3[4] * x * (4.f * zz - xx - yy)) * dL_dRGB);
                    DSH(14, (bSH_C3[5] * z * (xx - yy)) * dL_dRGB);
                    DSH(15, (bSH_C3[6] * x * (xx - 3.f * yy)) * dL_dRGB);
                    written = 16;
                }
            }
        }
        // coefficients above the active degree get no gradient (zeros in the reference's pre-filled tensor)
        if (!RAW || staged)
            for (int i = written; i < M; ++i) DSH(i, B3{0.f, 0.f, 0.f});

        const B3 dL_ddir = {dotB(dRGBdx, dL_dRGB), dotB(dRGBdy, dL_dRGB), dotB(dRGBdz, dL_dRGB)};
        // dnormvdv (auxiliary.h:120-131)
        const B3 v = dir_orig, dv = dL_ddir;
        const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
        const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
        B3 dn;
        dn.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
        dn.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
        dn.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
        dmean += dn;
    }
    if (RAW) {
        raw.g_xyz[3 * idx + 0] += dmean.x;
        raw.g_xyz[3 * idx + 1] += dmean.y;
        raw.g_xyz[3 * idx + 2] += dmean.z;
    } else {
        dL_dmean3D[3 * idx + 0] = dmean.x;
        dL_dmean3D[3 * idx + 1] = dmean.y;
        dL_dmean3D[3 * idx + 2] = dmean.z;
    }

    // ------------------------------------------------------------------ cov3D -> scale, quaternion
    if (scales) {
        float4 q = *reinterpret_cast<const float4*>(rotations + 4 * (size_t)idx);
        const float4 q_raw = q;
        float3 sc = {scales[3 * idx], scales[3 * idx + 1], scales[3 * idx + 2]};
        if (RAW) {   // the getters, as preprocess_kernel<RAW> applies them
            sc = make_float3(expf(sc.x), expf(sc.y), expf(sc.z));
            const float qn = fmaxf(torch_norm_inner4(q.x, q.y, q.z, q.w), 1e-12f);
            q = make_float4(q.x / qn, q.y / qn, q.z / qn, q.w / qn);
        }
        const float r = q.x, x = q.y, y = q.z, z = q.w;
        Mat3 R = mat3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                           2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                           2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
        Mat3 S = mat3_cols(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
        const float3 s = {scale_modifier * sc.x, scale_modifier * sc.y, scale_modifier * sc.z};
        S.m[0][0] = s.x;
        S.m[1][1] = s.y;
        S.m[2][2] = s.z;
        Mat3 Mm = mat3_mul(S, R);
        Mat3 dL_dSigma = mat3_cols(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                                   0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
        // dL_dM = 2 * M * dL_dSigma  (scalar*matrix first, then product)
        Mat3 M2;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) M2.m[cc][rr] = 2.0f * Mm.m[cc][rr];
        Mat3 dL_dM = mat3_mul(M2, dL_dSigma);
        Mat3 Rt = mat3_transpose(R);
        Mat3 dL_dMt = mat3_transpose(dL_dM);

        const float ds0 = Rt.m[0][0] * dL_dMt.m[0][0] + Rt.m[0][1] * dL_dMt.m[0][1] + Rt.m[0][2] * dL_dMt.m[0][2];
        const float ds1 = Rt.m[1][0] * dL_dMt.m[1][0] + Rt.m[1][1] * dL_dMt.m[1][1] + Rt.m[1][2] * dL_dMt.m[1][2];
        const float ds2 = Rt.m[2][0] * dL_dMt.m[2][0] + Rt.m[2][1] * dL_dMt.m[2][1] + Rt.m[2][2] * dL_dMt.m[2][2];
        if (RAW) {   // exp backward: grad * exp(x)
            raw.g_log_scale[3 * idx + 0] += ds0 * sc.x;
            raw.g_log_scale[3 * idx + 1] += ds1 * sc.y;
            raw.g_log_scale[3 * idx + 2] += ds2 * sc.z;
        } else {
            dL_dscale[3 * idx + 0] = ds0;
            dL_dscale[3 * idx + 1] = ds1;
            dL_dscale[3 * idx + 2] = ds2;
        }

#pragma unroll
        for (int k = 0; k < 3; ++k) {
            dL_dMt.m[0][k] *= s.x;
            dL_dMt.m[1][k] *= s.y;
            dL_dMt.m[2][k] *= s.z;
        }
        float4 dq;
        dq.x = 2 * z * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * y * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) +
               2 * x * (dL_dMt.m[1][2] - dL_dMt.m[2][1]);
        dq.y = 2 * y * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * z * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) +
               2 * r * (dL_dMt.m[1][2] - dL_dMt.m[2][1]) - 4 * x * (dL_dMt.m[2][2] + dL_dMt.m[1][1]);
        dq.z = 2 * x * (dL_dMt.m[1][0] + dL_dMt.m[0][1]) + 2 * r * (dL_dMt.m[2][0] - dL_dMt.m[0][2]) +
               2 * z * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * y * (dL_dMt.m[2][2] + dL_dMt.m[0][0]);
        dq.w = 2 * r * (dL_dMt.m[0][1] - dL_dMt.m[1][0]) + 2 * x * (dL_dMt.m[2][0] + dL_dMt.m[0][2]) +
               2 * y * (dL_dMt.m[1][2] + dL_dMt.m[2][1]) - 4 * z * (dL_dMt.m[1][1] + dL_dMt.m[0][0]);
        // no quaternion-normalisation Jacobian, as in the reference (backward.cu:345)
        if (RAW) {
            const float v[4] = {q_raw.x, q_raw.y, q_raw.z, q_raw.w};
            const float g[4] = {dq.x, dq.y, dq.z, dq.w};
            float o[4];
            gb_normalize_bwd(v, g, 4, o);
#pragma unroll
            for (int k = 0; k < 4; ++k) raw.g_rot[4 * (size_t)idx + k] += o[k];
        } else {
            *reinterpret_cast<float4*>(dL_drot + 4 * (size_t)idx) = dq;
        }
    }
}

// The kernel. RAW fast path (M = 16, i.e. 45 f_rest floats per Gaussian, full warps, 16-B aligned tensors): a warp's 32
// rows are 5760 contiguous bytes — it moves them with 12 coalesced 16-B loads per lane into its own shared-memory
// region (flat layout: a 45-float row stride is conflict-free), the body reads its row and overwrites it with the SH
// gradient, and the warp adds the region onto the leaf's gradient with 16-B loads / stores. Per lane that is 36
// memory instructions instead of 135 scalar ones; the scalar kernel is bound by the LSU queue (ncu: lg_throttle +
// long_scoreboard at 12 % issue utilisation). Only __syncwarp is needed: no CTA barrier, 23 KB of shared memory.
template <bool RAW>
__global__ void GB_BOUNDS
gaussian_backward_kernel(const int P, const int D, const int M, const float* __restrict__ means3D,
                         const int* __restrict__ radii, const float* __restrict__ shs,
                         const uint8_t* __restrict__ clamped, const float* __restrict__ scales,
                         const float* __restrict__ rotations, const float scale_modifier,
                         const float* __restrict__ cov3Ds, const float* __restrict__ view_matrix,
                         const float* __restrict__ proj, const float* __restrict__ campos, const float h_x,
                         const float h_y, const float tan_fovx, const float tan_fovy,
                         const float* __restrict__ accum,
                         float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic_out,
                         float* __restrict__ dL_dopacity, float* __restrict__ dL_dcolor,
                         float* __restrict__ dL_dnormal, float* __restrict__ dL_dalbedo,
                         float* __restrict__ dL_droughness, float* __restrict__ dL_dmetallic,
                         float* __restrict__ dL_dmean3D, float* __restrict__ dL_dcov3D, float* __restrict__ dL_dsh,
                         float* __restrict__ dL_dscale, float* __restrict__ dL_drot, const RawGrads raw, const int stage_rows)
{
    pdl_enter();
    extern __shared__ float4 gb_stage4[];
    constexpr int ROW = 45, V4 = 32 * ROW / 4;                 // 360 float4 per warp
    const int idx = blockIdx.x * GB_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx0w = blockIdx.x * GB_THREADS + warp * 32;
    float* s_row = nullptr;
    bool fast = false;
    if (RAW) {
        fast = stage_rows && (idx0w + 32 <= P);                 // warp-uniform
        if (fast) {
            float4* sw = gb_stage4 + warp * V4;
            const float4* g4 = reinterpret_cast<const float4*>(raw.f_rest + (size_t)idx0w * ROW);
#pragma unroll
            for (int j = 0; j < (V4 + 31) / 32; ++j) {
                const int q = lane + 32 * j;
                if (q < V4) sw[q] = g4[q];
            }
            __syncwarp();
            s_row = reinterpret_cast<float*>(sw) + lane * ROW;
        }
    }
    if (idx < P)
        gaussian_backward_body<RAW>(idx, s_row, P, D, M, means3D, radii, shs, clamped, scales, rotations, scale_modifier,
                                    cov3Ds, view_matrix, proj, campos, h_x, h_y, tan_fovx, tan_fovy, accum, dL_dmean2D,
                                    dL_dconic_out, dL_dopacity, dL_dcolor, dL_dnormal, dL_dalbedo, dL_droughness,
                                    dL_dmetallic, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, raw);
    if (RAW && fast) {
        __syncwarp();
        const float4* sw = gb_stage4 + warp * V4;
        float4* o4 = reinterpret_cast<float4*>(raw.g_f_rest + (size_t)idx0w * ROW);
#pragma unroll
        for (int j = 0; j < (V4 + 31) / 32; ++j) {
            const int q = lane + 32 * j;
            if (q < V4) {
                const float4 v = sw[q];
                if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
                    float4 o = o4[q];
                    o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
                    o4[q] = o;
                }
            }
        }
    }
}

int launch_gaussian_backward(const GigsRasterBwd* a, const Layout& L, cudaStream_t st)
{
    const GigsCamera& c = a->cam;
    const char* g = (const char*)a->geom;
    const float focal_y = c.height / (2.0f * c.tan_fovy);
    const float focal_x = c.width / (2.0f * c.tan_fovx);
    const float* cov3D_ptr = a->cov3D_precomp ? a->cov3D_precomp : (const float*)(g + L.off.g_cov3D);
    GIGS_CUDA(launch_k(gaussian_backward_kernel<false>, dim3((a->P + GB_THREADS - 1) / GB_THREADS), dim3(GB_THREADS), (size_t)(0), st, 
        a->P, c.sh_degree, c.sh_coeffs, a->means3D, a->radii, a->shs, (const uint8_t*)(g + L.off.g_clamped), a->scales,
        a->rotations, c.scale_modifier, cov3D_ptr, c.viewmatrix, c.projmatrix, c.campos, focal_x, focal_y, c.tan_fovx,
        c.tan_fovy, a->accum, a->dL_dmean2D, a->dL_dconic, a->dL_dopacity, a->dL_dcolor, a->dL_dnormal, a->dL_dalbedo,
        a->dL_droughness, a->dL_dmetallic, a->dL_dmean3D, a->dL_dcov3D, a->dL_dsh, a->dL_dscale, a->dL_drot, RawGrads{}, 0));
    GIGS_LAUNCH_CHECK("gaussian_backward_kernel");
    return 0;
}

// First-stage frame: raw leaves in, gradients accumulated into the leaves' gradient tensors (see RawGrads).
int launch_gaussian_backward_raw(int P, const GigsCamera& c, const void* geom, const Layout& L, const int32_t* radii,
                                 const float* accum, const float* xyz, const float* f_dc, const float* log_scale,
                                 const float* rot, float* g_means2D, const RawGrads& raw, cudaStream_t st)
{
    const char* g = (const char*)geom;
    const float focal_y = c.height / (2.0f * c.tan_fovy);
    const float focal_x = c.width / (2.0f * c.tan_fovx);
    // staged SH rows: 45 floats per Gaussian (M = 16) and 16-B aligned leaves (GIGS_GB_STAGE=0 turns it off)
    static const bool stage_env = !(getenv("GIGS_GB_STAGE") && atoi(getenv("GIGS_GB_STAGE")) == 0);
    const int stage_rows = stage_env && c.sh_coeffs == 16 && raw.f_rest && raw.g_f_rest &&
                           ((((uintptr_t)raw.f_rest | (uintptr_t)raw.g_f_rest) & 15) == 0);
    const size_t smem = stage_rows ? (size_t)(GB_THREADS / 32) * (32 * 45 / 4) * sizeof(float4) : 0;
    GIGS_CUDA(launch_k(gaussian_backward_kernel<true>, dim3((P + GB_THREADS - 1) / GB_THREADS), dim3(GB_THREADS), (size_t)(smem), st, 
        P, c.sh_degree, c.sh_coeffs, xyz, radii, f_dc, (const uint8_t*)(g + L.off.g_clamped), log_scale, rot,
        c.scale_modifier, (const float*)(g + L.off.g_cov3D), c.viewmatrix, c.projmatrix, c.campos, focal_x, focal_y,
        c.tan_fovx, c.tan_fovy, accum, g_means2D, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
        nullptr, nullptr, nullptr, nullptr, raw, stage_rows));
    GIGS_LAUNCH_CHECK("gaussian_backward_kernel<RAW>");
    return 0;
}

}  // namespace gigs
