// Image loss of the first training stage: (1 - lambda) * L1 + lambda * (1 - SSIM) with its gradient,
// /root/reference/train.py:320-322 + utils/loss_utils.py:19-20,40-100. The reference runs 5 grouped 11x11
// convolutions forward (mu1, mu2, E[xx], E[yy], E[xy]), ~15 elementwise kernels, and autograd replays all of them
// backward; here the forward is ONE kernel (separable 11-tap Gaussian in shared memory on a 16x16 tile with a 5-pixel
// zero-padded halo, SSIM map, L1, per-CTA partial sums, and the three partial-derivative maps the backward needs) and
// the backward is ONE kernel (the same separable filter applied to the three derivative maps: the window is symmetric
// and the padding is zero, so the adjoint of the convolution is the convolution).
//
//   S = A1*A2 / (B1*B2),  A1 = 2 mu1 mu2 + C1, A2 = 2 s12 + C2, B1 = mu1^2 + mu2^2 + C1, B2 = s11 + s22 + C2,
//   s11 = E[xx] - mu1^2, s22 = E[yy] - mu2^2, s12 = E[xy] - mu1 mu2   (operation order of loss_utils.py:77-95)
//   dS/dE[xx] = -S / B2,  dS/dE[xy] = 2 A1 / (B1 B2),
//   dS/dmu1 (through s11 and s12 as well) = 2 mu2 (A2 - A1) / (B1 B2) + 2 mu1 S (1/B2 - 1/B1)
//   dS/dx(p) = sum_q w(q - p) [ dS/dmu1(q) + 2 x(p) dS/dE[xx](q) + y(p) dS/dE[xy](q) ]
// HBM-bound in principle (forward: 8 B read + 12 B written per element; backward: 20 B read + 4 B written), in
// practice bound by the shared-memory passes; 800x800x3 is 1.9 M elements, the pair of kernels ~50 us.
#include <cmath>
#include <mutex>
#include "common.cuh"

namespace gigs {

constexpr int SS_T = 16;            // tile
constexpr int SS_R = 5;             // window radius (window_size 11)
constexpr int SS_E = SS_T + 2 * SS_R;   // 26

__constant__ float c_ssim_w[11];

struct SsimW { float w[11]; };

// loss_utils.py:40-51: float32 tensor of exp(-(x-5)^2 / (2*1.5^2)) (evaluated in double), divided by its float32 sum
static SsimW ssim_window()
{
    SsimW W;
    float s = 0.f;
    for (int i = 0; i < 11; i++) {
        W.w[i] = (float)std::exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5));
    }
    // torch.sum over 11 floats: a plain left-to-right float sum is within one ulp of any order; the weights' own
    // rounding (2^-24 relative) is far below the test tolerance
    for (int i = 0; i < 11; i++) s += W.w[i];
    for (int i = 0; i < 11; i++) W.w[i] = W.w[i] / s;
    return W;
}

// Both filter kernels: 16x16 output tile, 64 threads (16 x 4). Every thread produces 4 adjacent outputs per pass from a
// 14-value register window (horizontal pass: 4 columns of one row; vertical pass: 4 rows of one column), so a filtered
// value costs 3.5 shared-memory loads instead of 11 — the first version (one output per thread, 256 threads) was bound
// by the shared-memory instruction queue (ncu: mio_throttle the top stall, 62 + 55 us at 800x800x3).
constexpr int SS_NT = 64;            // threads per CTA
constexpr int SS_HS = 20;            // row stride of the horizontally filtered arrays: 4 rows apart = 16 banks apart

// partial sums layout in scratch: float2 partial[n_ctas] (sum S, sum |x-y|)
__global__ void __launch_bounds__(SS_NT) ssim_forward_kernel(int C, int W, int H, const float* __restrict__ X,
                                                             const float* __restrict__ Y, float* __restrict__ dmu,
                                                             float* __restrict__ dxx, float* __restrict__ dxy,
                                                             float2* __restrict__ partial)
{
    pdl_enter();
    __shared__ float sx[SS_E][SS_E + 1], sy[SS_E][SS_E + 1];
    __shared__ float h[5][SS_E][SS_HS];
    __shared__ float red[2][SS_NT / 32];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const int tid = threadIdx.y * SS_T + threadIdx.x;
    const size_t plane = (size_t)c * W * H;
    for (int i = tid; i < SS_E * SS_E; i += SS_NT) {
        const int ly = i / SS_E, lx = i - ly * SS_E;
        const int gx = x0 + lx - SS_R, gy = y0 + ly - SS_R;
        float a = 0.f, b = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            a = X[plane + (size_t)gy * W + gx];
            b = Y[plane + (size_t)gy * W + gx];
        }
        sx[ly][lx] = a;
        sy[ly][lx] = b;
    }
    __syncthreads();
    float wk[11];
#pragma unroll
    for (int k = 0; k < 11; k++) wk[k] = c_ssim_w[k];
    // horizontal pass: 26 rows x 4 groups of 4 columns
    for (int it = tid; it < SS_E * 4; it += SS_NT) {
        const int ly = it >> 2, cg = (it & 3) * 4;
        float va[14], vb[14];
#pragma unroll
        for (int i = 0; i < 14; i++) { va[i] = sx[ly][cg + i]; vb[i] = sy[ly][cg + i]; }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float m1 = 0.f, m2 = 0.f, xx = 0.f, yy = 0.f, xy = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) {
                const float w = wk[k], a = va[j + k], b = vb[j + k];
                m1 = fmaf(w, a, m1);
                m2 = fmaf(w, b, m2);
                xx = fmaf(w, a * a, xx);
                yy = fmaf(w, b * b, yy);
                xy = fmaf(w, a * b, xy);
            }
            h[0][ly][cg + j] = m1; h[1][ly][cg + j] = m2; h[2][ly][cg + j] = xx; h[3][ly][cg + j] = yy; h[4][ly][cg + j] = xy;
        }
    }
    __syncthreads();
    // vertical pass: column threadIdx.x, rows 4*threadIdx.y .. +3
    const int lx = threadIdx.x, r0 = threadIdx.y * 4;
    float f[5][4];
#pragma unroll
    for (int q = 0; q < 5; q++) {
        float v[14];
#pragma unroll
        for (int i = 0; i < 14; i++) v[i] = h[q][r0 + i][lx];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) a = fmaf(wk[k], v[j + k], a);
            f[q][j] = a;
        }
    }
    float S = 0.f, l1 = 0.f;
    const int gx = x0 + lx;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int gy = y0 + r0 + j;
        if (gx < W && gy < H) {
            const float mu1 = f[0][j], mu2 = f[1][j], exx = f[2][j], eyy = f[3][j], exy = f[4][j];
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
            const float s11 = exx - mu1_sq, s22 = eyy - mu2_sq, s12 = exy - mu12;
            const float A1 = 2.f * mu12 + C1, A2 = 2.f * s12 + C2, B1 = mu1_sq + mu2_sq + C1, B2 = s11 + s22 + C2;
            const float Sv = (A1 * A2) / (B1 * B2);
            S += Sv;
            l1 += fabsf(sx[r0 + j + SS_R][lx + SS_R] - sy[r0 + j + SS_R][lx + SS_R]);
            if (dmu) {
                const float inv = 1.f / (B1 * B2);
                const size_t o = plane + (size_t)gy * W + gx;
                dmu[o] = 2.f * mu2 * (A2 - A1) * inv + 2.f * mu1 * Sv * (1.f / B2 - 1.f / B1);
                dxx[o] = -Sv / B2;
                dxy[o] = 2.f * A1 * inv;
            }
        }
    }
    // deterministic CTA reduction (fixed shuffle tree, fixed order over warps)
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        S += __shfl_xor_sync(0xffffffffu, S, o);
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = S; red[1][tid >> 5] = l1; }
    __syncthreads();
    if (tid == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < SS_NT / 32; i++) { a += red[0][i]; b += red[1][i]; }
        partial[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = make_float2(a, b);
    }
}

// one CTA: sums the per-CTA partials in a fixed order (double accumulators), writes loss_out[0..2] = loss, l1, ssim
__global__ void __launch_bounds__(1024) ssim_finish_kernel(int n, const float2* __restrict__ partial, double inv_n,
                                                           float lambda, float loss_scale, float* __restrict__ out,
                                                           int accumulate)
{
    pdl_enter();
    __shared__ double rs[32], rl[32];
    double s = 0.0, l = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) { s += (double)partial[i].x; l += (double)partial[i].y; }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        l += __shfl_xor_sync(0xffffffffu, l, o);
    }
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rl[threadIdx.x >> 5] = l; }
    __syncthreads();
    if (threadIdx.x == 0) {
        s = 0.0; l = 0.0;
        for (int i = 0; i < 32; i++) { s += rs[i]; l += rl[i]; }
        const float ssim = (float)(s * inv_n), l1 = (float)(l * inv_n);
        const float loss = loss_scale * ((1.f - lambda) * l1 + lambda * (1.f - ssim));
        out[0] = accumulate ? out[0] + loss : loss;
        out[1] = l1;
        out[2] = ssim;
    }
}

__global__ void __launch_bounds__(SS_NT) ssim_backward_kernel(int C, int W, int H, const float* __restrict__ X,
                                                              const float* __restrict__ Y, const float* __restrict__ dmu,
                                                              const float* __restrict__ dxx, const float* __restrict__ dxy,
                                                              float k_ssim, float k_l1, const float* __restrict__ upstream,
                                                              float* __restrict__ grad, int accumulate)
{
    pdl_enter();
    __shared__ float s[3][SS_E][SS_E + 1];
    __shared__ float h[3][SS_E][SS_HS];
    const int c = blockIdx.z;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    const int tid = threadIdx.y * SS_T + threadIdx.x;
    const size_t plane = (size_t)c * W * H;
    for (int i = tid; i < SS_E * SS_E; i += SS_NT) {
        const int ly = i / SS_E, lx = i - ly * SS_E;
        const int gx = x0 + lx - SS_R, gy = y0 + ly - SS_R;
        float a = 0.f, b = 0.f, d = 0.f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t o = plane + (size_t)gy * W + gx;
            a = dmu[o]; b = dxx[o]; d = dxy[o];
        }
        s[0][ly][lx] = a; s[1][ly][lx] = b; s[2][ly][lx] = d;
    }
    __syncthreads();
    float wk[11];
#pragma unroll
    for (int k = 0; k < 11; k++) wk[k] = c_ssim_w[k];
    for (int it = tid; it < SS_E * 4; it += SS_NT) {
        const int ly = it >> 2, cg = (it & 3) * 4;
#pragma unroll
        for (int q = 0; q < 3; q++) {
            float v[14];
#pragma unroll
            for (int i = 0; i < 14; i++) v[i] = s[q][ly][cg + i];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float a = 0.f;
#pragma unroll
                for (int k = 0; k < 11; k++) a = fmaf(wk[k], v[j + k], a);
                h[q][ly][cg + j] = a;
            }
        }
    }
    __syncthreads();
    const int lx = threadIdx.x, r0 = threadIdx.y * 4;
    float f[3][4];
#pragma unroll
    for (int q = 0; q < 3; q++) {
        float v[14];
#pragma unroll
        for (int i = 0; i < 14; i++) v[i] = h[q][r0 + i][lx];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < 11; k++) a = fmaf(wk[k], v[j + k], a);
            f[q][j] = a;
        }
    }
    const int gx = x0 + lx;
    const float up = upstream ? upstream[0] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int gy = y0 + r0 + j;
        if (gx >= W || gy >= H) continue;
        const size_t o = plane + (size_t)gy * W + gx;
        const float x = X[o], y = Y[o];
        const float dS = f[0][j] + 2.f * x * f[1][j] + y * f[2][j];
        const float df = x - y;
        const float sg = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
        float g = k_ssim * dS + k_l1 * sg;
        if (upstream) g *= up;
        grad[o] = accumulate ? grad[o] + g : g;
    }
}

// ---- geometry terms of the first-stage loss (train.py:323-328) ------------------------------------------------------
//   normal_loss = F.l1_loss(normal_map[:, mask], normal_map_from_depth[:, mask])       (mask = normal_from_depth_mask)
//   normal_tv   = get_tv_loss(gt_image, normal_map, pad=1, step=1)                      (train.py:83-100: edge-aware TV)
// One pixel per thread; partial[cta] = (sum |n - nd| over masked pixels, masked pixel count, sum w_h*dh^2, sum w_w*dw^2).
__device__ __forceinline__ float tv_edge_weight(const float* __restrict__ gt, size_t N, size_t a, size_t b)
{
    const float m = (fabsf(gt[b] - gt[a]) + fabsf(gt[N + b] - gt[N + a]) + fabsf(gt[2 * N + b] - gt[2 * N + a])) / 3.f;
    return expf(-m);
}

__global__ void __launch_bounds__(256) normal_loss_forward_kernel(int W, int H, const float* __restrict__ nm,
                                                                  const float* __restrict__ nd,
                                                                  const uint8_t* __restrict__ mask,
                                                                  const float* __restrict__ gt, float4* __restrict__ partial)
{
    pdl_enter();
    __shared__ float red[4][8];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const size_t N = (size_t)W * H;
    float l1 = 0.f, cnt = 0.f, th = 0.f, tw = 0.f;
    if (x < W && y < H) {
        const size_t o = (size_t)y * W + x;
        const float n0 = nm[o], n1 = nm[N + o], n2 = nm[2 * N + o];
        if (!mask || mask[o]) {
            l1 = fabsf(n0 - nd[o]) + fabsf(n1 - nd[N + o]) + fabsf(n2 - nd[2 * N + o]);
            cnt = 1.f;
        }
        if (y + 1 < H) {
            const size_t b = o + W;
            const float d0 = nm[b] - n0, d1 = nm[N + b] - n1, d2 = nm[2 * N + b] - n2;
            th = (d0 * d0 + d1 * d1 + d2 * d2) * tv_edge_weight(gt, N, o, b);
        }
        if (x + 1 < W) {
            const size_t b = o + 1;
            const float d0 = nm[b] - n0, d1 = nm[N + b] - n1, d2 = nm[2 * N + b] - n2;
            tw = (d0 * d0 + d1 * d1 + d2 * d2) * tv_edge_weight(gt, N, o, b);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        l1 += __shfl_xor_sync(0xffffffffu, l1, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        th += __shfl_xor_sync(0xffffffffu, th, o);
        tw += __shfl_xor_sync(0xffffffffu, tw, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = l1; red[1][threadIdx.x >> 5] = cnt;
        red[2][threadIdx.x >> 5] = th; red[3][threadIdx.x >> 5] = tw;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i = 0; i < 8; i++) { a.x += red[0][i]; a.y += red[1][i]; a.z += red[2][i]; a.w += red[3][i]; }
        partial[blockIdx.y * gridDim.x + blockIdx.x] = a;
    }
}

// scal[0] = normal_weight * loss_scale / (3 * count)  (NaN when the mask is empty: F.l1_loss of an empty selection)
__global__ void __launch_bounds__(1024) normal_loss_finish_kernel(int n, const float4* __restrict__ partial, int W, int H,
                                                                  float normal_weight, float tv_weight, float loss_scale,
                                                                  float* __restrict__ out, int accumulate,
                                                                  float* __restrict__ scal)
{
    pdl_enter();
    __shared__ double r[4][32];
    double a = 0, b = 0, c = 0, d = 0;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const float4 p = partial[i];
        a += (double)p.x; b += (double)p.y; c += (double)p.z; d += (double)p.w;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if ((threadIdx.x & 31) == 0) {
        r[0][threadIdx.x >> 5] = a; r[1][threadIdx.x >> 5] = b; r[2][threadIdx.x >> 5] = c; r[3][threadIdx.x >> 5] = d;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = b = c = d = 0;
        for (int i = 0; i < 32; i++) { a += r[0][i]; b += r[1][i]; c += r[2][i]; d += r[3][i]; }
        const float l1 = (float)(a / (3.0 * b));       // 0/0 = NaN like the reference's mean over nothing
        const double nh = 3.0 * (double)(H - 1) * W, nw = 3.0 * (double)H * (W - 1);
        const float tv = (float)(c / nh) + (float)(d / nw);   // mean over an empty tensor (H or W == 1) is NaN too
        if (out) {
            const float loss = loss_scale * (normal_weight * l1 + tv_weight * tv);
            out[0] = accumulate ? out[0] + loss : loss;
            out[1] = l1;
            out[2] = tv;
        }
        scal[0] = (float)((double)normal_weight * (double)loss_scale / (3.0 * b));
    }
}

__global__ void __launch_bounds__(256) normal_loss_backward_kernel(int W, int H, const float* __restrict__ nm,
                                                                   const float* __restrict__ nd,
                                                                   const uint8_t* __restrict__ mask,
                                                                   const float* __restrict__ gt,
                                                                   const float* __restrict__ scal, float kh, float kw,
                                                                   const float* __restrict__ upstream,
                                                                   float* __restrict__ grad, int accumulate)
{
    pdl_enter();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const size_t N = (size_t)W * H, o = (size_t)y * W + x;
    const float k1 = scal[0];
    const bool m = !mask || mask[o];
    // the four edges of this pixel: weight * (2/count) folded into kh / kw
    const float wu = y > 0 ? kh * tv_edge_weight(gt, N, o - W, o) : 0.f;
    const float wd = y + 1 < H ? kh * tv_edge_weight(gt, N, o, o + W) : 0.f;
    const float wl = x > 0 ? kw * tv_edge_weight(gt, N, o - 1, o) : 0.f;
    const float wr = x + 1 < W ? kw * tv_edge_weight(gt, N, o, o + 1) : 0.f;
    const float up = upstream ? upstream[0] : 1.f;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const size_t q = c * N + o;
        const float v = nm[q];
        float g = 0.f;
        if (m) {
            const float df = v - nd[q];
            g = df > 0.f ? k1 : (df < 0.f ? -k1 : 0.f);
        }
        if (y > 0) g += wu * (v - nm[q - W]);
        if (y + 1 < H) g -= wd * (nm[q + W] - v);
        if (x > 0) g += wl * (v - nm[q - 1]);
        if (x + 1 < W) g -= wr * (nm[q + 1] - v);
        g *= up;
        grad[q] = accumulate ? grad[q] + g : g;
    }
}

}  // namespace gigs

using namespace gigs;

extern "C" {

int gigs_image_loss(int32_t C, int32_t W, int32_t H, const float* image, const float* gt, float lambda_dssim,
                    float loss_scale, void* scratch, uint64_t* scratch_bytes, float* loss_out, int32_t accumulate_loss,
                    float* grad_image, int32_t accumulate_grad, const float* upstream, void* stream)
{
    if (C <= 0 || W <= 0 || H <= 0 || !scratch_bytes) { set_error("gigs_image_loss: bad arguments"); return -1; }
    const dim3 grid((W + SS_T - 1) / SS_T, (H + SS_T - 1) / SS_T, C);
    const uint64_t n_cta = (uint64_t)grid.x * grid.y * grid.z;
    const uint64_t n = (uint64_t)C * W * H;
    const uint64_t maps = grad_image ? align_up(n * 4, 256) : 0;
    const uint64_t need = 3 * maps + align_up(n_cta * 8, 256);
    if (!scratch) { *scratch_bytes = 3 * align_up(n * 4, 256) + align_up(n_cta * 8, 256); return 0; }
    if (*scratch_bytes < need || !image || !gt || (!loss_out && !grad_image)) {
        set_error("gigs_image_loss: scratch too small (%llu < %llu) or NULL input", (unsigned long long)*scratch_bytes,
                  (unsigned long long)need);
        return -1;
    }
    if (grid.z > 65535 || grid.y > 65535) { set_error("gigs_image_loss: image too large"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    // the window is a constant of the algorithm: uploaded once per device (blocking copy, so that no stream of that
    // device can run the kernel before the constant is there), under a mutex
    {
        static std::mutex w_mutex;
        static unsigned long long w_devices = 0ull;
        int dev = 0;
        GIGS_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lock(w_mutex);
        if (dev < 0 || dev >= 64 || !((w_devices >> dev) & 1ull)) {
            const SsimW Wt = ssim_window();
            GIGS_CUDA(cudaMemcpyToSymbol(c_ssim_w, Wt.w, sizeof(Wt.w), 0, cudaMemcpyHostToDevice));
            if (dev >= 0 && dev < 64) w_devices |= 1ull << dev;
        }
    }
    char* base = (char*)scratch;
    float* dmu = grad_image ? (float*)base : nullptr;
    float* dxx = grad_image ? (float*)(base + maps) : nullptr;
    float* dxy = grad_image ? (float*)(base + 2 * maps) : nullptr;
    float2* partial = (float2*)(base + 3 * maps);
    ProfScope prof(27, st);
    GIGS_CUDA(launch_k(ssim_forward_kernel, dim3(grid), dim3(dim3(SS_T, 4)), (size_t)(0), st, C, W, H, image, gt, dmu, dxx, dxy, partial));
    GIGS_LAUNCH_CHECK("ssim_forward_kernel");
    if (loss_out) {
        GIGS_CUDA(launch_k(ssim_finish_kernel, dim3(1), dim3(1024), (size_t)(0), st, (int)n_cta, partial, 1.0 / (double)n, lambda_dssim, loss_scale, loss_out,
                                               accumulate_loss));
        GIGS_LAUNCH_CHECK("ssim_finish_kernel");
    }
    if (grad_image) {
        const float k_ssim = (float)(-(double)loss_scale * (double)lambda_dssim / (double)n);
        const float k_l1 = (float)((double)loss_scale * (1.0 - (double)lambda_dssim) / (double)n);
        GIGS_CUDA(launch_k(ssim_backward_kernel, dim3(grid), dim3(dim3(SS_T, 4)), (size_t)(0), st, C, W, H, image, gt, dmu, dxx, dxy, k_ssim, k_l1, upstream,
                                                                grad_image, accumulate_grad));
        GIGS_LAUNCH_CHECK("ssim_backward_kernel");
    }
    return 0;
}

int gigs_normal_loss(int32_t W, int32_t H, const float* normal_map, const float* normal_from_depth, const uint8_t* mask,
                     const float* gt_image, float normal_weight, float tv_weight, float loss_scale, void* scratch,
                     uint64_t* scratch_bytes, float* loss_out, int32_t accumulate_loss, float* grad_normal,
                     int32_t accumulate_grad, const float* upstream, void* stream)
{
    if (W <= 0 || H <= 0 || !scratch_bytes) { set_error("gigs_normal_loss: bad arguments"); return -1; }
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    const uint64_t n_cta = (uint64_t)grid.x * grid.y;
    const uint64_t need = align_up(n_cta * 16, 256) + 256;
    if (!scratch) { *scratch_bytes = need; return 0; }
    if (*scratch_bytes < need || !normal_map || !normal_from_depth || !gt_image || (!loss_out && !grad_normal)) {
        set_error("gigs_normal_loss: scratch too small or NULL input");
        return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float4* partial = (float4*)scratch;
    float* scal = (float*)((char*)scratch + align_up(n_cta * 16, 256));
    ProfScope prof(28, st);
    GIGS_CUDA(launch_k(normal_loss_forward_kernel, dim3(grid), dim3(256), (size_t)(0), st, W, H, normal_map, normal_from_depth, mask, gt_image, partial));
    GIGS_LAUNCH_CHECK("normal_loss_forward_kernel");
    GIGS_CUDA(launch_k(normal_loss_finish_kernel, dim3(1), dim3(1024), (size_t)(0), st, (int)n_cta, partial, W, H, normal_weight, tv_weight, loss_scale, loss_out,
                                                  accumulate_loss, scal));
    GIGS_LAUNCH_CHECK("normal_loss_finish_kernel");
    if (grad_normal) {
        const float kh = (float)(2.0 * (double)tv_weight * (double)loss_scale / (3.0 * (double)(H - 1) * (double)W));
        const float kw = (float)(2.0 * (double)tv_weight * (double)loss_scale / (3.0 * (double)H * (double)(W - 1)));
        GIGS_CUDA(launch_k(normal_loss_backward_kernel, dim3(grid), dim3(256), (size_t)(0), st, W, H, normal_map, normal_from_depth, mask, gt_image, scal, kh, kw,
                                                          upstream, grad_normal, accumulate_grad));
        GIGS_LAUNCH_CHECK("normal_loss_backward_kernel");
    }
    return 0;
}

}  // extern "C"
