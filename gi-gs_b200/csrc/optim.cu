// Fused multi-group Adam: the optimiser step of /root/reference/train.py:516-523
// (`gaussians.optimizer.step(); zero_grad(); light_optimizer.step(); zero_grad(); cubemap.clamp_(min=0)`), i.e.
// torch.optim.Adam over the 10 parameter groups of scene/gaussian_model.py:318-359 plus the light's base cubemap, as
// ONE launch over every group instead of ~10 foreach launches per optimiser plus the gradient clears.
//
// HBM bound: per element it reads param, grad, exp_avg, exp_avg_sq and writes param, exp_avg, exp_avg_sq (28 B, +4 B
// when the gradient is cleared in the same pass). A group whose gradient is KNOWN to be zero (grad == NULL: in the PBR
// stage the reference's backward returns all-zero gradients for everything but the materials, and the fused frame
// never writes them) skips the gradient read and clear: 24 B. The arithmetic is torch's `_multi_tensor_adam` /
// `_single_tensor_adam` sequence, operation by operation, with explicitly rounded intrinsics so that nvcc's
// contraction choices cannot change a bit:
//   exp_avg    = fma(1-beta1, grad - exp_avg, exp_avg)                     (lerp_, weight < 0.5)
//   exp_avg_sq = fma((1-beta2)*grad, grad, exp_avg_sq*beta2)               (mul_ then addcmul_)
//   denom      = sqrt(exp_avg_sq) / sqrt(1-beta2^t) + eps
//   param      = fma(-lr/(1-beta1^t), exp_avg / denom, param)              (addcdiv_)
// The per-group scalars are formed on the host in double exactly as torch forms them in Python floats and rounded to
// float once, as torch's kernels do when they receive a Python scalar.
#include <cmath>
#include <cstdlib>
#include "common.cuh"

namespace gigs {

constexpr int ADAM_MAX_GROUPS = 24;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC_DEFAULT = 1;    // float4 per array per thread. Measured on B200 at 300k Gaussians (20 M elements,
                                       // all gradients read): 1 -> 0.124 ms (5.5 TB/s), 2 -> 0.130, 4 -> 0.180 (fewer, fatter CTAs
                                       // load and store in lock-step phases). GIGS_ADAM_VEC=1|2|4 overrides, for experiments

struct AdamGroupDev {
    float* param;
    float* grad;   // NULL: gradient known to be zero
    float* m;
    float* v;
    unsigned long long count;
    unsigned int first_block;   // first CTA of this group
    float w1;          // 1 - beta1
    float beta2;
    float w2;          // 1 - beta2
    float neg_step_size;   // -(lr / (1 - beta1^t))
    float bc2_sqrt;        // sqrt(1 - beta2^t)
    float bc2_rcp;         // rn(1 / bc2_sqrt): division by the per-group constant as a correctly rounded 3-FMA sequence
    float eps;
    int flags;             // bit 0: clamp param at >= 0 after the update; bit 1: clear grad; bit 2: pointers 16-B aligned
};

struct AdamArgs {
    int n_groups;
    AdamGroupDev g[ADAM_MAX_GROUPS];
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamGroupDev& G)
{
    m = __fmaf_rn(G.w1, __fsub_rn(g, m), m);
    v = __fmaf_rn(__fmul_rn(G.w2, g), g, __fmul_rn(v, G.beta2));
    // sqrt(v) / bc2_sqrt with the divisor's reciprocal precomputed (Markstein: q = s*R, r = s - q*b exactly, q' = q + r*R
    // is the correctly rounded quotient for a correctly rounded R; 3 instructions instead of the ~9 of an IEEE division —
    // this kernel is issue bound: 84 % of issue slots busy at 5.5 TB/s)
    const float sv = __fsqrt_rn(v);
    const float q0 = __fmul_rn(sv, G.bc2_rcp);
    float hat = __fmaf_rn(__fmaf_rn(-q0, G.bc2_sqrt, sv), G.bc2_rcp, q0);
    if (!(q0 < 3.0e38f)) hat = q0;      // overflowed second moment: inf / b = inf (the residual would be inf - inf)
    const float denom = __fadd_rn(hat, G.eps);
    p = __fmaf_rn(G.neg_step_size, __fdiv_rn(m, denom), p);
    if (G.flags & 1) p = (p < 0.f) ? 0.f : p;   // clamp_(min=0); a NaN stays a NaN as in torch (fmaxf would drop it)
}

template <int ADAM_VEC_PER_THREAD>
__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const __grid_constant__ AdamArgs A)
{
    pdl_enter();
    constexpr int ADAM_CHUNK = ADAM_THREADS * ADAM_VEC_PER_THREAD * 4;   // floats per CTA
    // group of this CTA: the table is tiny and uniform over the CTA
    int gi = 0;
#pragma unroll 1
    for (int i = 1; i < A.n_groups; i++)
        if (blockIdx.x >= A.g[i].first_block) gi = i;
    const AdamGroupDev& G = A.g[gi];
    const unsigned long long base = (unsigned long long)(blockIdx.x - G.first_block) * ADAM_CHUNK;
    const bool has_g = G.grad != nullptr;
    const bool clear = (G.flags & 2) && has_g;

    if ((G.flags & 4) && base + ADAM_CHUNK <= G.count) {
        float4* __restrict__ P = reinterpret_cast<float4*>(G.param + base);
        float4* __restrict__ M = reinterpret_cast<float4*>(G.m + base);
        float4* __restrict__ V = reinterpret_cast<float4*>(G.v + base);
        float4* __restrict__ Gr = has_g ? reinterpret_cast<float4*>(G.grad + base) : nullptr;
        float4 p[ADAM_VEC_PER_THREAD], m[ADAM_VEC_PER_THREAD], v[ADAM_VEC_PER_THREAD], g[ADAM_VEC_PER_THREAD];
#pragma unroll
        for (int k = 0; k < ADAM_VEC_PER_THREAD; k++) {          // all loads first: 16 x 16 B in flight per thread
            const int i = k * ADAM_THREADS + threadIdx.x;
            p[k] = P[i];
            m[k] = M[i];
            v[k] = V[i];
            g[k] = has_g ? Gr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < ADAM_VEC_PER_THREAD; k++) {
            const int i = k * ADAM_THREADS + threadIdx.x;
            float4 q = p[k];
            adam_elem(q.x, g[k].x, m[k].x, v[k].x, G);
            adam_elem(q.y, g[k].y, m[k].y, v[k].y, G);
            adam_elem(q.z, g[k].z, m[k].z, v[k].z, G);
            adam_elem(q.w, g[k].w, m[k].w, v[k].w, G);
            P[i] = q;
            M[i] = m[k];
            V[i] = v[k];
            if (clear) Gr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    // ragged tail of a group, or unaligned pointers: scalar
    for (unsigned long long i = base + threadIdx.x; i < base + ADAM_CHUNK && i < G.count; i += ADAM_THREADS) {
        float p = G.param[i], m = G.m[i], v = G.v[i];
        const float g = has_g ? G.grad[i] : 0.f;
        adam_elem(p, g, m, v, G);
        G.param[i] = p;
        G.m[i] = m;
        G.v[i] = v;
        if (clear) G.grad[i] = 0.f;
    }
}

// add_densification_stats + the max_radii2D update of /root/reference/train.py:489-495,
// scene/gaussian_model.py:933-945, for every Gaussian with radii > 0 (the visibility filter):
//   max_radii2D = max(max_radii2D, radii); xyz_gradient_accum += ||grad2D.xy||;
//   xyz_gradient_accum_abs += |gx| + |gy| (the reference's norm over a single column is an absolute value);
//   xyz_gradient_accum_abs_max = max(., |gx| + |gy|); denom += 1.
__global__ void densify_stats_kernel(int P, const int* __restrict__ radii, const float* __restrict__ grad2D, int gstride,
                                     float* __restrict__ accum, float* __restrict__ accum_abs,
                                     float* __restrict__ accum_abs_max, float* __restrict__ denom,
                                     float* __restrict__ max_radii2D)
{
    pdl_enter();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    if (max_radii2D) max_radii2D[i] = fmaxf(max_radii2D[i], (float)r);
    const float gx = grad2D[(size_t)i * gstride], gy = grad2D[(size_t)i * gstride + 1];
    accum[i] = __fadd_rn(accum[i], __fsqrt_rn(__fmaf_rn(gy, gy, __fmul_rn(gx, gx))));
    const float s = __fadd_rn(fabsf(gx), fabsf(gy));
    if (accum_abs) accum_abs[i] = __fadd_rn(accum_abs[i], s);
    if (accum_abs_max) accum_abs_max[i] = fmaxf(accum_abs_max[i], s);
    denom[i] = __fadd_rn(denom[i], 1.f);
}

}  // namespace gigs

using namespace gigs;

// ---- zero_grad of selected spans of the flat gradient buffer, one launch (optimizer.zero_grad(), train.py:518,522) ----
constexpr int CLEAR_MAX_SPANS = 8;
struct ClearArgs {
    float* base;
    unsigned long long begin[CLEAR_MAX_SPANS], end[CLEAR_MAX_SPANS];
    int n;
};
__global__ void __launch_bounds__(256) clear_spans_kernel(const ClearArgs a)
{
    pdl_enter();
    const size_t gtid = (size_t)blockIdx.x * 256 + threadIdx.x, gstride = (size_t)gridDim.x * 256;
    for (int s = 0; s < a.n; ++s) {
        float* p = a.base + a.begin[s];
        const size_t len = a.end[s] - a.begin[s];
        size_t head = ((16 - ((uintptr_t)p & 15)) & 15) / 4;   // floats up to the next 16-byte boundary
        if (head > len) head = len;
        if (gtid < head) p[gtid] = 0.f;
        float4* q = reinterpret_cast<float4*>(p + head);
        const size_t n4 = (len - head) / 4;
        for (size_t i = gtid; i < n4; i += gstride) q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t done = head + 4 * n4;
        if (gtid < len - done) p[done + gtid] = 0.f;
    }
}

extern "C" {

int gigs_adam_step(int32_t n_groups, const GigsAdamGroup* groups, void* stream)
{
    if (n_groups < 0 || n_groups > ADAM_MAX_GROUPS || (n_groups && !groups)) {
        set_error("gigs_adam_step: bad arguments (n_groups must be 0..%d)", ADAM_MAX_GROUPS);
        return -1;
    }
    AdamArgs A;
    A.n_groups = 0;
    unsigned long long blocks = 0;
    static int vec = 0;
    if (!vec) {
        const char* e = getenv("GIGS_ADAM_VEC");
        vec = e ? atoi(e) : ADAM_VEC_DEFAULT;
        if (vec != 1 && vec != 2 && vec != 4) vec = ADAM_VEC_DEFAULT;
    }
    const unsigned long long ADAM_CHUNK = (unsigned long long)ADAM_THREADS * vec * 4;
    for (int i = 0; i < n_groups; i++) {
        const GigsAdamGroup& s = groups[i];
        if (s.count == 0) continue;
        if (!s.param || !s.exp_avg || !s.exp_avg_sq || s.step < 1) {
            set_error("gigs_adam_step: group %d has a NULL param / state pointer or step < 1", i);
            return -1;
        }
        AdamGroupDev& d = A.g[A.n_groups++];
        d.param = s.param; d.grad = s.grad; d.m = s.exp_avg; d.v = s.exp_avg_sq;
        d.count = s.count;
        d.first_block = (unsigned int)blocks;
        // torch/optim/adam.py (_single_tensor_adam / _multi_tensor_adam): Python-float (double) scalars
        const double b1 = s.beta1, b2 = s.beta2;
        const double bc1 = 1.0 - std::pow(b1, (double)s.step);
        const double bc2 = 1.0 - std::pow(b2, (double)s.step);
        d.w1 = (float)(1.0 - b1);
        d.beta2 = (float)b2;
        d.w2 = (float)(1.0 - b2);
        d.neg_step_size = (float)(-(s.lr / bc1));
        d.bc2_sqrt = (float)std::sqrt(bc2);
        d.bc2_rcp = (float)(1.0 / (double)d.bc2_sqrt);
        d.eps = (float)s.eps;
        const uintptr_t al = (uintptr_t)s.param | (uintptr_t)s.exp_avg | (uintptr_t)s.exp_avg_sq | (uintptr_t)s.grad;
        d.flags = (s.clamp_min0 ? 1 : 0) | (s.clear_grad ? 2 : 0) | ((al & 15) == 0 ? 4 : 0);
        blocks += (s.count + ADAM_CHUNK - 1) / ADAM_CHUNK;
        if (blocks > 0x7fffffffull) { set_error("gigs_adam_step: too many elements"); return -1; }
    }
    if (!blocks) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope prof(26, st);
    if (vec == 1) GIGS_CUDA(launch_k(adam_kernel<1>, dim3((unsigned int)blocks), dim3(ADAM_THREADS), (size_t)(0), st, A));
    else if (vec == 2) GIGS_CUDA(launch_k(adam_kernel<2>, dim3((unsigned int)blocks), dim3(ADAM_THREADS), (size_t)(0), st, A));
    else GIGS_CUDA(launch_k(adam_kernel<4>, dim3((unsigned int)blocks), dim3(ADAM_THREADS), (size_t)(0), st, A));
    GIGS_LAUNCH_CHECK("adam_kernel");
    return 0;
}

int gigs_densify_stats(int32_t P, const int32_t* radii, const float* grad2D, int32_t grad_stride,
                       float* xyz_gradient_accum, float* xyz_gradient_accum_abs, float* xyz_gradient_accum_abs_max,
                       float* denom, float* max_radii2D, void* stream)
{
    if (P < 0 || grad_stride < 2) { set_error("gigs_densify_stats: bad arguments"); return -1; }
    if (P == 0) return 0;
    if (!radii || !grad2D || !xyz_gradient_accum || !denom) { set_error("gigs_densify_stats: NULL input"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    GIGS_CUDA(launch_k(densify_stats_kernel, dim3((P + 255) / 256), dim3(256), (size_t)(0), st, P, radii, grad2D, grad_stride, xyz_gradient_accum,
                                                          xyz_gradient_accum_abs, xyz_gradient_accum_abs_max, denom,
                                                          max_radii2D));
    GIGS_LAUNCH_CHECK("densify_stats_kernel");
    return 0;
}

int gigs_clear_spans(float* base, int32_t n_spans, const uint64_t* begin, const uint64_t* end, void* stream)
{
    if (n_spans < 0 || n_spans > CLEAR_MAX_SPANS) { set_error("gigs_clear_spans: at most %d spans", CLEAR_MAX_SPANS); return -1; }
    if (n_spans == 0) return 0;
    if (!base || !begin || !end) { set_error("gigs_clear_spans: NULL argument"); return -1; }
    ClearArgs a;
    a.base = base;
    a.n = n_spans;
    size_t total = 0;
    for (int s = 0; s < n_spans; ++s) {
        if (end[s] < begin[s]) { set_error("gigs_clear_spans: span %d ends before it begins", s); return -1; }
        a.begin[s] = begin[s];
        a.end[s] = end[s];
        total += (size_t)(end[s] - begin[s]);
    }
    if (total == 0) return 0;
    size_t blocks = (total / 4 + 256 * 4 - 1) / (256 * 4);   // ~4 x 16 bytes per thread
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaStream_t st = (cudaStream_t)stream;
    GIGS_CUDA(launch_k(clear_spans_kernel, dim3((unsigned)blocks), dim3(256), (size_t)0, st, a));
    GIGS_LAUNCH_CHECK("clear_spans_kernel");
    return 0;
}

}  // extern "C"
