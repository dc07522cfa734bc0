// Densification / pruning rebuild: GaussianModel.densify_and_prune of
// /root/reference/scene/gaussian_model.py:905-931 (densify_and_clone :785-817, densify_and_split :741-783,
// densification_postfix / cat_tensors_to_optimizer :639-739, prune_points / _prune_optimizer :594-637).
//
// The reference edits the model in four rounds — concatenate the clones onto all 10 parameter tensors and both Adam
// moments, concatenate the split children, mask out the split parents, mask out the pruned points — i.e. ~120 tensor
// copies of the whole model. The outcome of the four rounds is a pure function of three per-point decisions, so the
// host computes ONE source map for the surviving rows (gigs/densify.py) and this kernel materialises all 10 parameter
// tensors and their 20 moment tensors in one launch:
//   kind 0 (kept point)   : parameter row and both moment rows copied from row src;
//   kind 1 (clone)        : parameters copied, moments zero; xyz = R(rot) * (noise * exp(log_scale)) + xyz   (:797-801)
//   kind 2 (split child)  : as a clone, and log_scale = log(exp(log_scale) / split_div), split_div = 0.8 * N (:757-760)
// `noise` holds one standard-normal 3-vector per output row (torch.normal(mean=0, std=s) is randn * s); rows of kept
// points are not read. Rotation matrix as utils/general_utils.py:89-110 (build_rotation), operation by operation.
#include "common.cuh"

namespace gigs {

constexpr int DG_MAX_GROUPS = 16;

struct DensifyGroupDev {
    const float* src; const float* src_m; const float* src_v;
    float* dst; float* dst_m; float* dst_v;
    int width, role;
    unsigned long long first_elem;   // prefix of n_out * width over the groups
};

struct DensifyArgs {
    int n_groups;
    int n_out;
    const int* src_index;
    const signed char* kind;
    const float* noise;
    const float* log_scale;
    const float* rot;
    float split_div;
    unsigned long long total;
    DensifyGroupDev g[DG_MAX_GROUPS];
};

__device__ __forceinline__ float3 sample_offset(const DensifyArgs& A, int row, int s)
{
    const float sx = expf(A.log_scale[(size_t)s * 3 + 0]), sy = expf(A.log_scale[(size_t)s * 3 + 1]),
                sz = expf(A.log_scale[(size_t)s * 3 + 2]);
    const float vx = __fmul_rn(A.noise[(size_t)row * 3 + 0], sx), vy = __fmul_rn(A.noise[(size_t)row * 3 + 1], sy),
                vz = __fmul_rn(A.noise[(size_t)row * 3 + 2], sz);
    const float q0 = A.rot[(size_t)s * 4 + 0], q1 = A.rot[(size_t)s * 4 + 1], q2 = A.rot[(size_t)s * 4 + 2],
                q3 = A.rot[(size_t)s * 4 + 3];
    const float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2)),
                                         __fmul_rn(q3, q3)));
    const float r = __fdiv_rn(q0, n), x = __fdiv_rn(q1, n), y = __fdiv_rn(q2, n), z = __fdiv_rn(q3, n);
#define M2(a, b) __fmul_rn(a, b)
    const float R00 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(M2(y, y), M2(z, z))));
    const float R01 = __fmul_rn(2.f, __fsub_rn(M2(x, y), M2(r, z)));
    const float R02 = __fmul_rn(2.f, __fadd_rn(M2(x, z), M2(r, y)));
    const float R10 = __fmul_rn(2.f, __fadd_rn(M2(x, y), M2(r, z)));
    const float R11 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(M2(x, x), M2(z, z))));
    const float R12 = __fmul_rn(2.f, __fsub_rn(M2(y, z), M2(r, x)));
    const float R20 = __fmul_rn(2.f, __fsub_rn(M2(x, z), M2(r, y)));
    const float R21 = __fmul_rn(2.f, __fadd_rn(M2(y, z), M2(r, x)));
    const float R22 = __fsub_rn(1.f, __fmul_rn(2.f, __fadd_rn(M2(x, x), M2(y, y))));
#undef M2
    // torch.bmm(rots, samples): a 3-term dot product per component (summation order of the library kernel unknown)
    return make_float3(fmaf(R02, vz, fmaf(R01, vy, R00 * vx)), fmaf(R12, vz, fmaf(R11, vy, R10 * vx)),
                       fmaf(R22, vz, fmaf(R21, vy, R20 * vx)));
}

__global__ void __launch_bounds__(256) densify_gather_kernel(const __grid_constant__ DensifyArgs A)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < A.total; e += stride) {
        int gi = 0;
#pragma unroll 1
        for (int i = 1; i < A.n_groups; i++)
            if (e >= A.g[i].first_elem) gi = i;
        const DensifyGroupDev& G = A.g[gi];
        const unsigned long long le = e - G.first_elem;
        const int row = (int)(le / (unsigned)G.width), col = (int)(le - (unsigned long long)row * G.width);
        const int s = A.src_index[row];
        const int k = A.kind[row];
        const size_t so = (size_t)s * G.width + col;
        float v = G.src[so];
        if (k != 0) {
            if (G.role == 1) {
                const float3 o = sample_offset(A, row, s);
                v = __fadd_rn(col == 0 ? o.x : (col == 1 ? o.y : o.z), v);
            } else if (G.role == 2 && k == 2) {
                v = logf(__fdiv_rn(expf(v), A.split_div));
            }
        }
        G.dst[le] = v;
        if (G.dst_m) G.dst_m[le] = (k == 0 && G.src_m) ? G.src_m[so] : 0.f;
        if (G.dst_v) G.dst_v[le] = (k == 0 && G.src_v) ? G.src_v[so] : 0.f;
    }
}

}  // namespace gigs

using namespace gigs;

extern "C" {

int gigs_densify_gather(int32_t n_out, const int32_t* src_index, const int8_t* kind, const float* noise,
                        const float* src_log_scale, const float* src_rot, float split_div, int32_t n_groups,
                        const GigsDensifyGroup* groups, void* stream)
{
    if (n_out < 0 || n_groups < 0 || n_groups > DG_MAX_GROUPS || (n_groups && !groups)) {
        set_error("gigs_densify_gather: bad arguments (n_groups must be 0..%d)", DG_MAX_GROUPS);
        return -1;
    }
    if (n_out == 0 || n_groups == 0) return 0;
    if (!src_index || !kind) { set_error("gigs_densify_gather: NULL source map"); return -1; }
    DensifyArgs A;
    A.n_groups = n_groups; A.n_out = n_out; A.src_index = src_index; A.kind = (const signed char*)kind;
    A.noise = noise; A.log_scale = src_log_scale; A.rot = src_rot; A.split_div = split_div;
    unsigned long long total = 0;
    for (int i = 0; i < n_groups; i++) {
        const GigsDensifyGroup& s = groups[i];
        if (!s.src || !s.dst || s.width <= 0 || s.role < 0 || s.role > 2 || (s.role == 1 && s.width != 3) ||
            (s.role == 2 && s.width != 3)) {
            set_error("gigs_densify_gather: group %d has a NULL tensor, a bad width or a bad role", i);
            return -1;
        }
        if (s.role != 0 && (!noise || !src_log_scale || !src_rot)) {
            set_error("gigs_densify_gather: xyz / scale groups need noise, src_log_scale and src_rot");
            return -1;
        }
        DensifyGroupDev& d = A.g[i];
        d.src = s.src; d.src_m = s.src_exp_avg; d.src_v = s.src_exp_avg_sq;
        d.dst = s.dst; d.dst_m = s.dst_exp_avg; d.dst_v = s.dst_exp_avg_sq;
        d.width = s.width; d.role = s.role; d.first_elem = total;
        total += (unsigned long long)n_out * s.width;
    }
    A.total = total;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned long long want = (total + 255) / 256;
    const unsigned int blocks = (unsigned int)(want < 148ull * 32 ? want : 148ull * 32);   // grid-stride over the rest
    densify_gather_kernel<<<blocks, 256, 0, st>>>(A);
    GIGS_LAUNCH_CHECK("densify_gather_kernel");
    return 0;
}

}  // extern "C"
