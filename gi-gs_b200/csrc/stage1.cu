// Fused FIRST-STAGE frame: one view of /root/reference/train.py:266-328 (iteration <= pbr_iteration) as two C-ABI calls.
//
//   forward : getters (inside preprocess_kernel<RAW>) -> rasterize (full G-buffer incl. SH radiance) -> depth -> normal
//             chain (geometry_chain_kernel) -> render()'s normal post-processing (normalise, 3x3 median, rotation into view
//             space, gaussian_renderer/__init__.py:157-190) -> image loss (L1 + SSIM) and normal loss (masked L1 + TV),
//             both writing the gradient of their input map (csrc/loss.cu)
//   backward: normal post-processing backward (rotation, median as a gather over the stored selections, normalise) ->
//             general blend backward (dL/dcolour and dL/dnormal maps) -> per-Gaussian backward chained through the
//             getters straight into the leaves' gradient tensors (gaussian_backward_kernel<RAW>).
// The operator path (gigs.step.first_stage_step(fused=False): GaussianRasterizer + framework ops + autograd, ~200
// launches, 2.6 ms at 300k Gaussians / 800x800) computes the same thing and is what the parity tests compare against.
// SSAO is not run: the first-stage loss reads neither the occlusion nor any material map.
#include <cstring>
#include "common.cuh"
#include "filters.cuh"

namespace gigs {

int launch_preprocess(const GigsRasterFwd* a, const Layout& L, cudaStream_t st, const float* sh_rest);
int forward_finish_impl(GigsRasterFwd* a, bool lite);
int read_back_num_rendered(GigsRasterFwd* a, const Layout& L, cudaStream_t st);
int launch_blend_backward(const GigsRasterBwd* a, const Layout& L, cudaStream_t st);


constexpr int S1_TW = 32, S1_TH = 8, S1_HW = S1_TW + 2, S1_HH = S1_TH + 2;
constexpr int ST_S1_NORMALS = 29, ST_S1_NORMALS_BWD = 30;

__device__ __forceinline__ float3 s1_normalize_where_positive(float3 v)
{
    const float n = torch_norm_outer3(v.x, v.y, v.z);
    if (n > 0.f) {
        const float d = fmaxf(n, 1e-12f);
        return make_float3(v.x / d, v.y / d, v.z / d);
    }
    return v;
}

// render() post-processing of the two normal maps (gaussian_renderer/__init__.py:157-190):
//   normal_from_depth_mask = (nfd != 0).all(0); nfd = where(|nfd| > 0, normalize(nfd), nfd)
//   normal_map = where(|n| > 0, normalize(n), n); normal_map = median_blur3x3(normal_map); normals_view = -(normal_map^T R)
__global__ void __launch_bounds__(S1_TW* S1_TH)
stage1_normals_forward_kernel(const int W, const int H, const float* __restrict__ viewmatrix,
                              const float* __restrict__ normal_map, const float* __restrict__ nfd,
                              float* __restrict__ normals_view, float* __restrict__ nfd_unit, uint8_t* __restrict__ sel,
                              uint8_t* __restrict__ mask)
{
    pdl_enter();
    __shared__ float s_n[3][S1_HH][S1_HW];
    __shared__ float s_R[9];
    const int tid = threadIdx.y * S1_TW + threadIdx.x;
    const size_t HW = (size_t)W * H;
    const int x0 = blockIdx.x * S1_TW, y0 = blockIdx.y * S1_TH;
    if (tid < 9) s_R[tid] = viewmatrix[4 * (tid / 3) + (tid % 3)];   // world_view_transform[:3,:3][i][j]
    for (int i = tid; i < S1_HH * S1_HW; i += S1_TW * S1_TH) {
        const int lx = i % S1_HW, ly = i / S1_HW;
        const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
        float3 n = make_float3(0.f, 0.f, 0.f);     // median_blur pads with zeros
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t id = (size_t)gy * W + gx;
            n = s1_normalize_where_positive(make_float3(normal_map[id], normal_map[HW + id], normal_map[2 * HW + id]));
        }
        s_n[0][ly][lx] = n.x; s_n[1][ly][lx] = n.y; s_n[2][ly][lx] = n.z;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t id = (size_t)y * W + x;
    float mn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float a[9], b[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) a[dy * 3 + dx] = b[dy * 3 + dx] = s_n[c][threadIdx.y + dy][threadIdx.x + dx];
        mn[c] = median9(a);
        const int k = (mn[c] == mn[c]) ? median9_select(b, mn[c]) : -1;
        sel[c * HW + id] = (uint8_t)(k < 0 ? 255 : k);
    }
    normals_view[id] = -(mn[0] * s_R[0] + mn[1] * s_R[3] + mn[2] * s_R[6]);
    normals_view[HW + id] = -(mn[0] * s_R[1] + mn[1] * s_R[4] + mn[2] * s_R[7]);
    normals_view[2 * HW + id] = -(mn[0] * s_R[2] + mn[1] * s_R[5] + mn[2] * s_R[8]);
    const float3 d = make_float3(nfd[id], nfd[HW + id], nfd[2 * HW + id]);
    mask[id] = (d.x != 0.f && d.y != 0.f && d.z != 0.f) ? 1 : 0;
    const float3 du = s1_normalize_where_positive(d);
    nfd_unit[id] = du.x; nfd_unit[HW + id] = du.y; nfd_unit[2 * HW + id] = du.z;
}

// dL/dnormals_view -> dL/d(rasterizer normal map): rotation^T, median backward in gather form (pixel p collects the
// gradient of every neighbour whose stored selection points at p), normalise backward.
__global__ void __launch_bounds__(S1_TW* S1_TH)
stage1_normals_backward_kernel(const int W, const int H, const float* __restrict__ viewmatrix,
                               const float* __restrict__ normal_map, const float* __restrict__ g_view,
                               const uint8_t* __restrict__ sel, float* __restrict__ g_normal)
{
    pdl_enter();
    __shared__ float s_g[3][S1_HH][S1_HW];      // gradient w.r.t. the median's output (world frame)
    __shared__ uint8_t s_sel[3][S1_HH][S1_HW];
    __shared__ float s_R[9];
    const int tid = threadIdx.y * S1_TW + threadIdx.x;
    const size_t HW = (size_t)W * H;
    const int x0 = blockIdx.x * S1_TW, y0 = blockIdx.y * S1_TH;
    if (tid < 9) s_R[tid] = viewmatrix[4 * (tid / 3) + (tid % 3)];
    __syncthreads();
    for (int i = tid; i < S1_HH * S1_HW; i += S1_TW * S1_TH) {
        const int lx = i % S1_HW, ly = i / S1_HW;
        const int gx = x0 - 1 + lx, gy = y0 - 1 + ly;
        float g0 = 0.f, g1 = 0.f, g2 = 0.f;
        uint8_t k0 = 255, k1 = 255, k2 = 255;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t id = (size_t)gy * W + gx;
            const float a = g_view[id], b = g_view[HW + id], c = g_view[2 * HW + id];
            // normals_view_j = -sum_i m_i R[i][j]  =>  dL/dm_i = -sum_j R[i][j] g_j
            g0 = -(s_R[0] * a + s_R[1] * b + s_R[2] * c);
            g1 = -(s_R[3] * a + s_R[4] * b + s_R[5] * c);
            g2 = -(s_R[6] * a + s_R[7] * b + s_R[8] * c);
            k0 = sel[id]; k1 = sel[HW + id]; k2 = sel[2 * HW + id];
        }
        s_g[0][ly][lx] = g0; s_g[1][ly][lx] = g1; s_g[2][ly][lx] = g2;
        s_sel[0][ly][lx] = k0; s_sel[1][ly][lx] = k1; s_sel[2][ly][lx] = k2;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    const size_t id = (size_t)y * W + x;
    float g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float a = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                // neighbour q = p + (dx, dy); its window element k sits at q + (k%3 - 1, k/3 - 1) == p  <=>
                const int k = (1 - dy) * 3 + (1 - dx);
                if (s_sel[c][threadIdx.y + 1 + dy][threadIdx.x + 1 + dx] == k) a += s_g[c][threadIdx.y + 1 + dy][threadIdx.x + 1 + dx];
            }
        g[c] = a;
    }
    // where(|v| > 0, v / max(|v|, 1e-12), v) backward
    const float v[3] = {normal_map[id], normal_map[HW + id], normal_map[2 * HW + id]};
    const float nn = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    float o[3] = {g[0], g[1], g[2]};
    if (nn > 1e-12f) {
        const float u[3] = {v[0] / nn, v[1] / nn, v[2] / nn};
        const float d = u[0] * g[0] + u[1] * g[1] + u[2] * g[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = (g[c] - u[c] * d) / nn;
    } else if (nn > 0.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = g[c] / 1e-12f;
    }
    g_normal[id] = o[0]; g_normal[HW + id] = o[1]; g_normal[2 * HW + id] = o[2];
}

static GigsStage1Layout stage1_layout(int W, int H)
{
    GigsStage1Layout L;
    memset(&L, 0, sizeof(L));
    const uint64_t N = (uint64_t)W * H;
    uint64_t o = 0;
    auto take = [&](uint64_t bytes) { const uint64_t at = o; o = align_up(o + bytes, 256); return at; };
    L.color = take(3 * N * 4); L.opacity = take(N * 4); L.depth = take(N * 4); L.normal = take(3 * N * 4);
    L.normal_view = take(3 * N * 4); L.pos = take(3 * N * 4); L.albedo = take(3 * N * 4); L.roughness = take(N * 4);
    L.metallic = take(N * 4); L.normal_from_depth = take(3 * N * 4); L.depth_pos = take(3 * N * 4);
    L.normals_view = take(3 * N * 4); L.nfd_unit = take(3 * N * 4); L.g_color = take(3 * N * 4);
    L.g_normals_view = take(3 * N * 4); L.g_normal = take(3 * N * 4); L.median_sel = take(3 * N); L.mask = take(N);
    uint64_t need = 0;
    gigs_image_loss(3, W, H, nullptr, nullptr, 0.f, 1.f, nullptr, &need, nullptr, 0, nullptr, 0, nullptr, nullptr);
    L.loss_scratch = take(need);
    L.loss_scratch_bytes = need;
    gigs_normal_loss(W, H, nullptr, nullptr, nullptr, nullptr, 1.f, 1.f, 1.f, nullptr, &need, nullptr, 0, nullptr, 0, nullptr,
                     nullptr);
    L.nloss_scratch = take(need);
    L.nloss_scratch_bytes = need;
    L.stats = take(64);
    L.total_bytes = o;
    return L;
}

static int stage1_check(const GigsStage1* f)
{
    if (!f) { set_error("stage1: null args"); return -1; }
    if (f->P <= 0) { set_error("stage1: P must be positive"); return -1; }
    const GigsCamera& c = f->cam;
    if (c.width <= 1 || c.height <= 1 || !c.viewmatrix || !c.projmatrix || !c.campos || !c.bg) { set_error("stage1: bad camera"); return -1; }
    if (!f->xyz || !f->f_dc || !f->opacity || !f->normal || !f->albedo || !f->roughness || !f->metallic || !f->log_scale ||
        !f->rot || (c.sh_coeffs > 1 && !f->f_rest)) { set_error("stage1: a parameter pointer is NULL"); return -1; }
    if (!f->geom || !f->img || !f->maps || !f->radii) { set_error("stage1: workspace pointer is NULL"); return -1; }
    const GigsStage1Layout FL = stage1_layout(c.width, c.height);
    if (f->maps_bytes < FL.total_bytes) { set_error("stage1: maps blob too small (%llu < %llu)", (unsigned long long)f->maps_bytes, (unsigned long long)FL.total_bytes); return -2; }
    return 0;
}

}  // namespace gigs

using namespace gigs;

extern "C" {

int gigs_stage1_layout(int32_t W, int32_t H, GigsStage1Layout* out)
{
    if (!out || W <= 0 || H <= 0) { set_error("gigs_stage1_layout: bad arguments"); return -1; }
    *out = stage1_layout(W, H);
    return 0;
}

int gigs_stage1_forward(GigsStage1* f)
{
    if (int e = stage1_check(f)) return e;
    const GigsCamera& c = f->cam;
    cudaStream_t st = (cudaStream_t)f->stream;
    const GigsStage1Layout FL = stage1_layout(c.width, c.height);
    char* m = (char*)f->maps;
    GigsRasterFwd a;
    memset(&a, 0, sizeof(a));
    a.P = f->P;
    a.material_only = 2;   // radiance + normal + depth + opacity: the first-stage loss reads nothing else
    a.cam = c;
    a.cam.prefiltered = 0; a.cam.argmax_depth = 0; a.cam.inference = 0;
    a.means3D = f->xyz; a.shs = f->f_dc; a.opacities = f->opacity; a.normal = f->normal; a.albedo = f->albedo;
    a.roughness = f->roughness; a.metallic = f->metallic; a.scales = f->log_scale; a.rotations = f->rot;
    a.out_color = (float*)(m + FL.color); a.out_opacity = (float*)(m + FL.opacity); a.out_depth = (float*)(m + FL.depth);
    a.out_normal = (float*)(m + FL.normal); a.out_normal_view = (float*)(m + FL.normal_view);
    a.out_pos = (float*)(m + FL.pos); a.out_albedo = (float*)(m + FL.albedo);
    a.out_roughness = (float*)(m + FL.roughness); a.out_metallic = (float*)(m + FL.metallic);
    a.radii = f->radii;
    a.geom = f->geom; a.geom_bytes = f->geom_bytes; a.img = f->img; a.img_bytes = f->img_bytes;
    a.binning = f->binning; a.binning_bytes = f->binning_bytes; a.sort = f->sort; a.sort_bytes = f->sort_bytes;
    a.pinned_num_rendered = (uint32_t*)f->pinned_num_rendered;
    a.num_rendered = f->num_rendered;
    a.stream = f->stream;
    if (!f->resume) {
        const Layout L0 = make_layout(f->P, c.width, c.height, 0);
        if (f->geom_bytes < L0.size.geom_bytes || f->img_bytes < L0.size.img_bytes) { set_error("stage1: geom/img workspace too small"); return -2; }
        {
            ProfScope ps(ST_PREPROCESS, st);
            if (int e = launch_preprocess(&a, L0, st, f->f_rest ? f->f_rest : f->f_dc)) return e;
        }
        if (int e = read_back_num_rendered(&a, L0, st)) return e;
        f->num_rendered = a.num_rendered;
    }
    a.num_rendered = f->num_rendered;
    const Layout L = make_layout(f->P, c.width, c.height, (uint64_t)f->num_rendered);
    f->need_binning_bytes = L.size.binning_bytes;
    f->need_sort_bytes = L.size.sort_bytes;
    if (!f->binning || !f->sort || f->binning_bytes < L.size.binning_bytes || f->sort_bytes < L.size.sort_bytes) {
        set_error("stage1: binning/sort workspace too small (need %llu / %llu bytes)", (unsigned long long)L.size.binning_bytes,
                  (unsigned long long)L.size.sort_bytes);
        return -5;   // GIGS_E_GROW
    }
    if (int e = forward_finish_impl(&a, false)) return e;

    const int W = c.width, H = c.height;
    const float fx = W / (2.0f * c.tan_fovx), fy = H / (2.0f * c.tan_fovy);
    if (int e = gigs_geometry_chain(W, H, fx, fy, c.viewmatrix, (float*)(m + FL.depth), 1, (float*)(m + FL.normal_from_depth),
                                    (float*)(m + FL.depth_pos), f->stream)) return e;
    dim3 grid((W + S1_TW - 1) / S1_TW, (H + S1_TH - 1) / S1_TH), block(S1_TW, S1_TH);
    {
        ProfScope ps(ST_S1_NORMALS, st);
        GIGS_CUDA(launch_k(stage1_normals_forward_kernel, dim3(grid), dim3(block), (size_t)(0), st, W, H, c.viewmatrix, (float*)(m + FL.normal),
                                                              (float*)(m + FL.normal_from_depth), (float*)(m + FL.normals_view),
                                                              (float*)(m + FL.nfd_unit), (uint8_t*)(m + FL.median_sel),
                                                              (uint8_t*)(m + FL.mask)));
        GIGS_LAUNCH_CHECK("stage1_normals_forward_kernel");
    }
    if (f->gt_image) {
        if (f->gt_ready_event) GIGS_CUDA(cudaStreamWaitEvent(st, (cudaEvent_t)f->gt_ready_event, 0));
        float* stats = (float*)(m + FL.stats);
        uint64_t nb = FL.loss_scratch_bytes;
        if (int e = gigs_image_loss(3, W, H, (float*)(m + FL.color), f->gt_image, f->lambda_dssim, f->loss_scale,
                                    m + FL.loss_scratch, &nb, stats, 0, (float*)(m + FL.g_color), 0, nullptr, f->stream))
            return e;
        nb = FL.nloss_scratch_bytes;
        if (int e = gigs_normal_loss(W, H, (float*)(m + FL.normals_view), (float*)(m + FL.nfd_unit), (uint8_t*)(m + FL.mask),
                                     f->gt_image, f->normal_weight, f->normal_tv_weight, f->loss_scale, m + FL.nloss_scratch,
                                     &nb, stats + 4, 0, (float*)(m + FL.g_normals_view), 0, nullptr, f->stream))
            return e;
    }
    if (c.debug) GIGS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int gigs_stage1_backward(GigsStage1* f)
{
    if (int e = stage1_check(f)) return e;
    if (!f->gt_image) { set_error("stage1_backward: the forward ran without a ground-truth image"); return -1; }
    if (!f->accum || !f->binning || !f->g_xyz || !f->g_f_dc || !f->g_opacity || !f->g_normal || !f->g_albedo ||
        !f->g_roughness || !f->g_metallic || !f->g_log_scale || !f->g_rot || (f->cam.sh_coeffs > 1 && !f->g_f_rest)) {
        set_error("stage1_backward: a gradient / accum / binning pointer is NULL");
        return -1;
    }
    const GigsCamera& c = f->cam;
    cudaStream_t st = (cudaStream_t)f->stream;
    const GigsStage1Layout FL = stage1_layout(c.width, c.height);
    const int W = c.width, H = c.height;
    char* m = (char*)f->maps;
    dim3 grid((W + S1_TW - 1) / S1_TW, (H + S1_TH - 1) / S1_TH), block(S1_TW, S1_TH);
    {
        ProfScope ps(ST_S1_NORMALS_BWD, st);
        GIGS_CUDA(launch_k(stage1_normals_backward_kernel, dim3(grid), dim3(block), (size_t)(0), st, W, H, c.viewmatrix, (float*)(m + FL.normal),
                                                               (float*)(m + FL.g_normals_view), (uint8_t*)(m + FL.median_sel),
                                                               (float*)(m + FL.g_normal)));
        GIGS_LAUNCH_CHECK("stage1_normals_backward_kernel");
    }
    GigsRasterBwd b;
    memset(&b, 0, sizeof(b));
    b.P = f->P; b.num_rendered = f->num_rendered; b.cam = c;
    b.geom = f->geom; b.binning = f->binning; b.img = f->img;
    b.dL_dpix = (float*)(m + FL.g_color);
    b.dL_dpix_normal = (float*)(m + FL.g_normal);
    b.accum = f->accum;
    const Layout L = make_layout(f->P, W, H, (uint64_t)f->num_rendered);
    {
        ProfScope ps(ST_BLEND_BWD, st);
        GIGS_CUDA(cudaMemsetAsync(f->accum, 0, (size_t)f->P * ACC_FLOATS * sizeof(float), st));
        if (f->num_rendered > 0)
            if (int e = launch_blend_backward(&b, L, st)) return e;
    }
    {
        ProfScope ps(ST_GAUSS_BWD, st);
        RawGrads r;
        r.f_rest = f->f_rest; r.opacity = f->opacity; r.normal = f->normal; r.albedo = f->albedo;
        r.roughness = f->roughness; r.metallic = f->metallic;
        r.g_xyz = f->g_xyz; r.g_f_dc = f->g_f_dc; r.g_f_rest = f->g_f_rest; r.g_opacity = f->g_opacity;
        r.g_normal = f->g_normal; r.g_albedo = f->g_albedo; r.g_roughness = f->g_roughness; r.g_metallic = f->g_metallic;
        r.g_log_scale = f->g_log_scale; r.g_rot = f->g_rot;
        if (int e = launch_gaussian_backward_raw(f->P, c, f->geom, L, f->radii, f->accum, f->xyz, f->f_dc, f->log_scale,
                                                 f->rot, f->g_means2D, r, st))
            return e;
    }
    if (c.debug) GIGS_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
