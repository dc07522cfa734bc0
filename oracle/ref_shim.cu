// TEST INFRASTRUCTURE — not product code.
//
// C-ABI shim around the *unmodified* reference CUDA rasterizer so that tests and
// bench.py can run the reference's own kernels on the GPU box next to ours.
// The reference sources are compiled where they lie under /root/reference by
// oracle/Makefile; nothing from the reference is copied into this repo.
//
// Wraps (reference file:line):
//   CudaRasterizer::Rasterizer::forward        cuda_rasterizer/rasterizer_impl.cu:486
//   CudaRasterizer::Rasterizer::backward       cuda_rasterizer/rasterizer_impl.cu:676
//   CudaRasterizer::Rasterizer::lite_forward   cuda_rasterizer/rasterizer_impl.cu:338
//   CudaRasterizer::Rasterizer::depthToNormal  cuda_rasterizer/rasterizer_impl.cu:200
//   CudaRasterizer::Rasterizer::SSAO           cuda_rasterizer/rasterizer_impl.cu:220
//   CudaRasterizer::Rasterizer::SSR            cuda_rasterizer/rasterizer_impl.cu:250
//   CudaRasterizer::Rasterizer::markVisible    cuda_rasterizer/rasterizer_impl.cu:141
//   SimpleKNN::knn                             simple-knn/simple_knn.cu:165
// The three workspaces the reference grows through std::function callbacks
// (rasterize_points.cu:31-37) are plain cudaMalloc'd blobs owned by a context.
#include <cstdint>
#include <cfloat>
#include <cstdio>
#include <functional>
#include <stdexcept>
#include <string>
#include <cuda_runtime.h>
#include "rasterizer_impl.h"   // reference: GeometryState/ImageState/BinningState::fromChunk
#include "simple_knn.h"        // reference: SimpleKNN::knn

using namespace CudaRasterizer;

struct RefBlob {
    char* ptr = nullptr;
    size_t cap = 0;
    char* ensure(size_t n) {
        if (n > cap) {
            if (ptr) cudaFree(ptr);
            cap = n + (n >> 2) + 1024;
            if (cudaMalloc(&ptr, cap) != cudaSuccess) { ptr = nullptr; cap = 0; }
        }
        return ptr;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

struct RefCtx {
    RefBlob geom, binning, img;
    int P = 0, R = 0, N = 0;
};

static thread_local std::string g_err;

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void* ref_ctx_create() { return new RefCtx(); }
void ref_ctx_destroy(void* c) {
    RefCtx* ctx = (RefCtx*)c;
    ctx->geom.release(); ctx->binning.release(); ctx->img.release();
    delete ctx;
}

// Returns num_rendered (>=0) or -1 on exception.
int ref_forward(void* c, int P, int D, int M, const float* background, int W, int H,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* normal, const float* albedo,
                const float* roughness, const float* metallic, const float* scales,
                float scale_modifier, const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                float tan_fovx, float tan_fovy, int prefiltered, int argmax_depth, int inference,
                float* out_color, float* out_opacity, float* out_depth, float* out_normal,
                float* out_normal_view, float* out_pos, float* out_albedo, float* out_roughness,
                float* out_metallic, int* radii, int debug)
{
    RefCtx* ctx = (RefCtx*)c;
    try {
        std::function<char*(size_t)> gf = [ctx](size_t n) { return ctx->geom.ensure(n); };
        std::function<char*(size_t)> bf = [ctx](size_t n) { return ctx->binning.ensure(n); };
        std::function<char*(size_t)> imf = [ctx](size_t n) { return ctx->img.ensure(n); };
        int R = Rasterizer::forward(gf, bf, imf, P, D, M, background, W, H, means3D, shs,
            colors_precomp, opacities, normal, albedo, roughness, metallic, scales, scale_modifier,
            rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy,
            prefiltered != 0, argmax_depth != 0, inference != 0, out_color, out_opacity, out_depth,
            out_normal, out_normal_view, out_pos, out_albedo, out_roughness, out_metallic, radii,
            debug != 0);
        ctx->P = P; ctx->R = R; ctx->N = W * H;
        return R;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// The radiance-only rasterizer (lite_rasterize_gaussians, rasterize_points.cu:40-128). Returns num_rendered or -1.
int ref_lite_forward(void* c, int P, int D, int M, const float* background, int W, int H,
                     const float* means3D, const float* shs, const float* colors_precomp,
                     const float* opacities, const float* scales, float scale_modifier,
                     const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                     const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy,
                     int prefiltered, int argmax_depth, float* out_color, float* out_opacity,
                     float* out_depth, int* radii)
{
    RefCtx* ctx = (RefCtx*)c;
    try {
        std::function<char*(size_t)> gf = [ctx](size_t n) { return ctx->geom.ensure(n); };
        std::function<char*(size_t)> bf = [ctx](size_t n) { return ctx->binning.ensure(n); };
        std::function<char*(size_t)> imf = [ctx](size_t n) { return ctx->img.ensure(n); };
        int R = Rasterizer::lite_forward(gf, bf, imf, P, D, M, background, W, H, means3D, shs,
            colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix,
            projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered != 0, argmax_depth != 0, out_color,
            out_opacity, out_depth, radii);
        ctx->P = P; ctx->R = R; ctx->N = W * H;
        return R;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

int ref_backward(void* c, int P, int D, int M, int R, const float* background, int W, int H,
                 const float* means3D, const float* shs, const float* colors_precomp,
                 const float* normal, const float* albedo, const float* roughness,
                 const float* metallic, const float* scales, const float* rotations,
                 const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                 const float* cam_pos, const int* radii, float scale_modifier, float tan_fovx,
                 float tan_fovy, const float* dL_dpix_depth, const float* dL_dpix,
                 const float* dL_dpix_opacity, const float* dL_dpix_normal,
                 const float* dL_dpix_albedo, const float* dL_dpix_roughness,
                 const float* dL_dpix_metallic, float* dL_dmean2D, float* dL_dconic,
                 float* dL_depth, float* dL_dopacity, float* dL_dnormal, float* dL_dalbedo,
                 float* dL_droughness, float* dL_dmetallic, float* dL_dcolor, float* dL_dmean3D,
                 float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot, int debug)
{
    RefCtx* ctx = (RefCtx*)c;
    try {
        Rasterizer::backward(P, D, M, R, background, W, H, means3D, shs, colors_precomp, normal,
            albedo, roughness, metallic, scales, rotations, cov3D_precomp, viewmatrix, projmatrix,
            cam_pos, radii, scale_modifier, tan_fovx, tan_fovy, ctx->geom.ptr, ctx->binning.ptr,
            ctx->img.ptr, dL_dpix_depth, dL_dpix, dL_dpix_opacity, dL_dpix_normal, dL_dpix_albedo,
            dL_dpix_roughness, dL_dpix_metallic, dL_dmean2D, dL_dconic, dL_depth, dL_dopacity,
            dL_dnormal, dL_dalbedo, dL_droughness, dL_dmetallic, dL_dcolor, dL_dmean3D, dL_dcov3D,
            dL_dsh, dL_dscale, dL_drot, debug != 0);
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// Decoded pointers into the reference workspaces of the last forward (device addresses).
// order: depths, pos_view, clamped, means2D, cov3D, conic_opacity, rgb, tiles_touched,
//        point_offsets, keys_unsorted, keys_sorted, vals_unsorted, vals_sorted(point_list),
//        accum_alpha, n_contrib, ranges
void ref_state_ptrs(void* c, void** out16)
{
    RefCtx* ctx = (RefCtx*)c;
    char* g = ctx->geom.ptr;
    GeometryState gs = GeometryState::fromChunk(g, ctx->P);
    char* b = ctx->binning.ptr;
    BinningState bs = BinningState::fromChunk(b, ctx->R);
    char* i = ctx->img.ptr;
    ImageState is = ImageState::fromChunk(i, ctx->N);
    out16[0] = gs.depths; out16[1] = gs.pos_view; out16[2] = gs.clamped; out16[3] = gs.means2D;
    out16[4] = gs.cov3D; out16[5] = gs.conic_opacity; out16[6] = gs.rgb; out16[7] = gs.tiles_touched;
    out16[8] = gs.point_offsets; out16[9] = bs.point_list_keys_unsorted; out16[10] = bs.point_list_keys;
    out16[11] = bs.point_list_unsorted; out16[12] = bs.point_list; out16[13] = is.accum_alpha;
    out16[14] = is.n_contrib; out16[15] = is.ranges;
}

void ref_depth_to_normal(int W, int H, float fx, float fy, const float* viewmatrix,
                         const float* depth, float* normal_map, float* depth_pos)
{
    Rasterizer::depthToNormal(W, H, fx, fy, viewmatrix, depth, normal_map, depth_pos);
}

void ref_ssao(int W, int H, float fx, float fy, float radius, float bias, float thick, float delta,
              int step, int start, const float* normal, const float* pos, float* occlusion)
{
    Rasterizer::SSAO(W, H, fx, fy, radius, bias, thick, delta, step, start, normal, pos, occlusion);
}

void ref_ssr(int W, int H, float fx, float fy, float radius, float bias, float thick, float delta,
             int step, int start, const float* normal, const float* pos, const float* rgb,
             const float* albedo, const float* roughness, const float* metallic, const float* F0,
             float* color, float* abd)
{
    Rasterizer::SSR(W, H, fx, fy, radius, bias, thick, delta, step, start, normal, pos, rgb, albedo,
                    roughness, metallic, F0, color, abd);
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
    Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, present);
}

void ref_knn(int P, float* points, float* mean_dists)
{
    SimpleKNN::knn(P, (float3*)points, mean_dists);
}

int ref_sync() { return (int)cudaDeviceSynchronize(); }

int ref_memcpy_d2d(void* dst, const void* src, size_t n) { return (int)cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice); }

}  // extern "C"
