// Shared device/host helpers for the gigs_b200 kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../include/gigs_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "gigs_b200 is written for sm_100a (B200) only"
#endif

namespace gigs {

constexpr int TILE_X = 16;       // reference config.h:16-17 (the tile size is part of the key contract)
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;
constexpr int REC_FLOATS = 24;   // packed blend record, 96 B, 16-B aligned (DESIGN.md "HBM layout")
constexpr int WARP_MASK_WORDS = 8;   // warps of a 16x16 tile
constexpr int ACC_FLOATS = 20;   // packed per-Gaussian gradient accumulator, 80 B

// record slots
enum : int {
    R_X = 0, R_Y = 1, R_CA = 2, R_CB = 3,           // float4 #0: mean2D.xy, conic.x, conic.y
    R_CC = 4, R_OP = 5, R_DEPTH = 6, R_TAU = 7,      // float4 #1: conic.z, opacity, depth, ln(255*opacity)
    R_RGB = 8, R_ROUGH = 11,                         // float4 #2
    R_ALB = 12, R_METAL = 15,                        // float4 #3
    R_NRM = 16, R_PX = 19,                           // float4 #4
    R_PY = 20, R_PZ = 21                             // float4 #5 (22,23 spare)
};
// accumulator slots
enum : int {
    A_COL = 0, A_DEPTH = 3,      // {color3, depth}
    A_NRM = 4, A_ROUGH = 7,      // {normal3, roughness}
    A_ALB = 8, A_METAL = 11,     // {albedo3, metallic}
    A_M2X = 12, A_M2Y = 13, A_M2Z = 14, A_OPAC = 15,  // {mean2D.xyz, opacity}
    A_CX = 16, A_CY = 17, A_CW = 18                   // {conic.x, .y, .w, -}
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of a function: set it once per (device,
// kernel) under a mutex (api.cu). Returns 0 or a cudaError_t (message in gigs_last_error()).
int ensure_dynamic_smem(const void* kernel, int bytes);
#define GIGS_SMEM_ATTR(kernel, bytes)                                                    \
    do {                                                                                 \
        if (int _e = gigs::ensure_dynamic_smem((const void*)(kernel), (int)(bytes))) return _e; \
    } while (0)

// num_rendered reaches the host through a page-locked word the preprocess scan kernel stores to directly (api.cu)
constexpr uint32_t NUM_RENDERED_PENDING = 0xffffffffu;
struct HostSlot { uint32_t* host; uint32_t* dev; };   // dev == nullptr: the word is not mapped, use a copy
int host_total_slot(const GigsRasterFwd* a, HostSlot* s);

// Programmatic dependent launch (PTX griddepcontrol): a kernel launched through launch_k may be scheduled while the
// previous kernel of the stream is still draining; it calls pdl_wait() FIRST (every thread, before any global access:
// the wait returns once the previous grid has completed and its writes are visible) and pdl_trigger() right after (lets
// the next launch_k kernel be scheduled). Both are no-ops in a kernel launched the ordinary way. Only kernels that
// start with pdl_wait() may be launched with launch_k: completion then stays transitive along the stream.
bool pdl_enabled();
void note_launch();   // counts the kernels launched through launch_k (gigs_launch_count)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    note_launch();
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define GIGS_CUDA(call)                                          \
    do {                                                         \
        cudaError_t _e = (call);                                 \
        if (_e != cudaSuccess) return gigs::cuda_fail(_e, #call); \
    } while (0)
#define GIGS_LAUNCH_CHECK(name)                                  \
    do {                                                         \
        cudaError_t _e = cudaGetLastError();                     \
        if (_e != cudaSuccess) return gigs::cuda_fail(_e, name); \
    } while (0)

// optional per-stage event timing (api.cu)
enum Stage : int {
    ST_PREPROCESS = 0, ST_EMIT_KEYS, ST_SORT, ST_RANGES, ST_BLEND_FWD, ST_BLEND_BWD, ST_GAUSS_BWD, ST_GEOM_CHAIN,
    ST_SSAO, ST_SSR, ST_SHADE_FWD, ST_SHADE_BWD, ST_MEDIAN, ST_MEDIAN_BWD, ST_BILATERAL, ST_D2N, ST_SSR_BWD, ST_DIST2,
    ST_DEFER_SHADE, ST_DEFER_LOSS, ST_DEFER_BWD, ST_PARAM_GRAD, ST_SORT_PASS, ST_DEPTH_SORT, ST_CUBEMAP, ST_CUBEMAP_BWD,
    ST_DEFER_BWD_KERNEL = 31, ST_PEER_ALLREDUCE = 32
};
int prof_begin(int stage, cudaStream_t st);   // returns a token (<0 when profiling is off)
void prof_end(int token, cudaStream_t st);
struct ProfScope {
    int tok;
    cudaStream_t st;
    ProfScope(int stage, cudaStream_t s) : tok(prof_begin(stage, s)), st(s) {}
    ~ProfScope() { prof_end(tok, st); }
};

__host__ __device__ inline uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Workspace layout: pure function of (P, W, H, R).
// ---------------------------------------------------------------------------------------------
struct Layout {
    GigsLayout off;
    GigsSizes size;
    uint32_t tiles_x, tiles_y, num_tiles, num_blocks;
    // instance sort (by tile id) scratch internals
    uint64_t s_keys_a, s_keys_b, s_vals_b, s_hist, s_joint, s_status, s_ticket, s_zero_bytes;
    uint32_t sort_tiles, sort_passes, sort_bits;
    // Gaussian depth argsort internals (geom blob) + block sums of tiles_touched in depth order
    uint64_t p_keys_a, p_keys_b, p_vals_b, p_hist, p_status, p_ticket, p_zero_bytes, g_block_sums2;
    // binning blob: the forward blend's per-(chunk of 32 list entries, warp) footprint-test results, reused by the
    // backward kernels (WARP_MASK_WORDS words per chunk; chunk index of tile t = (range.x >> 5) + t + chunk in tile)
    uint64_t b_warp_masks;
};
Layout make_layout(int P, int W, int H, uint64_t R);
uint32_t higher_msb(uint32_t n);
int radix_digit_bits(int bits);
int radix_sort_passes(int bits);
uint64_t radix_joint_bytes();

// ---------------------------------------------------------------------------------------------
// Pixel block of one warp inside the 16x16 tile. The blend kernels reject, once per warp, the Gaussians whose 1/255
// iso-ellipse misses the warp's block; a compact 8x4 block is missed ~15-20 % more often than the 16x2 strip a
// row-major thread layout gives (the Minkowski sum of block and footprint is smaller), so fewer (warp, Gaussian) pairs
// are walked. Per pixel the blend order is unchanged, so the forward results are bit-identical either way.
// ---------------------------------------------------------------------------------------------
#ifndef GIGS_WARP_COLS
#define GIGS_WARP_COLS 8
#endif
constexpr int WARP_COLS = GIGS_WARP_COLS, WARP_ROWS = 32 / WARP_COLS;
// tile-local pixel of thread `tid` (0..255): warps tile the 16x16 block in WARP_COLS x WARP_ROWS pieces
__device__ __forceinline__ void warp_block_pixel(int tid, int& lx, int& ly)
{
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int per_row = 16 / WARP_COLS;                 // warp blocks per tile row
    lx = (warp % per_row) * WARP_COLS + (lane % WARP_COLS);
    ly = (warp / per_row) * WARP_ROWS + (lane / WARP_COLS);
}
// tile-local origin of the warp's block
__device__ __forceinline__ void warp_block_origin(int tid, int& ox, int& oy)
{
    const int warp = tid >> 5;
    constexpr int per_row = 16 / WARP_COLS;
    ox = (warp % per_row) * WARP_COLS;
    oy = (warp / per_row) * WARP_ROWS;
}
// lower bound of the conic's quadratic form 0.5 d^T C d over the block, d = mean2D - pixel: d.x in [hx-(COLS-1), hx],
// d.y in {hy, hy-1, ..., hy-(ROWS-1)}; exact per row (clamped vertex), with a rounding margin
__device__ __forceinline__ float warp_block_qmin(float cA, float cB, float cC, float hx, float hy)
{
    float qmin;
#pragma unroll
    for (int r = 0; r < WARP_ROWS; ++r) {
        const float dy = hy - (float)r;
        const float dxs = fminf(hx, fmaxf(hx - (float)(WARP_COLS - 1), __fdividef(-cB * dy, cA)));
        const float ta = 0.5f * cA * dxs * dxs, tb = cB * dxs * dy, tc = 0.5f * cC * dy * dy;
        const float q = (ta + tb + tc) - 4e-6f * (fabsf(ta) + fabsf(tb) + fabsf(tc));
        qmin = (r == 0) ? q : fminf(qmin, q);
    }
    return qmin;
}

// gaussian_backward_kernel<RAW> (gauss_bwd.cu): raw leaves in, gradients accumulated into the leaves' gradient tensors
struct RawGrads {
    const float* f_rest; const float* opacity; const float* normal; const float* albedo; const float* roughness;
    const float* metallic;
    float* g_xyz; float* g_f_dc; float* g_f_rest; float* g_opacity; float* g_normal; float* g_albedo; float* g_roughness;
    float* g_metallic; float* g_log_scale; float* g_rot;
};
int launch_gaussian_backward_raw(int P, const GigsCamera& c, const void* geom, const Layout& L, const int32_t* radii,
                                 const float* accum, const float* xyz, const float* f_dc, const float* log_scale,
                                 const float* rot, float* g_means2D, const RawGrads& raw, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// Minimal column-major 3x3 matrix with the SAME expression shapes as the math library the
// reference uses for mat3*mat3 / transpose, so that nvcc contracts FMAs identically and tile keys
// come out bit-exact (SURVEY.md "Hard parts"). m[c][r], constructor takes columns.
// ---------------------------------------------------------------------------------------------
struct Mat3 {
    float m[3][3];
};
__device__ __forceinline__ Mat3 mat3_cols(float c00, float c01, float c02, float c10, float c11, float c12,
                                          float c20, float c21, float c22)
{
    Mat3 r;
    r.m[0][0] = c00; r.m[0][1] = c01; r.m[0][2] = c02;
    r.m[1][0] = c10; r.m[1][1] = c11; r.m[1][2] = c12;
    r.m[2][0] = c20; r.m[2][1] = c21; r.m[2][2] = c22;
    return r;
}
__device__ __forceinline__ Mat3 mat3_mul(const Mat3& a, const Mat3& b)
{
    // result[c][r] = a[0][r]*b[c][0] + a[1][r]*b[c][1] + a[2][r]*b[c][2]
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
            r.m[c][rr] = a.m[0][rr] * b.m[c][0] + a.m[1][rr] * b.m[c][1] + a.m[2][rr] * b.m[c][2];
        }
    }
    return r;
}
__device__ __forceinline__ Mat3 mat3_transpose(const Mat3& a)
{
    Mat3 r;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) r.m[c][rr] = a.m[rr][c];
    return r;
}

// reference auxiliary.h:58-66 / :68-77 expression order (column-major 4x4 in a float[16])
__device__ __forceinline__ float3 xform_point_4x3(const float3& p, const float* __restrict__ M)
{
    float3 t;
    t.x = M[0] * p.x + M[4] * p.y + M[8] * p.z + M[12];
    t.y = M[1] * p.x + M[5] * p.y + M[9] * p.z + M[13];
    t.z = M[2] * p.x + M[6] * p.y + M[10] * p.z + M[14];
    return t;
}
__device__ __forceinline__ float4 xform_point_4x4(const float3& p, const float* __restrict__ M)
{
    float4 t;
    t.x = M[0] * p.x + M[4] * p.y + M[8] * p.z + M[12];
    t.y = M[1] * p.x + M[5] * p.y + M[9] * p.z + M[13];
    t.z = M[2] * p.x + M[6] * p.y + M[10] * p.z + M[14];
    t.w = M[3] * p.x + M[7] * p.y + M[11] * p.z + M[15];
    return t;
}

// reference auxiliary.h:41-44: evaluated in double because of the un-suffixed literals
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// reference auxiliary.h:46-56
__device__ __forceinline__ void tile_rect(float px, float py, int max_radius, uint32_t gx, uint32_t gy,
                                          uint2& rmin, uint2& rmax)
{
    rmin.x = min(gx, (uint32_t)max((int)0, (int)((px - max_radius) / TILE_X)));
    rmin.y = min(gy, (uint32_t)max((int)0, (int)((py - max_radius) / TILE_Y)));
    rmax.x = min(gx, (uint32_t)max((int)0, (int)((px + max_radius + TILE_X - 1) / TILE_X)));
    rmax.y = min(gy, (uint32_t)max((int)0, (int)((py + max_radius + TILE_Y - 1) / TILE_Y)));
}

// float3 helpers with the reference's vec_math.h shapes (normalize = v * (1/sqrtf(dot)))
__device__ __forceinline__ float dot3(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(const float3& a, const float3& b)
{
    return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize3(const float3& v)
{
    float inv = 1.0f / sqrtf(dot3(v, v));
    return make_float3(v.x * inv, v.y * inv, v.z * inv);
}
__device__ __forceinline__ float3 sub3(const float3& a, const float3& b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 add3(const float3& a, const float3& b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }

// ---------------------------------------------------------------------------------------------
// L2 norms with the summation order of torch's CUDA reduction kernels (no fused multiply-add), so that the getters
// fused into our kernels return the bits of F.normalize / torch.norm (measured against torch 2.11 on B200 over
// 300k random vectors: 0 differing values; tools/raw_getter_check.py):
//   reduction over a contiguous inner dim of 4 / 3 (rotation [P,4], normal [P,3], rays [HW,3]):
//     threads along the reduced dim take a strided share, then a shuffle tree: (x0^2 + x2^2) + (x1^2 + x3^2),
//     (x0^2 + x2^2) + x1^2
//   reduction over the outer dim of a [3,H,W] map: one thread sums in order, (x0^2 + x1^2) + x2^2
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float torch_norm_inner4(float x0, float x1, float x2, float x3)
{
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x2, x2)), __fadd_rn(__fmul_rn(x1, x1), __fmul_rn(x3, x3))));
}
__device__ __forceinline__ float torch_norm_inner3(float x0, float x1, float x2)
{
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x2, x2)), __fmul_rn(x1, x1)));
}
__device__ __forceinline__ float torch_norm_outer3(float x0, float x1, float x2)
{
    return sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x0, x0), __fmul_rn(x1, x1)), __fmul_rn(x2, x2)));
}

// ---------------------------------------------------------------------------------------------
// PTX: mbarrier + bulk async copy (TMA, SASS UBLKCP) + vector reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes multiple of 16, both addresses 16-B aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// ask the copy engine to pull [gmem, gmem + bytes) into L2 (bytes multiple of 16, address 16-B aligned); no completion
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_f32(float* addr, float a)
{
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
// Blackwell packed FP32: d = a * b + c on two lanes of a 64-bit register pair (SASS FFMA2). Each half rounds exactly
// like a scalar fma, so accumulations stay bit-identical while the FMA instruction count halves.
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c)
{
    float2 d;
    asm("{\n"
        ".reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %5};\n"
        "mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;\n"
        "}\n"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float4 ld_nc_f4(const float* p)
{
    float4 r;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

}  // namespace gigs
