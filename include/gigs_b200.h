/*
 * gigs_b200 — C-ABI of the B200-native (sm_100a) GI-GS differentiable-rendering hot path.
 *
 * This is the drop-in boundary: plain pointers, sizes and a cudaStream_t (as void*); no torch
 * types.  Every entry point returns 0 on success, a negative value for an argument error and a
 * positive cudaError_t for a CUDA failure; gigs_last_error() gives the message.  All data
 * pointers are DEVICE pointers unless a field says "host".  NULL = "not provided", exactly as
 * the reference's C++ sees nullptr from empty tensors (cuda_rasterizer/forward.cu:217,254).
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to /root/reference/submodules/diff-gaussian-rasterization unless noted).
 */
#ifndef GIGS_B200_H
#define GIGS_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GIGS_ABI_VERSION 4

/* Per-view constants: the non-tensor fields of GaussianRasterizationSettings
 * (diff_gaussian_rasterization/__init__.py:31-51). Matrices are the transposed (column-major)
 * world_view / full_proj matrices, as in the reference. */
typedef struct GigsCamera {
    int32_t width, height;
    float tan_fovx, tan_fovy;
    float scale_modifier;
    int32_t sh_degree;   /* D: active degree */
    int32_t sh_coeffs;   /* M: stored coefficients per Gaussian (0 when colors_precomp is used) */
    int32_t prefiltered, debug, inference, argmax_depth;
    const float* viewmatrix; /* [16] */
    const float* projmatrix; /* [16] */
    const float* campos;     /* [3]  */
    const float* bg;         /* [3]  */
} GigsCamera;

/* Workspace sizes. geom/img/binning are the three opaque blobs the reference saves for backward
 * (rasterize_points.cu:183-188, rasterizer_impl.cu:155-199); their layout here is our own, a pure
 * function of (P, W, H, R). sort_bytes is transient scratch that need not outlive the call. */
typedef struct GigsSizes {
    uint64_t geom_bytes, img_bytes, binning_bytes, sort_bytes;
} GigsSizes;

int gigs_abi_version(void);
const char* gigs_last_error(void);
/* sizeof() of the argument structs, so a foreign-language binding can verify its mirror of the layout:
 * which = 0 GigsCamera, 1 GigsSizes, 2 GigsLayout, 3 GigsRasterFwd, 4 GigsRasterBwd, 5 GigsShade,
 * 6 GigsFrameLayout, 7 GigsFrame, 8 GigsLightLayout, 9 GigsAdamGroup, 10 GigsDensifyGroup,
 * 11 GigsStage1Layout, 12 GigsStage1; negative for an unknown id. */
int gigs_sizeof(int32_t which);
int gigs_raster_sizes(int32_t P, int32_t W, int32_t H, uint64_t R, GigsSizes* out);

/* Byte offsets of the fields inside the blobs (for tests that check keys / sort order / ranges
 * bit-for-bit against the reference's GeometryState / BinningState / ImageState).
 * Binning here is two small stable sorts instead of the reference's one 44-bit sort: the Gaussians
 * are argsorted by depth bits (g_order), instances are emitted in that order as (tile id, Gaussian id) pairs and
 * sorted by tile id only. The reference's 64-bit key of sorted slot i is
 * (s_tiles_sorted[i] << 32) | depth_bits(record[b_point_list[i]]). */
typedef struct GigsLayout {
    /* geom blob */
    uint64_t g_record;        /* float[P][24] packed blend record, see DESIGN.md */
    uint64_t g_cov3D;         /* float[P][6]  */
    uint64_t g_clamped;       /* uint8[P][4]  (3 used) */
    uint64_t g_tiles_touched; /* uint32[P] */
    uint64_t g_depth_keys;    /* uint32[P] depth bits (0xFFFFFFFF for Gaussians that touch no tile) */
    uint64_t g_order;         /* uint32[P] Gaussian ids in ascending (depth bits, id) order */
    uint64_t g_block_sums;    /* uint32[ceil(P/256)+1] */
    uint64_t g_num_rendered;  /* uint32[1] */
    /* img blob */
    uint64_t i_final_T;       /* float[N] */
    uint64_t i_n_contrib;     /* uint32[N] */
    uint64_t i_ranges;        /* uint2[T] */
    /* binning blob */
    uint64_t b_point_list;    /* uint32[R] sorted Gaussian ids */
    /* sort scratch */
    uint64_t s_tiles_sorted;   /* uint32[R] tile id of each sorted slot */
    uint64_t s_tiles_unsorted; /* uint32[R] tile ids in emission order */
    uint64_t s_vals_unsorted;  /* uint32[R] Gaussian ids in emission order */
} GigsLayout;
int gigs_raster_layout(int32_t P, int32_t W, int32_t H, uint64_t R, GigsLayout* out);

/* Replaces RasterizeGaussiansCUDA (rasterize_points.cu:130-252) ->
 * CudaRasterizer::Rasterizer::forward (cuda_rasterizer/rasterizer_impl.cu:486-672).
 * Two phases because num_rendered sizes the binning blob, which the caller owns:
 *   begin : preprocess + tile-count scan, num_rendered read back (the depth argsort of the Gaussians is queued
 *           behind the read-back so the GPU keeps working while the host sizes the binning blob)
 *   finish: pair emission in depth order, onesweep radix sort by tile id, tile ranges, G-buffer blend
 * Output maps are CHW planar float32, fully written by finish (no pre-fill needed). */
typedef struct GigsRasterFwd {
    int32_t P;
    int32_t material_only; /* 1: the caller does not read out_color / out_pos (the PBR stage: train.py uses the SH radiance
                              image only in the first stage and for logging): SH evaluation and those 6 blend channels are
                              skipped, the two outputs are left untouched. 2: the caller reads out_color, out_normal, out_depth and
                              out_opacity only (the first training stage): the material channels, out_pos and out_normal_view
                              are neither blended nor written. 0 = everything, like the reference. */
    GigsCamera cam;
    const float* means3D;        /* [P,3] */
    const float* shs;            /* [P,M,3] or NULL */
    const float* colors_precomp; /* [P,3] or NULL */
    const float* opacities;      /* [P] */
    const float* normal;         /* [P,3] */
    const float* albedo;         /* [P,3] */
    const float* roughness;      /* [P] */
    const float* metallic;       /* [P] */
    const float* scales;         /* [P,3] or NULL */
    const float* rotations;      /* [P,4] or NULL */
    const float* cov3D_precomp;  /* [P,6] or NULL */
    float* out_color;       /* [3,H,W] */
    float* out_opacity;     /* [1,H,W] */
    float* out_depth;       /* [1,H,W] */
    float* out_normal;      /* [3,H,W] */
    float* out_normal_view; /* [3,H,W] */
    float* out_pos;         /* [3,H,W] */
    float* out_albedo;      /* [3,H,W] */
    float* out_roughness;   /* [1,H,W] */
    float* out_metallic;    /* [1,H,W] */
    int32_t* radii;         /* [P] */
    void* geom;    uint64_t geom_bytes;
    void* img;     uint64_t img_bytes;
    void* binning; uint64_t binning_bytes; /* finish only */
    void* sort;    uint64_t sort_bytes;    /* finish only */
    uint32_t* pinned_num_rendered;         /* host, page-locked, 4 bytes; may be NULL */
    int64_t num_rendered;                  /* out of begin, in of finish */
    void* stream;                          /* cudaStream_t */
} GigsRasterFwd;
int gigs_raster_forward_begin(GigsRasterFwd* a);
int gigs_raster_forward_finish(GigsRasterFwd* a);

/* Replaces LiteRasterizeGaussiansCUDA (rasterize_points.cu:39-127): colour + opacity + depth only.
 * Same struct; normal/albedo/roughness/metallic inputs and the six PBR outputs are ignored. */
int gigs_lite_forward_finish(GigsRasterFwd* a);

/* Replaces RasterizeGaussiansBackwardCUDA (rasterize_points.cu:254-364) ->
 * CudaRasterizer::Rasterizer::backward (cuda_rasterizer/rasterizer_impl.cu:676-803).
 * Upstream gradients may be NULL (treated as all-zero; the result is identical to the
 * reference fed a materialised zero tensor). Every output element is written exactly once. */
typedef struct GigsRasterBwd {
    int32_t P;
    int32_t _pad;
    int64_t num_rendered;
    GigsCamera cam;
    const float* means3D; const float* shs; const float* colors_precomp;
    const float* normal; const float* albedo; const float* roughness; const float* metallic;
    const float* scales; const float* rotations; const float* cov3D_precomp;
    const int32_t* radii;
    const void* geom; const void* binning; const void* img;
    const float* dL_dpix_depth;     /* [1,H,W] */
    const float* dL_dpix;           /* [3,H,W] */
    const float* dL_dpix_opacity;   /* [1,H,W] */
    const float* dL_dpix_normal;    /* [3,H,W] */
    const float* dL_dpix_albedo;    /* [3,H,W] */
    const float* dL_dpix_roughness; /* [1,H,W] */
    const float* dL_dpix_metallic;  /* [1,H,W] */
    float* accum;        /* scratch float[P][20], zero-initialised by the call */
    float* dL_dmean2D;   /* [P,3] (xy + |grad| in z, backward.cu:616-619) */
    float* dL_dconic;    /* [P,4] may be NULL (intermediate in the reference) */
    float* dL_dopacity;  /* [P]   */
    float* dL_dcolor;    /* [P,3] */
    float* dL_dnormal;   /* [P,3] */
    float* dL_dalbedo;   /* [P,3] */
    float* dL_droughness;/* [P]   */
    float* dL_dmetallic; /* [P]   */
    float* dL_dmean3D;   /* [P,3] */
    float* dL_dcov3D;    /* [P,6] */
    float* dL_dsh;       /* [P,M,3] or NULL */
    float* dL_dscale;    /* [P,3] or NULL */
    float* dL_drot;      /* [P,4] or NULL */
    void* stream;
} GigsRasterBwd;
int gigs_raster_backward(GigsRasterBwd* a);

/* Replaces markVisible (rasterize_points.cu:366-385, rasterizer_impl.cu:54-66). present: uint8[P]. */
int gigs_mark_visible(int32_t P, const float* means3D, const float* viewmatrix, uint8_t* present, void* stream);

/* Replaces depthToNormal (rasterize_points.cu:387-405, forward.cu:914-1032).
 * normal_map/depth_pos [3,H,W]; fully written (zeros where the reference leaves its fill). */
int gigs_depth_to_normal(int32_t W, int32_t H, float fx, float fy, const float* viewmatrix,
                         const float* depth, float* normal_map, float* depth_pos, void* stream);

/* 3x3 zero-padded lower-median over C planes (kornia.filters.median_blur as called at
 * diff_gaussian_rasterization/__init__.py:478,504); window containing NaN/inf -> NaN. */
int gigs_median3x3(int32_t C, int32_t W, int32_t H, const float* in, float* out, void* stream);
/* backward of the above: routes each output gradient to the selected input element. */
int gigs_median3x3_backward(int32_t C, int32_t W, int32_t H, const float* in, const float* grad_out,
                            float* grad_in /* zero-filled by the call */, void* stream);

/* 3x3 bilateral blur, reflect border, L1 colour distance, sigma_color, sigma_space
 * (kornia.filters.bilateral_blur as called at diff_gaussian_rasterization/__init__.py:491). */
int gigs_bilateral3x3(int32_t C, int32_t W, int32_t H, float sigma_color, float sigma_space,
                      const float* in, float* out, void* stream);

/* Fused GaussianRasterizer.forward post-pass (diff_gaussian_rasterization/__init__.py:475-504):
 * median(depth) -> depth_to_normal -> bilateral(normal) and median(depth_pos), one kernel. */
int gigs_geometry_chain(int32_t W, int32_t H, float fx, float fy, const float* viewmatrix,
                        const float* depth, int32_t derive_normal,
                        float* normal_from_depth /*[3,H,W]*/, float* depth_pos_filter /*[3,H,W]*/,
                        void* stream);

/* Scratch of the SSAO / SSR march for a W x H frame: the block (min, max) table of the position map's depth plane
 * that lets the march skip the depth gather of probes that cannot hit (<= 40 KB + 16). The caller owns it, 16-byte
 * aligned; it is written by every call and holds nothing between calls. With scratch == NULL (or too small) the
 * march runs without the block test: same result, slower. */
uint64_t gigs_gi_scratch_bytes(int32_t W, int32_t H);

/* Replaces SSAO (rasterize_points.cu:407-436, forward.cu:635-724). */
int gigs_ssao(int32_t W, int32_t H, float fx, float fy, float radius, float bias, float thick,
              float delta, int32_t step, int32_t start, const float* normal, const float* pos,
              float* occlusion, void* scratch, uint64_t scratch_bytes, void* stream);

/* Replaces SSR (rasterize_points.cu:438-477, forward.cu:726-909). */
int gigs_ssr(int32_t W, int32_t H, float fx, float fy, float radius, float bias, float thick,
             float delta, int32_t step, int32_t start, const float* normal, const float* pos,
             const float* rgb, const float* albedo, const float* roughness, const float* metallic,
             const float* F0, float* color, float* abd, void* scratch, uint64_t scratch_bytes, void* stream);

/* Number of probes (one sample position -> one depth test) the SSAO/SSR march of this G-buffer has to evaluate:
 * the reference's loop (forward.cu:691-716 / :805-846) counted up to and including the probe that ends a direction,
 * without the pixels whose sample positions are all NaN and the zero-weight directions (theta = 0), neither of which
 * can change the result. Measurement helper (roofline numerator); count is two device words: [0] the probes,
 * [1] (when block_minmax != NULL) the probes whose depth window intersects the (min, max) of pos.z over the
 * block x block pixel block they land in — block_minmax is [ceil(H/block), ceil(W/block), 2] floats. */
int gigs_gi_count_probes(int32_t W, int32_t H, float fx, float fy, float radius, float bias, float thick,
                         float delta, int32_t step, int32_t start, const float* normal, const float* pos,
                         const float* block_minmax, int32_t block, uint64_t* count, void* stream);

/* Tuning knobs of the march: probe pairs evaluated per inner step (1 or 2; 0 = the reference-order loop for every
 * pixel), and whether the block (min, max) depth test runs before a probe's depth gather. Results are bit-identical
 * for every setting. Process-wide; defaults 1, 1. */
int gigs_gi_tune(int32_t pairs_per_step, int32_t block_test);

/* Self-check of the cube-map seam handling: the shading kernels fold a bilinear tap that steps over a face edge onto
 * the adjacent face with a 24-entry table; this compares the table against the reference statement of the fold for
 * every tap position of a level of resolution res (coordinates in [-1, res]) and writes the number of mismatches to
 * *mismatches (device, one int32). */
int gigs_cube_wrap_selfcheck(int32_t res, int32_t* mismatches, void* stream);

/* The kernels of the frame are launched with programmatic stream serialization (PTX griddepcontrol): each waits for
 * the previous grid before its first global access, so results do not change; only the launch latency between the
 * ~30 dependent kernels of a frame is hidden. on = 0 / 1 switches it for the process (default 1, or the environment
 * variable GIGS_PDL=0); on < 0 only queries. Returns the previous setting. */
int gigs_set_dependent_launch(int32_t on);

/* Kernels this library has launched in this process so far (frame, first-stage, loss, optimiser and light-build
 * paths; a benchmark reads it before and after its timed region). */
uint64_t gigs_launch_count(void);

/* Replaces SSR_BACKWARD (rasterize_points.cu:479-510). The reference's Python never calls its
 * kernel (diff_gaussian_rasterization/__init__.py:666-673); the live semantics are
 * grad_albedo = grad_color * abd, zeros for roughness/metallic, which is what this computes. */
int gigs_ssr_backward(int32_t W, int32_t H, const float* grad_color, const float* abd,
                      float* grad_albedo, float* grad_roughness, float* grad_metallic, void* stream);

/* Fused split-sum deferred shading: pbr_shading (/root/reference/pbr/shade.py:104-237) with
 * CubemapLight.get_mip (pbr/light.py:142-152) and nvdiffrast texture semantics restated.
 * Maps are CHW planar (the layout the rasterizer produces); the HWC permutes of train.py:343-348
 * are folded into the addressing. */
typedef struct GigsShade {
    int32_t W, H;
    int32_t n_spec_levels;          /* len(light.specular), <= 8 */
    int32_t spec_res[8];            /* per-level face resolution */
    const float* spec[8];           /* per-level [6,res,res,3] */
    int32_t diffuse_res;            /* 16 */
    const float* diffuse;           /* [6,res,res,3] */
    const float* brdf_lut;          /* [256,256,2] */
    int32_t lut_res;
    int32_t tone, gamma, has_metallic, has_occlusion;
    float min_roughness, max_roughness; /* 0.08, 0.5 */
    const float* normals;    /* [3,H,W] */
    const float* view_dirs;  /* [3,H,W] */
    const float* albedo;     /* [3,H,W] */
    const float* roughness;  /* [1,H,W] */
    const float* metallic;   /* [1,H,W] or NULL */
    const float* occlusion;  /* [1,H,W] or NULL */
    const uint8_t* mask;     /* [H,W] */
    const float* background; /* [3,H,W] or NULL (zeros) */
    /* forward outputs, CHW */
    float* render_rgb; float* diffuse_rgb; float* specular_rgb; float* diffuse_light;
    /* backward (all may be NULL in forward) */
    const float* g_render_rgb; const float* g_diffuse_rgb; const float* g_specular_rgb;
    float* g_albedo; float* g_roughness; float* g_metallic; /* written */
    float* g_diffuse_tex;  /* [6,res,res,3], accumulated (+=) */
    float* g_spec[8];      /* accumulated (+=) */
    void* stream;
} GigsShade;
int gigs_shade_forward(GigsShade* a);
int gigs_shade_backward(GigsShade* a);

/* ------------------------------------------------------------------------------------------------
 * Fused PBR-stage frame: the whole per-view training step of /root/reference/train.py:266-404 that lies on the
 * hot path, as TWO calls (forward, backward) instead of ~200 framework launches:
 *   forward : GaussianModel getters fused into preprocess (scene/gaussian_model.py:178-266, when raw_params) ->
 *             rasterize (rasterizer_impl.cu:486-672) -> geometry chain + SSAO
 *             (diff_gaussian_rasterization/__init__.py:475-517) -> render() post-processing
 *             (gaussian_renderer/__init__.py:157-199) + pbr_shading (pbr/shade.py:104-237) + srgb_to_linear
 *             (train.py:70-75) in one deferred kernel -> SSR (forward.cu:726-909) -> linear_to_srgb + 3x3 median +
 *             L1 / "lamb" loss (train.py:380-386,402-404) in one kernel.
 *   backward: median / sRGB / SSR (g*abd, __init__.py:671-673) / shading backward in one kernel -> material-only
 *             blend backward (backward.cu:404-630 with dL/dcolor = dL/dopacity = 0, which is what the PBR stage
 *             feeds it) -> sigmoid backward into the parameter gradients (accumulated, +=).
 * In the PBR stage normals / occlusion / SSR inputs are detached by the caller (train.py:343-351,378), so albedo,
 * roughness, metallic and the light textures are the only parameters that receive a gradient; that is the
 * reference's semantics, not a shortcut.
 * Every intermediate map lives in ONE caller-owned blob (`maps`, layout from gigs_frame_layout) and stays
 * inspectable after the call. */
typedef struct GigsFrameLayout {
    /* float planes, element offsets in BYTES into the maps blob; [c,H,W] planar */
    uint64_t color, opacity, depth, normal, normal_view, pos, albedo, roughness, metallic; /* rasterizer outputs */
    uint64_t normal_from_depth, depth_pos, occlusion;                                      /* geometry chain, SSAO */
    uint64_t shade_normal;   /* [3] normalised, median-filtered, view-rotated normal the shading uses */
    uint64_t ssr_normal;     /* [3] normalised + median-filtered out_normal_view the SSR uses */
    uint64_t render_direct, linear_rgb, F0, rough_remap, metal_used;
    uint64_t ssr_color, ssr_abd, render_rgb;
    uint64_t g_rgb;          /* [3] dL/d render_rgb */
    uint64_t g_albedo, g_roughness, g_metallic; /* dL/d G-buffer maps (backward) */
    uint64_t mask;           /* uint8 [H,W] normal_mask */
    uint64_t median_sel;     /* uint8 [3,H,W] window index the IRR median selected (255 = none) */
    uint64_t tex_scratch;    /* float scratch: private copies of the small light-gradient textures (backward) */
    uint64_t partials;       /* float scratch for the deterministic loss reduction */
    uint64_t stats;          /* float[8]: loss, l1_mean, mask_count, sum((1-rough)*mask), sum(metal*mask), brdf_tv, - */
    uint64_t tv_edge;        /* float[2,H,W]: edge weights of the BRDF TV prior (below / right of each pixel) */
    uint64_t total_bytes;
} GigsFrameLayout;
int gigs_frame_layout(int32_t W, int32_t H, GigsFrameLayout* out);

typedef struct GigsFrame {
    int32_t P;
    int32_t raw_params;   /* 1: parameter pointers are pre-activation leaves; sh_dc=[P,1,3], sh_rest=[P,M-1,3].
                             0: activated tensors as GaussianRasterizer takes them; sh_dc=[P,M,3], sh_rest=NULL */
    GigsCamera cam;
    const float* means3D; const float* sh_dc; const float* sh_rest; const float* opacities; const float* normal;
    const float* albedo; const float* roughness; const float* metallic; const float* scales; const float* rotations;
    float radius, bias, thick, delta; int32_t step, start;      /* GI settings (GaussianRasterizationSettings) */
    int32_t indirect, use_metallic, tone, gamma;                 /* train.py --indirect --metallic --tone --gamma */
    int32_t n_spec_levels; int32_t spec_res[8]; const float* spec[8];
    int32_t diffuse_res; const float* diffuse; const float* brdf_lut; int32_t lut_res;
    float min_roughness, max_roughness;
    const float* canonical_rays; /* [H*W,3] (scene/__init__.py:157-167) */
    const float* gt_image;       /* [3,H,W]; NULL = forward only, no loss */
    float loss_scale, lamb_weight;
    float brdf_tv_weight;        /* train.py:388-402 BRDF smoothness prior (get_masked_tv_loss); 0 = off */
    int32_t material_only;       /* GigsRasterFwd.material_only for this frame: the `color` / `pos` maps are not produced */
    void* geom; uint64_t geom_bytes; void* img; uint64_t img_bytes;
    void* binning; uint64_t binning_bytes; void* sort; uint64_t sort_bytes;
    void* maps; uint64_t maps_bytes;
    int32_t* radii;              /* [P] */
    float* accum;                /* [P,20] scratch (backward) */
    uint32_t* pinned_num_rendered;
    int64_t num_rendered;        /* out of forward */
    int32_t resume;              /* in: 1 = preprocess already ran for this frame (retry after GIGS_E_GROW) */
    int32_t skip_geometry;       /* in: 1 = when no GI march runs (start >= step) the depth -> normal / position chain is
                                    left out: normal_from_depth is a first-stage loss term and depth_pos only feeds the
                                    march (train.py:290-381), so loss, gradients and every other map are unchanged; the
                                    two maps and the march's other inputs (ssr_normal, linear_rgb, F0) are then not written.
                                    0 (what a zero-filled struct says): always run it */
    uint64_t need_binning_bytes, need_sort_bytes; /* out: sizes this frame needs (set when returning GIGS_E_GROW) */
    /* backward outputs, accumulated (+=); per-Gaussian ones are in raw-parameter space when raw_params */
    float* g_albedo; float* g_roughness; float* g_metallic;   /* [P,3], [P], [P] */
    float* g_diffuse_tex; float* g_spec[8];
    void* gt_ready_event;        /* cudaEvent_t or NULL: the stream waits on it right before the loss kernel, so the
                                    host->device copy of gt_image (issued on another stream) overlaps the rasterizer */
    void* light_ready_event;     /* cudaEvent_t or NULL: recorded in backward once the light-texture gradients are final
                                    (before the blend backward), so their all-reduce can overlap the rest */
    void* stream;
} GigsFrame;
#define GIGS_E_GROW (-5) /* binning / sort workspace too small: grow to need_*_bytes and call again with resume=1 */
int gigs_frame_forward(GigsFrame* f);
int gigs_frame_backward(GigsFrame* f);

/* Replaces latlong_to_cubemap (relight.py:92-112, render.py:64-84: the relight / eval sweeps turn a lat-long HDR
 * environment map [env_h, env_w, channels] into the light's base cubemap [6, res, res, channels] before build_mips):
 * per cube texel direction d = normalize(cube_to_dir), (u, v) = (atan2(d.x, -d.z) / 2pi + 0.5, acos(d.y) / pi),
 * bilinear with both axes wrapping (nvdiffrast dr.texture, filter_mode "linear", boundary mode "wrap"). */
int gigs_latlong_to_cubemap(int32_t env_h, int32_t env_w, int32_t channels, const float* env, int32_t res,
                            float* cube, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cubemap prefilter = CubemapLight.build_mips (/root/reference/pbr/light.py:154-170), SURVEY.md §8f-1: the producer of
 * the light textures the shading consumes, run by the reference once per PBR training step (train.py:340).
 * Replaces the nvdiffrec renderutils plugin entry points diffuse_cubemap_fwd/bwd, specular_bounds,
 * specular_cubemap_fwd/bwd (pbr/renderutils/c_src/torch_bindings.cpp:740-889 -> c_src/cubemap.cu:110-350) and
 * cubemap_mip (pbr/light.py:54-79). Cubemaps are [6,res,res,3] float32 (the reference's NHWC layout).
 *   table  : float[6*res*res][4]  = unit direction of each texel centre (cubemap.cu:32-47) + pixel_area (:17-30)
 *   bounds : int16[6*res*res][6][4] = (xmin,xmax,ymin,ymax) of the GGX cone on each face (cubemap.cu:182-246)
 * Both depend only on (res, cutoff): build once, reuse every step. Backward entry points WRITE grad_in (gather, no
 * atomics, deterministic); the specular pair is fused with the col / wsum division of renderutils/ops.py:456. */
int gigs_cubemap_table(int32_t res, float* table, void* stream);
int gigs_specular_bounds(int32_t res, float costheta_cutoff, const float* table, int16_t* bounds, void* stream);
int gigs_cubemap_mip_forward(int32_t res_out, const float* in /*[6,2r,2r,3]*/, float* out /*[6,r,r,3]*/, void* stream);
int gigs_cubemap_mip_backward(int32_t res_coarse, const float* grad_coarse, float* grad_fine /*[6,2r,2r,3]*/,
                              int32_t accumulate, void* stream);
int gigs_diffuse_cubemap_forward(int32_t res, const float* table, const float* cubemap, float* out, void* stream);
int gigs_diffuse_cubemap_backward(int32_t res, const float* table, const float* grad_out, float* grad_in, void* stream);
int gigs_specular_cubemap_forward(int32_t res, const float* table, const int16_t* bounds, float roughness,
                                  float costheta_cutoff, const float* cubemap, float* out /*[6,r,r,3]*/,
                                  float* wsum /*[6,r,r]*/, void* stream);
int gigs_specular_cubemap_backward(int32_t res, const float* table, const int16_t* bounds, float roughness,
                                   float costheta_cutoff, const float* grad_out, const float* wsum, float* grad_in,
                                   void* stream);

/* The whole light in one go = CubemapLight.build_mips (pbr/light.py:154-170) and its backward, for the training step
 * (train.py:340 rebuilds the mips every PBR iteration). Two caller-owned blobs:
 *   workspace (layout.total_bytes, a pure function of (base_res, min_res)): tables, bounds, the mip chain, the filtered
 *     levels, wsum, the texture-gradient span and scratch;
 *   weights (layout.weights_bytes, known after gigs_light_prepare; ~1.5 GB at base_res 256): the filters as STORED
 *     sparse operators. Their weights depend only on (resolution, roughness, cutoff), so they are evaluated once, with
 *     the reference's per-pair arithmetic to the bit, and every step streams them from HBM instead of recomputing
 *     ~100 instructions per (output texel, light texel) pair. Pass weights = NULL to gigs_light_build / _backward to
 *     compute the same sums on the fly instead (no extra memory, ~5x slower).
 * Sequence:
 *   gigs_light_layout  fills res[] (base_res, base_res/2, ... min_res), roughness[] (light.py:165-170) and the offsets;
 *                      the caller then sets cutoff[l] = cos(theta) keeping `cutoff` of the GGX energy at roughness[l]
 *                      (renderutils/ops.py:430-441 __ndfBounds — a 1e6-sample numpy cumulative sum in the reference,
 *                      kept on the host side of the binding so that the threshold is the reference's to the bit);
 *   gigs_light_prepare builds table[] / bounds[], counts the operators' runs and weights (one stream synchronise), fills
 *                      n_runs[] / n_weights[] / w_*[] / weights_bytes and clears the gradient span;
 *   gigs_light_weights fills the weights blob (run records, forward operators + wsum in the reference's summation
 *                      order, transposed backward operators with 1 / wsum folded in);
 *   gigs_light_build   base [6,R,R,3] -> chain[] (average-pool mips, 16-byte padded texels), spec[l] [6,r,r,3] (GGX
 *                      filtered, divided by wsum[l]) and diffuse [6,min_res,min_res,3]: 2 launches;
 *   gigs_light_backward reads the texture gradients g_spec[l] / g_diffuse (the span [grad_begin, +grad_bytes) that the
 *                      shading backward accumulates into: pass those pointers as GigsFrame.g_spec / g_diffuse_tex) and
 *                      writes (accumulate=0) or adds (accumulate=1) d loss / d base; clear_grads=1 re-zeroes the span.
 * Filter index f in the [GIGS_MAX_LIGHT_LEVELS + 1] arrays: 0..n_levels-1 = GGX levels, n_levels = the cosine filter. */
#define GIGS_MAX_LIGHT_LEVELS 8
typedef struct GigsLightLayout {
    int32_t n_levels;
    int32_t res[GIGS_MAX_LIGHT_LEVELS];
    float roughness[GIGS_MAX_LIGHT_LEVELS];
    float cutoff[GIGS_MAX_LIGHT_LEVELS];
    int32_t pad_;
    /* offsets into the workspace blob */
    uint64_t table[GIGS_MAX_LIGHT_LEVELS], bounds[GIGS_MAX_LIGHT_LEVELS], chain[GIGS_MAX_LIGHT_LEVELS];
    uint64_t spec[GIGS_MAX_LIGHT_LEVELS], wsum[GIGS_MAX_LIGHT_LEVELS], gq[GIGS_MAX_LIGHT_LEVELS];
    uint64_t g_chain[GIGS_MAX_LIGHT_LEVELS], g_spec[GIGS_MAX_LIGHT_LEVELS];
    uint64_t rowptr[GIGS_MAX_LIGHT_LEVELS + 1], wptr[GIGS_MAX_LIGHT_LEVELS + 1];
    uint64_t counts, totals;
    uint64_t diffuse, gq_diffuse, g_diffuse_in, g_diffuse;
    uint64_t grad_begin, grad_bytes, total_bytes;
    /* the stored operators (filled by gigs_light_prepare): sizes and offsets into the weights blob */
    int32_t lanes_log2[GIGS_MAX_LIGHT_LEVELS + 1];   /* lanes per 4-texel block = length the runs are cut into */
    int32_t pad2_;
    uint64_t n_runs[GIGS_MAX_LIGHT_LEVELS + 1], n_weights[GIGS_MAX_LIGHT_LEVELS + 1];   /* run pieces, float4 entries */
    uint64_t w_rows[GIGS_MAX_LIGHT_LEVELS + 1], w_fwd[GIGS_MAX_LIGHT_LEVELS + 1], w_bwd[GIGS_MAX_LIGHT_LEVELS + 1];
    uint64_t weights_bytes;
} GigsLightLayout;
int gigs_light_layout(int32_t base_res, int32_t min_res, GigsLightLayout* layout);
int gigs_light_prepare(GigsLightLayout* layout, void* workspace, void* stream);
int gigs_light_weights(const GigsLightLayout* layout, void* workspace, void* weights, void* stream);
int gigs_light_build(const GigsLightLayout* layout, const float* base, void* workspace, const void* weights, void* stream);
int gigs_light_backward(const GigsLightLayout* layout, void* workspace, const void* weights, float* grad_base,
                        int32_t accumulate, int32_t clear_grads, void* stream);

/* Env-map smoothness prior of the PBR-stage loss (/root/reference/train.py:406-420): a seamless bilinear lookup of the
 * base cubemap [6,R,R,3] along `dirs` [env_h*env_w,3] (get_envmap_dirs, train.py:145-157: a 512x1024 lat-long grid),
 * loss = mean((env[1:] - env[:-1])^2) + mean((env[:,1:] - env[:,:-1])^2).
 * loss_out (device float, may be NULL) = [accumulate_loss ? loss_out : 0] + scale * loss; grad_base (may be NULL) +=
 * scale * d loss / d base. scale = env_tv_weight * loss_scale. scratch == NULL: size query into *scratch_bytes. */
int gigs_env_tv(int32_t base_res, const float* base, const float* dirs, int32_t env_h, int32_t env_w, float scale,
                void* scratch, uint64_t* scratch_bytes, float* grad_base, float* loss_out, int32_t accumulate_loss,
                void* stream);

/* ---- The fused FIRST-STAGE frame ------------------------------------------------------------------------------------
 * One view of /root/reference/train.py:266-328 (iteration <= pbr_iteration) as two calls: GaussianModel getters +
 * gaussian_renderer.render(derive_normal=True) + loss = (1 - lambda_dssim) * L1 + lambda_dssim * (1 - SSIM)
 * + normal_weight * F.l1_loss(normal_map[:, mask], normal_map_from_depth[:, mask]) + normal_tv_weight * get_tv_loss(gt,
 * normal_map), then the whole backward into the gradient tensors of the ten raw parameter tensors (accumulated, +=,
 * like autograd). Parameters are the trainer's PRE-activation leaves (scene/gaussian_model.py:55-66: _xyz,
 * _features_dc [P,1,3], _features_rest [P,M-1,3], _opacity, _normal, _albedo, _roughness, _metallic, _scaling,
 * _rotation). Workspaces as in GigsFrame; maps is sized by gigs_stage1_layout. stats (float[8] at layout.stats):
 * [0] loss_scale * image loss, [1] L1, [2] SSIM, [4] loss_scale * normal loss, [5] normal L1, [6] normal TV; the
 * step's loss is stats[0] + stats[4]. g_means2D ([P,3], written, may be NULL) is the screen-space gradient the
 * densification statistics read (viewspace_point_tensor.grad). Returns GIGS_E_GROW (-5) like gigs_frame_forward. */
typedef struct GigsStage1Layout {
    uint64_t color, opacity, depth, normal, normal_view, pos, albedo, roughness, metallic;  /* rasterizer outputs */
    uint64_t normal_from_depth, depth_pos;                  /* geometry chain */
    uint64_t normals_view, nfd_unit;                        /* render()'s "normal_map" and "normal_map_from_depth" */
    uint64_t g_color, g_normals_view, g_normal;             /* gradients of the image, normals_view, rasterizer normal */
    uint64_t median_sel, mask;                              /* uint8 [3,H,W] / [H,W] */
    uint64_t loss_scratch, loss_scratch_bytes, nloss_scratch, nloss_scratch_bytes, stats, total_bytes;
} GigsStage1Layout;
typedef struct GigsStage1 {
    int32_t P; int32_t pad0_;
    GigsCamera cam;
    const float* xyz; const float* f_dc; const float* f_rest; const float* opacity; const float* normal;
    const float* albedo; const float* roughness; const float* metallic; const float* log_scale; const float* rot;
    const float* gt_image;           /* [3,H,W]; NULL = forward only, no loss */
    float lambda_dssim, normal_weight, normal_tv_weight, loss_scale;
    void* geom; uint64_t geom_bytes; void* img; uint64_t img_bytes;
    void* binning; uint64_t binning_bytes; void* sort; uint64_t sort_bytes;
    void* maps; uint64_t maps_bytes;
    int32_t* radii; float* accum; void* pinned_num_rendered;
    int64_t num_rendered; int32_t resume; int32_t pad1_;
    uint64_t need_binning_bytes, need_sort_bytes;
    float* g_xyz; float* g_f_dc; float* g_f_rest; float* g_opacity; float* g_normal; float* g_albedo;
    float* g_roughness; float* g_metallic; float* g_log_scale; float* g_rot;
    float* g_means2D;
    void* gt_ready_event;
    void* stream;
} GigsStage1;
int gigs_stage1_layout(int32_t W, int32_t H, GigsStage1Layout* out);
int gigs_stage1_forward(GigsStage1* f);
int gigs_stage1_backward(GigsStage1* f);

/* ---- The optimiser step -------------------------------------------------------------------------
 * Replaces `gaussians.optimizer.step(); gaussians.optimizer.zero_grad(); light_optimizer.step();
 * light_optimizer.zero_grad(); cubemap.clamp_(min=0.0)` of /root/reference/train.py:516-523 — torch.optim.Adam over
 * the 10 groups of scene/gaussian_model.py:318-359 (eps 1e-15) and the light's base cubemap (train.py:215-218, eps
 * 1e-8) — by ONE launch over all groups. Each group is one contiguous float tensor with its two moment tensors.
 * grad == NULL says "this gradient is all zero" (what the reference's rasterizer backward returns for every
 * non-material input in the PBR stage): the moments still decay and the parameter still moves exactly as torch's
 * Adam moves it, but the gradient is neither read nor cleared. `step` is torch's state['step'] AFTER its increment
 * (1 for the first update). Scalars are doubles because torch forms lr / (1 - beta1^t) and sqrt(1 - beta2^t) from
 * Python floats. The group array is HOST memory; at most 24 groups per call. */
typedef struct GigsAdamGroup {
    float* param;
    float* grad;        /* may be NULL (see above) */
    float* exp_avg;
    float* exp_avg_sq;
    uint64_t count;     /* elements */
    double lr, beta1, beta2, eps;
    int32_t step;
    int32_t clamp_min0; /* param.clamp_(min=0) after the update (the cubemap) */
    int32_t clear_grad; /* write zeros over grad after reading it (zero_grad fused into the pass) */
    int32_t pad_;
} GigsAdamGroup;
int gigs_adam_step(int32_t n_groups, const GigsAdamGroup* groups, void* stream);

/* optimizer.zero_grad() (train.py:518,522) of up to 8 spans [begin, end) (element offsets) of one float buffer, in one
 * launch that chains with the frame's kernels (the fused frame path writes material and light gradients only, so a
 * training step clears those two spans instead of the whole buffer). */
int gigs_clear_spans(float* base, int32_t n_spans, const uint64_t* begin, const uint64_t* end, void* stream);

/* Densification statistics of one view (/root/reference/train.py:489-495 + GaussianModel.add_densification_stats,
 * scene/gaussian_model.py:933-945) for the Gaussians with radii > 0: max_radii2D = max(max_radii2D, radii),
 * xyz_gradient_accum += |grad2D.xy|, xyz_gradient_accum_abs += |gx| + |gy|, xyz_gradient_accum_abs_max =
 * max(., |gx| + |gy|), denom += 1. grad2D is [P, grad_stride] (the reference's means2D gradient is [P,3]);
 * the *_abs, *_abs_max and max_radii2D outputs may be NULL. */
int gigs_densify_stats(int32_t P, const int32_t* radii, const float* grad2D, int32_t grad_stride,
                       float* xyz_gradient_accum, float* xyz_gradient_accum_abs, float* xyz_gradient_accum_abs_max,
                       float* denom, float* max_radii2D, void* stream);

/* ---- Image loss of the first training stage -------------------------------------------------------
 * Replaces `(1 - lambda_dssim) * l1_loss(image, gt) + lambda_dssim * (1 - ssim(image, gt))` and its autograd backward
 * (/root/reference/train.py:320-322, utils/loss_utils.py:19-20 l1_loss, :40-100 ssim: 11x11 Gaussian window sigma 1.5,
 * zero padding, per channel, mean over all elements) by one forward and one backward kernel. image, gt: [C,H,W].
 * loss_out (device float[3], may be NULL): [0] = (accumulate_loss ? loss_out[0] : 0) + loss_scale * loss, [1] = the L1
 * term, [2] = SSIM. grad_image (may be NULL): d(loss_scale * loss)/d image, times *upstream when upstream (a device
 * scalar, e.g. autograd's grad_output) is given; added to grad_image when accumulate_grad. lambda_dssim = 1 and
 * loss_scale = -1 give -(1 - ssim), i.e. plain SSIM's gradient up to the constant.
 * scratch == NULL: size query into *scratch_bytes (3 derivative maps + per-CTA partial sums). */
int gigs_image_loss(int32_t C, int32_t W, int32_t H, const float* image, const float* gt, float lambda_dssim,
                    float loss_scale, void* scratch, uint64_t* scratch_bytes, float* loss_out, int32_t accumulate_loss,
                    float* grad_image, int32_t accumulate_grad, const float* upstream, void* stream);

/* Geometry terms of the first-stage loss (/root/reference/train.py:323-328):
 *   normal_weight * F.l1_loss(normal_map[:, mask], normal_map_from_depth[:, mask])
 * + tv_weight     * get_tv_loss(gt_image, normal_map, pad=1, step=1)       (train.py:83-100, edge-aware total variation)
 * and the gradient with respect to normal_map (normal_map_from_depth comes from depth_to_normal, which the reference
 * runs outside autograd). All maps are [3,H,W]; mask is uint8 [H,W] (NULL = every pixel). An empty mask gives NaN,
 * like the reference's mean over an empty selection. loss_out (float[3], may be NULL) = [loss_scale * total (added to
 * the existing value when accumulate_loss), normal L1 term, TV term]; grad_normal (may be NULL) as in gigs_image_loss. */
int gigs_normal_loss(int32_t W, int32_t H, const float* normal_map, const float* normal_from_depth, const uint8_t* mask,
                     const float* gt_image, float normal_weight, float tv_weight, float loss_scale, void* scratch,
                     uint64_t* scratch_bytes, float* loss_out, int32_t accumulate_loss, float* grad_normal,
                     int32_t accumulate_grad, const float* upstream, void* stream);

/* ---- Densification / pruning rebuild ------------------------------------------------------------------
 * Replaces the tensor surgery of GaussianModel.densify_and_prune (/root/reference/scene/gaussian_model.py:905-931 with
 * densify_and_clone :785-817, densify_and_split :741-783, cat_tensors_to_optimizer :639-662, _prune_optimizer
 * :594-612): given the source map of the surviving rows — src_index[n_out] (row of the old model), kind[n_out] (0 kept
 * point, 1 clone, 2 split child) — it writes every parameter tensor and both Adam moment tensors of the new model in
 * one launch. Kept rows copy parameter and moments; new rows copy the parameter and get zero moments; in the group
 * with role 1 (xyz, width 3) a new row is R(rot) * (noise * exp(log_scale)) + xyz of its source; in the group with
 * role 2 (log-scale, width 3) a split child gets log(exp(log_scale) / split_div), split_div = 0.8 * N.
 * noise: [n_out,3] standard normals (rows of kept points unused). The group array is HOST memory (<= 16 groups);
 * moment pointers may be NULL (no optimiser state yet). */
typedef struct GigsDensifyGroup {
    const float* src; const float* src_exp_avg; const float* src_exp_avg_sq;
    float* dst; float* dst_exp_avg; float* dst_exp_avg_sq;
    int32_t width;   /* floats per Gaussian */
    int32_t role;    /* 0 copy, 1 xyz, 2 log-scale */
} GigsDensifyGroup;
int gigs_densify_gather(int32_t n_out, const int32_t* src_index, const int8_t* kind, const float* noise,
                        const float* src_log_scale, const float* src_rot, float split_div, int32_t n_groups,
                        const GigsDensifyGroup* groups, void* stream);

/* ---- Gradient all-reduce over NVLink peer memory ------------------------------------------------------
 * The exchange step of the view-sharded training step (the reference is single-GPU; /root/repo/BASELINE.json: "per-
 * Gaussian gradient allreduce over NVLink") as one kernel per rank instead of a sequence of library collectives: a
 * cross-rank barrier through flag words in peer memory, an in-place two-shot all-reduce (rank r sums slice r of every
 * span over all ranks' buffers in rank order and stores it into all of them), a second barrier.
 * peer_bufs[world] / peer_flags[world] (HOST arrays of device addresses): every rank's gradient buffer and flag block as
 * mapped into THIS process (symmetric allocations; entry `rank` is the local one). A flag block is 2*world + 2 uint32,
 * zero before the first call; calls on one flag block must be stream-ordered. multicast_buf: the NVLS multicast mapping
 * of the same buffer (0 if the fabric offers none): 16-byte-aligned spans are then reduced by the switch
 * (multimem.ld_reduce / multimem.st) instead of N peer loads and N peer stores per element. epoch = 1, 2, 3, ... must
 * advance by one per call and agree on all ranks, as must the spans (float offsets [begin, end) into the buffer).
 * n_ctas: 0 = scaled with the bytes (8 .. 128). Without multicast the sums are formed in rank order: bit-identical on
 * all ranks and independent of timing; with it every rank still receives the same bits (one rank reduces a slice and
 * broadcasts it). A peer that has not arrived after about a minute of GPU clock is fatal: the error word
 * flags[2*world + 1] = epoch is set and the kernel traps (the process's next CUDA call fails) instead of the ranks
 * training on with unreduced gradients. */
int gigs_peer_allreduce(int32_t world, int32_t rank, const uint64_t* peer_bufs, const uint64_t* peer_flags,
                        uint64_t multicast_buf, uint32_t epoch, int32_t n_spans, const uint64_t* span_begin,
                        const uint64_t* span_end, int32_t n_ctas, void* stream);

/* Replaces distCUDA2 / SimpleKNN::knn (/root/reference/submodules/simple-knn/spatial.cu,
 * simple_knn.cu:165-207): mean squared distance to the 3 nearest other points.
 * scratch_bytes: call with scratch==NULL to query. */
int gigs_dist2(int32_t P, const float* points, float* mean_dist2, void* scratch,
               uint64_t* scratch_bytes, void* stream);

/* Per-stage device timing (CUDA events on the launching stream), for bench.py's roofline numbers.
 * Stage ids: 0 preprocess+scan, 1 emit_keys, 2 radix_sort, 3 tile_ranges, 4 blend_forward, 5 blend_backward,
 * 6 gaussian_backward, 7 geometry_chain, 8 ssao, 9 ssr, 10 shade_forward, 11 shade_backward, 12 median3x3,
 * 13 median3x3_backward, 14 bilateral3x3, 15 depth_to_normal, 16 ssr_backward, 17 dist2, 18 deferred_shade,
 * 19 deferred_loss, 20 deferred_backward, 21 param_grad, 22 one radix-sort pass (nested inside 2 / 23),
 * 23 depth argsort of the Gaussians, 24 cubemap prefilter forward, 25 cubemap prefilter backward, 26 Adam step,
 * 27 image loss (L1 + SSIM forward, finish and backward), 28 normal loss (L1 + TV forward, finish, backward),
 * 29 first-stage normal post-processing, 30 its backward, 31 deferred_backward_kernel alone (nested inside 20),
 * 32 the peer all-reduce kernel.
 * gigs_profile_read synchronises the recorded events, writes up to cap (stage, ms) pairs, clears the log and
 * returns the number written (negative on error). Off by default; costs two event records per stage when on. */
int gigs_profile_enable(int32_t on);
int gigs_profile_read(int32_t* stages, float* ms, int32_t cap);

/* Measured-roofline helper: register-only FFMA throughput (flop/s) over all SMs. */
int gigs_ffma_peak(double* tflops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GIGS_B200_H */
