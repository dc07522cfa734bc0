"""ctypes binding of the C-ABI in include/gigs_b200.h.

There is NO fallback: if libgigs_b200.so is missing or a symbol is absent the import fails loudly
(build with `python -c "import __graft_entry__ as g; g.build()"` or `make -C gi-gs_b200/csrc`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GIGS_LIB: an alternative build of the same library (kernel-tuning experiments: tools/build_variant.sh)
LIB_PATH = os.environ.get("GIGS_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libgigs_b200.so")

c_f32p = C.c_void_p  # device pointers travel as integers


class GigsCamera(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32),
        ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
        ("scale_modifier", C.c_float),
        ("sh_degree", C.c_int32), ("sh_coeffs", C.c_int32),
        ("prefiltered", C.c_int32), ("debug", C.c_int32), ("inference", C.c_int32), ("argmax_depth", C.c_int32),
        ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p), ("bg", C.c_void_p),
    ]


class GigsSizes(C.Structure):
    _fields_ = [("geom_bytes", C.c_uint64), ("img_bytes", C.c_uint64), ("binning_bytes", C.c_uint64),
                ("sort_bytes", C.c_uint64)]


class GigsLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "g_record", "g_cov3D", "g_clamped", "g_tiles_touched", "g_depth_keys", "g_order", "g_block_sums",
        "g_num_rendered", "i_final_T", "i_n_contrib", "i_ranges", "b_point_list", "s_tiles_sorted",
        "s_tiles_unsorted", "s_vals_unsorted")]


class GigsRasterFwd(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("material_only", C.c_int32),
        ("cam", GigsCamera),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p),
        ("normal", C.c_void_p), ("albedo", C.c_void_p), ("roughness", C.c_void_p), ("metallic", C.c_void_p),
        ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("out_color", C.c_void_p), ("out_opacity", C.c_void_p), ("out_depth", C.c_void_p), ("out_normal", C.c_void_p),
        ("out_normal_view", C.c_void_p), ("out_pos", C.c_void_p), ("out_albedo", C.c_void_p),
        ("out_roughness", C.c_void_p), ("out_metallic", C.c_void_p),
        ("radii", C.c_void_p),
        ("geom", C.c_void_p), ("geom_bytes", C.c_uint64),
        ("img", C.c_void_p), ("img_bytes", C.c_uint64),
        ("binning", C.c_void_p), ("binning_bytes", C.c_uint64),
        ("sort", C.c_void_p), ("sort_bytes", C.c_uint64),
        ("pinned_num_rendered", C.c_void_p),
        ("num_rendered", C.c_int64),
        ("stream", C.c_void_p),
    ]


class GigsRasterBwd(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("_pad", C.c_int32),
        ("num_rendered", C.c_int64),
        ("cam", GigsCamera),
        ("means3D", C.c_void_p), ("shs", C.c_void_p), ("colors_precomp", C.c_void_p),
        ("normal", C.c_void_p), ("albedo", C.c_void_p), ("roughness", C.c_void_p), ("metallic", C.c_void_p),
        ("scales", C.c_void_p), ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
        ("radii", C.c_void_p),
        ("geom", C.c_void_p), ("binning", C.c_void_p), ("img", C.c_void_p),
        ("dL_dpix_depth", C.c_void_p), ("dL_dpix", C.c_void_p), ("dL_dpix_opacity", C.c_void_p),
        ("dL_dpix_normal", C.c_void_p), ("dL_dpix_albedo", C.c_void_p), ("dL_dpix_roughness", C.c_void_p),
        ("dL_dpix_metallic", C.c_void_p),
        ("accum", C.c_void_p),
        ("dL_dmean2D", C.c_void_p), ("dL_dconic", C.c_void_p), ("dL_dopacity", C.c_void_p), ("dL_dcolor", C.c_void_p),
        ("dL_dnormal", C.c_void_p), ("dL_dalbedo", C.c_void_p), ("dL_droughness", C.c_void_p),
        ("dL_dmetallic", C.c_void_p), ("dL_dmean3D", C.c_void_p), ("dL_dcov3D", C.c_void_p), ("dL_dsh", C.c_void_p),
        ("dL_dscale", C.c_void_p), ("dL_drot", C.c_void_p),
        ("stream", C.c_void_p),
    ]


class GigsShade(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("H", C.c_int32),
        ("n_spec_levels", C.c_int32),
        ("spec_res", C.c_int32 * 8),
        ("spec", C.c_void_p * 8),
        ("diffuse_res", C.c_int32),
        ("diffuse", C.c_void_p),
        ("brdf_lut", C.c_void_p),
        ("lut_res", C.c_int32),
        ("tone", C.c_int32), ("gamma", C.c_int32), ("has_metallic", C.c_int32), ("has_occlusion", C.c_int32),
        ("min_roughness", C.c_float), ("max_roughness", C.c_float),
        ("normals", C.c_void_p), ("view_dirs", C.c_void_p), ("albedo", C.c_void_p), ("roughness", C.c_void_p),
        ("metallic", C.c_void_p), ("occlusion", C.c_void_p), ("mask", C.c_void_p), ("background", C.c_void_p),
        ("render_rgb", C.c_void_p), ("diffuse_rgb", C.c_void_p), ("specular_rgb", C.c_void_p),
        ("diffuse_light", C.c_void_p),
        ("g_render_rgb", C.c_void_p), ("g_diffuse_rgb", C.c_void_p), ("g_specular_rgb", C.c_void_p),
        ("g_albedo", C.c_void_p), ("g_roughness", C.c_void_p), ("g_metallic", C.c_void_p),
        ("g_diffuse_tex", C.c_void_p),
        ("g_spec", C.c_void_p * 8),
        ("stream", C.c_void_p),
    ]


class GigsFrameLayout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "color", "opacity", "depth", "normal", "normal_view", "pos", "albedo", "roughness", "metallic",
        "normal_from_depth", "depth_pos", "occlusion", "shade_normal", "ssr_normal", "render_direct", "linear_rgb",
        "F0", "rough_remap", "metal_used", "ssr_color", "ssr_abd", "render_rgb", "g_rgb", "g_albedo", "g_roughness",
        "g_metallic", "mask", "median_sel", "tex_scratch", "partials", "stats", "tv_edge", "total_bytes")]


class GigsLightLayout(C.Structure):
    _fields_ = ([("n_levels", C.c_int32), ("res", C.c_int32 * 8), ("roughness", C.c_float * 8),
                 ("cutoff", C.c_float * 8), ("pad_", C.c_int32)]
                + [(n, C.c_uint64 * 8) for n in ("table", "bounds", "chain", "spec", "wsum", "gq", "g_chain", "g_spec")]
                + [(n, C.c_uint64 * 9) for n in ("rowptr", "wptr")]
                + [(n, C.c_uint64) for n in ("counts", "totals", "diffuse", "gq_diffuse", "g_diffuse_in", "g_diffuse",
                                             "grad_begin", "grad_bytes", "total_bytes")]
                + [("lanes_log2", C.c_int32 * 9), ("pad2_", C.c_int32)]
                + [(n, C.c_uint64 * 9) for n in ("n_runs", "n_weights", "w_rows", "w_fwd", "w_bwd")]
                + [("weights_bytes", C.c_uint64)])


class GigsFrame(C.Structure):
    _fields_ = [
        ("P", C.c_int32), ("raw_params", C.c_int32),
        ("cam", GigsCamera),
        ("means3D", C.c_void_p), ("sh_dc", C.c_void_p), ("sh_rest", C.c_void_p), ("opacities", C.c_void_p),
        ("normal", C.c_void_p), ("albedo", C.c_void_p), ("roughness", C.c_void_p), ("metallic", C.c_void_p),
        ("scales", C.c_void_p), ("rotations", C.c_void_p),
        ("radius", C.c_float), ("bias", C.c_float), ("thick", C.c_float), ("delta", C.c_float),
        ("step", C.c_int32), ("start", C.c_int32),
        ("indirect", C.c_int32), ("use_metallic", C.c_int32), ("tone", C.c_int32), ("gamma", C.c_int32),
        ("n_spec_levels", C.c_int32), ("spec_res", C.c_int32 * 8), ("spec", C.c_void_p * 8),
        ("diffuse_res", C.c_int32), ("diffuse", C.c_void_p), ("brdf_lut", C.c_void_p), ("lut_res", C.c_int32),
        ("min_roughness", C.c_float), ("max_roughness", C.c_float),
        ("canonical_rays", C.c_void_p), ("gt_image", C.c_void_p),
        ("loss_scale", C.c_float), ("lamb_weight", C.c_float), ("brdf_tv_weight", C.c_float), ("material_only", C.c_int32),
        ("geom", C.c_void_p), ("geom_bytes", C.c_uint64), ("img", C.c_void_p), ("img_bytes", C.c_uint64),
        ("binning", C.c_void_p), ("binning_bytes", C.c_uint64), ("sort", C.c_void_p), ("sort_bytes", C.c_uint64),
        ("maps", C.c_void_p), ("maps_bytes", C.c_uint64),
        ("radii", C.c_void_p), ("accum", C.c_void_p), ("pinned_num_rendered", C.c_void_p),
        ("num_rendered", C.c_int64), ("resume", C.c_int32), ("skip_geometry", C.c_int32),
        ("need_binning_bytes", C.c_uint64), ("need_sort_bytes", C.c_uint64),
        ("g_albedo", C.c_void_p), ("g_roughness", C.c_void_p), ("g_metallic", C.c_void_p),
        ("g_diffuse_tex", C.c_void_p), ("g_spec", C.c_void_p * 8),
        ("gt_ready_event", C.c_void_p),
        ("light_ready_event", C.c_void_p),
        ("stream", C.c_void_p),
    ]


class GigsAdamGroup(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("count", C.c_uint64), ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double),
                ("eps", C.c_double), ("step", C.c_int32), ("clamp_min0", C.c_int32), ("clear_grad", C.c_int32),
                ("pad_", C.c_int32)]


class GigsDensifyGroup(C.Structure):
    _fields_ = [("src", C.c_void_p), ("src_exp_avg", C.c_void_p), ("src_exp_avg_sq", C.c_void_p),
                ("dst", C.c_void_p), ("dst_exp_avg", C.c_void_p), ("dst_exp_avg_sq", C.c_void_p),
                ("width", C.c_int32), ("role", C.c_int32)]


class GigsStage1Layout(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "color", "opacity", "depth", "normal", "normal_view", "pos", "albedo", "roughness", "metallic",
        "normal_from_depth", "depth_pos", "normals_view", "nfd_unit", "g_color", "g_normals_view", "g_normal",
        "median_sel", "mask", "loss_scratch", "loss_scratch_bytes", "nloss_scratch", "nloss_scratch_bytes", "stats",
        "total_bytes")]


class GigsStage1(C.Structure):
    _fields_ = ([("P", C.c_int32), ("pad0_", C.c_int32), ("cam", GigsCamera)]
                + [(n, C.c_void_p) for n in ("xyz", "f_dc", "f_rest", "opacity", "normal", "albedo", "roughness",
                                             "metallic", "log_scale", "rot", "gt_image")]
                + [(n, C.c_float) for n in ("lambda_dssim", "normal_weight", "normal_tv_weight", "loss_scale")]
                + [("geom", C.c_void_p), ("geom_bytes", C.c_uint64), ("img", C.c_void_p), ("img_bytes", C.c_uint64),
                   ("binning", C.c_void_p), ("binning_bytes", C.c_uint64), ("sort", C.c_void_p), ("sort_bytes", C.c_uint64),
                   ("maps", C.c_void_p), ("maps_bytes", C.c_uint64),
                   ("radii", C.c_void_p), ("accum", C.c_void_p), ("pinned_num_rendered", C.c_void_p),
                   ("num_rendered", C.c_int64), ("resume", C.c_int32), ("pad1_", C.c_int32),
                   ("need_binning_bytes", C.c_uint64), ("need_sort_bytes", C.c_uint64)]
                + [(n, C.c_void_p) for n in ("g_xyz", "g_f_dc", "g_f_rest", "g_opacity", "g_normal", "g_albedo",
                                             "g_roughness", "g_metallic", "g_log_scale", "g_rot", "g_means2D",
                                             "gt_ready_event", "stream")])


GIGS_E_GROW = -5

# every symbol include/gigs_b200.h declares: (name, restype, argtypes)
_i32, _f, _vp, _u64 = C.c_int32, C.c_float, C.c_void_p, C.c_uint64
SYMBOLS = {
    "gigs_abi_version": (C.c_int, []),
    "gigs_last_error": (C.c_char_p, []),
    "gigs_sizeof": (C.c_int, [_i32]),
    "gigs_raster_sizes": (C.c_int, [_i32, _i32, _i32, _u64, C.POINTER(GigsSizes)]),
    "gigs_raster_layout": (C.c_int, [_i32, _i32, _i32, _u64, C.POINTER(GigsLayout)]),
    "gigs_raster_forward_begin": (C.c_int, [C.POINTER(GigsRasterFwd)]),
    "gigs_raster_forward_finish": (C.c_int, [C.POINTER(GigsRasterFwd)]),
    "gigs_lite_forward_finish": (C.c_int, [C.POINTER(GigsRasterFwd)]),
    "gigs_raster_backward": (C.c_int, [C.POINTER(GigsRasterBwd)]),
    "gigs_mark_visible": (C.c_int, [_i32, _vp, _vp, _vp, _vp]),
    "gigs_depth_to_normal": (C.c_int, [_i32, _i32, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "gigs_median3x3": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp]),
    "gigs_median3x3_backward": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "gigs_bilateral3x3": (C.c_int, [_i32, _i32, _i32, _f, _f, _vp, _vp, _vp]),
    "gigs_geometry_chain": (C.c_int, [_i32, _i32, _f, _f, _vp, _vp, _i32, _vp, _vp, _vp]),
    "gigs_gi_scratch_bytes": (C.c_uint64, [_i32, _i32]),
    "gigs_ssao": (C.c_int, [_i32, _i32, _f, _f, _f, _f, _f, _f, _i32, _i32, _vp, _vp, _vp, _vp, _u64, _vp]),
    "gigs_ssr": (C.c_int, [_i32, _i32, _f, _f, _f, _f, _f, _f, _i32, _i32] + [_vp] * 10 + [_u64, _vp]),
    "gigs_gi_count_probes": (C.c_int, [_i32, _i32, _f, _f, _f, _f, _f, _f, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "gigs_gi_tune": (C.c_int, [_i32, _i32]),
    "gigs_set_dependent_launch": (C.c_int, [_i32]),
    "gigs_cube_wrap_selfcheck": (C.c_int, [_i32, _vp, _vp]),
    "gigs_launch_count": (C.c_uint64, []),
    "gigs_ssr_backward": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gigs_shade_forward": (C.c_int, [C.POINTER(GigsShade)]),
    "gigs_shade_backward": (C.c_int, [C.POINTER(GigsShade)]),
    "gigs_frame_layout": (C.c_int, [_i32, _i32, C.POINTER(GigsFrameLayout)]),
    "gigs_frame_forward": (C.c_int, [C.POINTER(GigsFrame)]),
    "gigs_frame_backward": (C.c_int, [C.POINTER(GigsFrame)]),
    "gigs_latlong_to_cubemap": (C.c_int, [_i32, _i32, _i32, _vp, _i32, _vp, _vp]),
    "gigs_cubemap_table": (C.c_int, [_i32, _vp, _vp]),
    "gigs_specular_bounds": (C.c_int, [_i32, _f, _vp, _vp, _vp]),
    "gigs_cubemap_mip_forward": (C.c_int, [_i32, _vp, _vp, _vp]),
    "gigs_cubemap_mip_backward": (C.c_int, [_i32, _vp, _vp, _i32, _vp]),
    "gigs_diffuse_cubemap_forward": (C.c_int, [_i32, _vp, _vp, _vp, _vp]),
    "gigs_diffuse_cubemap_backward": (C.c_int, [_i32, _vp, _vp, _vp, _vp]),
    "gigs_specular_cubemap_forward": (C.c_int, [_i32, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "gigs_specular_cubemap_backward": (C.c_int, [_i32, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "gigs_light_layout": (C.c_int, [_i32, _i32, C.POINTER(GigsLightLayout)]),
    "gigs_light_prepare": (C.c_int, [C.POINTER(GigsLightLayout), _vp, _vp]),
    "gigs_light_weights": (C.c_int, [C.POINTER(GigsLightLayout), _vp, _vp, _vp]),
    "gigs_light_build": (C.c_int, [C.POINTER(GigsLightLayout), _vp, _vp, _vp, _vp]),
    "gigs_light_backward": (C.c_int, [C.POINTER(GigsLightLayout), _vp, _vp, _vp, _i32, _i32, _vp]),
    "gigs_env_tv": (C.c_int, [_i32, _vp, _vp, _i32, _i32, _f, _vp, C.POINTER(C.c_uint64), _vp, _vp, _i32, _vp]),
    "gigs_stage1_layout": (C.c_int, [_i32, _i32, C.POINTER(GigsStage1Layout)]),
    "gigs_stage1_forward": (C.c_int, [C.POINTER(GigsStage1)]),
    "gigs_stage1_backward": (C.c_int, [C.POINTER(GigsStage1)]),
    "gigs_adam_step": (C.c_int, [_i32, C.POINTER(GigsAdamGroup), _vp]),
    "gigs_clear_spans": (C.c_int, [_vp, _i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _vp]),
    "gigs_densify_stats": (C.c_int, [_i32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gigs_image_loss": (C.c_int, [_i32, _i32, _i32, _vp, _vp, _f, _f, _vp, C.POINTER(C.c_uint64), _vp, _i32, _vp, _i32,
                                  _vp, _vp]),
    "gigs_normal_loss": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, C.POINTER(C.c_uint64), _vp, _i32, _vp,
                                   _i32, _vp, _vp]),
    "gigs_densify_gather": (C.c_int, [_i32, _vp, _vp, _vp, _vp, _vp, _f, _i32, C.POINTER(GigsDensifyGroup), _vp]),
    "gigs_peer_allreduce": (C.c_int, [_i32, _i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _u64, C.c_uint32, _i32,
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _i32, _vp]),
    "gigs_dist2": (C.c_int, [_i32, _vp, _vp, _vp, C.POINTER(C.c_uint64), _vp]),
    "gigs_ffma_peak": (C.c_int, [C.POINTER(C.c_double), _vp]),
    "gigs_profile_enable": (C.c_int, [_i32]),
    "gigs_profile_read": (C.c_int, [C.POINTER(C.c_int32), C.POINTER(C.c_float), _i32]),
}

_lib = None


def load():
    """Load libgigs_b200.so and bind every declared symbol. Raises (never falls back) on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"gigs_b200: {LIB_PATH} not found. This package has no CPU or PyTorch fallback; build the CUDA "
            f"library first (python -c 'import __graft_entry__ as g; g.build()').")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.gigs_abi_version() != 4:
        raise ImportError("gigs_b200: ABI version mismatch between the python binding and libgigs_b200.so")
    for which, st in enumerate((GigsCamera, GigsSizes, GigsLayout, GigsRasterFwd, GigsRasterBwd, GigsShade,
                                GigsFrameLayout, GigsFrame, GigsLightLayout, GigsAdamGroup,
                                GigsDensifyGroup, GigsStage1Layout, GigsStage1)):
        if lib.gigs_sizeof(which) != C.sizeof(st):
            raise ImportError(f"gigs_b200: struct {st.__name__} is {C.sizeof(st)} bytes in the python binding but "
                              f"{lib.gigs_sizeof(which)} in libgigs_b200.so")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        msg = load().gigs_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed with status {status}: {msg}")


def ptr(t):
    """Device pointer of a tensor, or None (-> NULL) for None / empty tensors, as the reference's
    C++ sees nullptr for torch.Tensor([]) (diff_gaussian_rasterization/__init__.py:435-445)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()
