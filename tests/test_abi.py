"""The C-ABI library loads on a CPU-only box and exports every symbol include/gigs_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gigs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gigs_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ("gigs_raster_forward_begin", "gigs_raster_forward_finish", "gigs_raster_backward", "gigs_ssao",
                 "gigs_ssr", "gigs_shade_forward", "gigs_shade_backward", "gigs_dist2", "gigs_geometry_chain",
                 "gigs_mark_visible", "gigs_depth_to_normal", "gigs_lite_forward_finish", "gigs_ssr_backward",
                 "gigs_frame_forward", "gigs_frame_backward", "gigs_frame_layout", "gigs_sizeof"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from gigs import _lib
    assert os.path.exists(_lib.LIB_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/gigs_b200.h but not exported by libgigs_b200.so"


def test_python_binding_covers_every_declared_symbol():
    from gigs import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()
    lib = _lib.load()
    assert lib.gigs_abi_version() == 4


def test_workspace_sizes_are_a_pure_function_of_shape():
    from gigs import _lib
    L = _lib.load()
    a, b = _lib.GigsSizes(), _lib.GigsSizes()
    assert L.gigs_raster_sizes(1000, 800, 800, 12345, ctypes.byref(a)) == 0
    assert L.gigs_raster_sizes(1000, 800, 800, 12345, ctypes.byref(b)) == 0
    assert (a.geom_bytes, a.img_bytes, a.binning_bytes, a.sort_bytes) == (b.geom_bytes, b.img_bytes, b.binning_bytes,
                                                                          b.sort_bytes)
    assert a.geom_bytes % 128 == 0 and a.img_bytes % 128 == 0 and a.binning_bytes >= 4 * 12345
    lay = _lib.GigsLayout()
    assert L.gigs_raster_layout(1000, 800, 800, 12345, ctypes.byref(lay)) == 0
    for f, _ in lay._fields_:
        assert getattr(lay, f) % 128 == 0, f
    # bad arguments -> negative status + message, never a crash
    assert L.gigs_raster_sizes(-1, 800, 800, 0, ctypes.byref(a)) < 0
    assert b"bad arguments" in L.gigs_last_error()


def test_missing_library_is_a_loud_import_error(tmp_path, monkeypatch):
    from gigs import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _lib.load()
    except ImportError as e:
        assert "no CPU or PyTorch fallback" in str(e)
    else:
        raise AssertionError("expected ImportError")


def test_struct_mirrors_match_the_library():
    """ctypes mirrors of the argument structs have the C sizes (checked again at load time)."""
    from gigs import _lib
    L = _lib.load()
    for which, st in enumerate((_lib.GigsCamera, _lib.GigsSizes, _lib.GigsLayout, _lib.GigsRasterFwd,
                                _lib.GigsRasterBwd, _lib.GigsShade, _lib.GigsFrameLayout, _lib.GigsFrame)):
        assert L.gigs_sizeof(which) == ctypes.sizeof(st), st.__name__
    assert L.gigs_sizeof(99) < 0


def test_frame_layout_is_aligned_disjoint_and_shape_determined():
    from gigs import _lib
    L = _lib.load()
    a, b = _lib.GigsFrameLayout(), _lib.GigsFrameLayout()
    assert L.gigs_frame_layout(800, 800, ctypes.byref(a)) == 0
    assert L.gigs_frame_layout(800, 800, ctypes.byref(b)) == 0
    offs = [getattr(a, f) for f, _ in a._fields_]
    assert offs == [getattr(b, f) for f, _ in b._fields_]
    fields = offs[:-1]
    assert all(o % 256 == 0 for o in fields) and fields == sorted(fields) and len(set(fields)) == len(fields)
    assert a.total_bytes > fields[-1]
    # 62 float planes + 4 byte planes at 800x800, the 9.4 MB private-texel scratch, plus small tails
    assert 62 * 4 * 640000 + 4 * 640000 <= a.total_bytes <= 62 * 4 * 640000 + 4 * 640000 + (11 << 20)
    assert L.gigs_frame_layout(0, 10, ctypes.byref(a)) < 0
    # frame entry points validate their arguments without touching the GPU
    f = _lib.GigsFrame()
    assert L.gigs_frame_forward(ctypes.byref(f)) < 0 and L.gigs_frame_backward(ctypes.byref(f)) < 0
    assert L.gigs_frame_forward(None) < 0


def test_binning_blob_has_room_for_the_forward_footprint_masks():
    """The binning blob holds the sorted ids (4 R bytes) and, for the backward kernels, the forward's footprint-test
    ballots: 8 words per 32-entry chunk, chunk index of tile t = range.x/32 + t + chunk in tile, which needs at most
    R/32 + T + 1 chunks whatever the tile lengths are (csrc/common.cuh: Layout.b_warp_masks)."""
    from gigs import _lib
    L = _lib.load()
    s = _lib.GigsSizes()
    for (P, W, H, R) in [(1000, 800, 800, 0), (1000, 800, 800, 12345), (300000, 800, 800, 4227388),
                         (6000000, 1237, 822, 31991500), (10, 16, 16, 1), (10, 3840, 2160, 77)]:
        assert L.gigs_raster_sizes(P, W, H, R, ctypes.byref(s)) == 0
        T = ((W + 15) // 16) * ((H + 15) // 16)
        assert s.binning_bytes >= 4 * R + 32 * (R // 32 + T + 1), (P, W, H, R, s.binning_bytes)


def test_dependent_launch_switch_and_launch_counter_are_host_side():
    """gigs_set_dependent_launch / gigs_launch_count touch no device: usable (and consistent) on a CPU-only box."""
    from gigs import _lib
    L = _lib.load()
    prev = L.gigs_set_dependent_launch(-1)
    assert prev in (0, 1)
    assert L.gigs_set_dependent_launch(0) == prev and L.gigs_set_dependent_launch(-1) == 0
    assert L.gigs_set_dependent_launch(1) == 0 and L.gigs_set_dependent_launch(-1) == 1
    L.gigs_set_dependent_launch(prev)
    assert int(L.gigs_launch_count()) >= 0
