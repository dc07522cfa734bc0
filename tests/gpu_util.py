"""Helpers shared by the -m gpu tests (test infrastructure)."""
import torch

import diff_gaussian_rasterization as dgr

E = torch.Tensor([])


def ours_forward(g, cam, bg, deg=3, inference=False, argmax=False, colors_precomp=None, cov3D_precomp=None,
                 scale_modifier=1.0, debug=False, prefiltered=False):
    W, H = cam.image_width, cam.image_height
    shs = E if colors_precomp is not None else g["shs"]
    cp = colors_precomp if colors_precomp is not None else E
    sc = E if cov3D_precomp is not None else g["scales"]
    rt = E if cov3D_precomp is not None else g["rotations"]
    cv = cov3D_precomp if cov3D_precomp is not None else E
    res = dgr._C.rasterize_gaussians(bg, g["means3D"], cp, g["opacity"], g["normal"], g["albedo"], g["roughness"],
                                     g["metallic"], sc, rt, cv, shs, cam.camera_center, cam.world_view_transform,
                                     cam.full_proj_transform, scale_modifier, cam.tanfovx, cam.tanfovy, H, W, deg,
                                     prefiltered, argmax, inference, debug)
    names = ("num_rendered", "color", "radii", "geom", "binning", "img", "opacity", "depth", "normal", "normal_view",
             "pos", "albedo", "roughness", "metallic")
    return dict(zip(names, res))


def ours_backward(g, cam, bg, fwd, grads, deg=3, colors_precomp=None, cov3D_precomp=None, scale_modifier=1.0):
    W, H = cam.image_width, cam.image_height
    shs = E if colors_precomp is not None else g["shs"]
    cp = colors_precomp if colors_precomp is not None else E
    sc = E if cov3D_precomp is not None else g["scales"]
    rt = E if cov3D_precomp is not None else g["rotations"]
    cv = cov3D_precomp if cov3D_precomp is not None else E
    out = dgr._C.rasterize_gaussians_backward(
        bg, g["means3D"], fwd["radii"], cp, g["normal"], g["albedo"], g["roughness"], g["metallic"], sc, rt, cv, shs,
        cam.camera_center, cam.world_view_transform, cam.full_proj_transform, scale_modifier, cam.tanfovx, cam.tanfovy,
        deg, grads.get("depth"), grads.get("color"), grads.get("opacity"), grads.get("normal"), grads.get("albedo"),
        grads.get("roughness"), grads.get("metallic"), fwd["geom"], fwd["binning"], fwd["img"], fwd["num_rendered"],
        False, image_height=H, image_width=W)
    names = ("means2D", "colors", "opacity", "normal", "albedo", "roughness", "metallic", "means3D", "cov3D", "sh",
             "scales", "rotations")
    return dict(zip(names, out))


def decode_state(fwd, P, W, H):
    """Our workspace blobs -> the same named arrays refshim.RefRasterizer.state() returns.

    Our binning never materialises the reference's 64-bit keys (depth argsort of the Gaussians + stable
    sort of (tile, id) pairs by tile id), so they are REBUILT here from what it does produce:
      keys_sorted[i]  = tile_of_sorted_slot[i] << 32 | depth_bits(point_list[i])
      keys_unsorted / vals_unsorted = our emitted (tile, id) pairs put back into the reference's emission order
                        (ascending Gaussian id, row-major tiles inside a Gaussian's rect) with the depth bits attached
      point_offsets   = inclusive scan of tiles_touched (rasterizer_impl.cu:585)
    which makes the comparisons against the reference's arrays bit-for-bit checks of our pair multiset, of every
    Gaussian's tile rect, of the final order and of the ranges."""
    R = fwd["num_rendered"]
    lay = dgr.raster_layout(P, W, H, R)
    sc = dgr.sort_scratch(fwd["color"].device)
    geom, img, binning = fwd["geom"], fwd["img"], fwd["binning"]

    def view(buf, off, nbytes, dtype, shape):
        return buf[off:off + nbytes].view(dtype).reshape(shape)

    rec = view(geom, lay.g_record, P * 96, torch.float32, (P, 24))
    T = ((W + 15) // 16) * ((H + 15) // 16)
    N = W * H
    depth_bits = rec[:, 6].contiguous().view(torch.int32).long() & 0xFFFFFFFF
    tiles_touched = view(geom, lay.g_tiles_touched, 4 * P, torch.int32, (P,))
    tiles_sorted = view(sc, lay.s_tiles_sorted, 4 * R, torch.int32, (R,)).long()
    tiles_unsorted = view(sc, lay.s_tiles_unsorted, 4 * R, torch.int32, (R,)).long()
    vals_emitted = view(sc, lay.s_vals_unsorted, 4 * R, torch.int32, (R,))
    point_list = view(binning, lay.b_point_list, 4 * R, torch.int32, (R,))
    ref_order = torch.sort(vals_emitted.long() * (1 << 20) + tiles_unsorted, stable=True).indices
    return dict(
        record=rec, means2D=rec[:, 0:2], conic=rec[:, 2:5], depths=rec[:, 6], rgb=rec[:, 8:11],
        cov3D=view(geom, lay.g_cov3D, 24 * P, torch.float32, (P, 6)),
        clamped=view(geom, lay.g_clamped, 4 * P, torch.uint8, (P, 4))[:, :3],
        tiles_touched=tiles_touched,
        point_offsets=torch.cumsum(tiles_touched.long(), 0).int(),
        depth_keys=view(geom, lay.g_depth_keys, 4 * P, torch.int32, (P,)),
        order=view(geom, lay.g_order, 4 * P, torch.int32, (P,)),
        tiles_emitted=tiles_unsorted, vals_emitted=vals_emitted, tiles_sorted=tiles_sorted,
        keys_unsorted=(tiles_unsorted[ref_order] << 32) | depth_bits[vals_emitted.long()[ref_order]],
        vals_unsorted=vals_emitted[ref_order],
        keys_sorted=(tiles_sorted << 32) | depth_bits[point_list.long()],
        point_list=point_list,
        final_T=view(img, lay.i_final_T, 4 * N, torch.float32, (N,)),
        n_contrib=view(img, lay.i_n_contrib, 4 * N, torch.int32, (N,)),
        ranges=view(img, lay.i_ranges, 8 * T, torch.int32, (T, 2)))


def assert_close_map(a, b, tol, name):
    a, b = a.float(), b.float()
    na, nb = torch.isnan(a), torch.isnan(b)
    assert torch.equal(na, nb), f"{name}: NaN positions differ"
    ok = ~na
    d = (a[ok] - b[ok]).abs().max().item() if ok.any() else 0.0
    assert d <= tol, f"{name}: max abs diff {d} > {tol}"


def assert_grad_close(a, b, name, rel_tol=1e-3):
    """norm-wise relative error per tensor, with a per-element absolute floor (north_star: 1e-3 relative)."""
    a, b = a.float().reshape(-1), b.float().reshape(-1)
    floor = 1e-9 * (a.numel() ** 0.5)
    rel = ((a - b).norm() / (b.norm() + floor)).item()
    assert rel <= rel_tol, f"{name}: relative error {rel} > {rel_tol}"
    mx = (a - b).abs().max().item()
    assert mx <= rel_tol * b.abs().max().item() + 1e-8, f"{name}: element error {mx}"
