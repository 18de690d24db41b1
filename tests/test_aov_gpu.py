"""Depth-0 feature planes for a denoiser (setting "aov", rfwb200_read_aov): the albedo / normal sums the reference's OptiX
backend keeps when built with ALLOW_DENOISER (OptiX6Context/assets/kernels/kernels.cu:122-133,206-221,316-330)."""
import numpy as np
import pytest

import rfwb200 as R
import scenes as S

pytestmark = pytest.mark.gpu


def render(product_lib, W, H, spp, frames=1, **settings):
    sc = S.cornell_box(unit_scale=True)
    ctx = R.RenderContext(product_lib)
    S.upload(ctx, sc, W, H)
    ctx.set_setting("spp", spp)
    for k, v in settings.items():
        ctx.set_setting(k, v)
    cam = sc.camera(W, H)
    for f in range(frames):
        ctx.render_frame(cam, R.RESET if f == 0 else R.CONVERGE)
    return ctx, sc, cam


def test_feature_planes_of_the_first_vertex(product_lib, oracle_lib):
    W, H, spp = 128, 96, 4
    ctx, sc, cam = render(product_lib, W, H, spp, frames=2, aov="on")
    img, alb, nor = ctx.read_image().copy(), ctx.read_aov(0), ctx.read_aov(1)
    plain, _, _ = render(product_lib, W, H, spp, frames=2)
    assert np.array_equal(plain.read_image(), img)  # the planes do not touch the frame
    # what the camera rays hit, from the oracle: sky pixels carry no normal and the (black) sky as albedo
    o = R.RenderContext(oracle_lib)
    S.upload(o, sc, W, H)
    hit = np.ones((H, W), bool)
    inst = np.zeros((H, W), np.int64)
    for s in range(2 * spp):
        origins, dirs = o.generate_primary(cam, s)
        h = o.trace_closest(origins, dirs)
        hit &= (h["prim_id"] >= 0).reshape(H, W)
        # a surface = one face of one instance (the quads and box faces of the scene are two consecutive triangles each)
        face = (h["inst_id"].astype(np.int64) * 64 + h["prim_id"].astype(np.int64) // 2).reshape(H, W)
        inst = np.where(s == 0, face, np.where(inst == face, inst, -2))
    nlen = np.linalg.norm(nor[..., :3], axis=-1)
    assert np.isfinite(alb).all() and np.isfinite(nor).all()
    assert (nlen <= 1.0 + 1e-5).all()
    same_surface = hit & (inst >= 0)  # every sample of the pixel hit the same (flat-shaded) quad
    assert same_surface.mean() > 0.05
    # the sum of equal unit normals / samples (a camera ray on a shared edge may pick the neighbouring face on the GPU)
    assert (np.abs(nlen[same_surface] - 1.0) < 1e-4).mean() > 0.999
    # normals face the viewer: the Cornell camera looks along +z
    assert (nor[..., 2][same_surface] <= 1e-6).mean() > 0.999
    # diffuse albedo = colour * |cos| <= colour <= 1; the light (colour > 1) is clamped to 1 (kernels.cu:209)
    assert (alb[..., :3] >= 0).all() and (alb[..., :3] <= 1.0 + 1e-6).all()
    assert alb[..., :3][same_surface].mean() > 0.05


@pytest.mark.parametrize("settings", [{"spp_batch": 1}, {"sort": "off"}, {"spp_batch": 3, "sort_cell_bits": 4}, {"sample_layout": "planes"}])
def test_feature_planes_do_not_depend_on_the_schedule(product_lib, settings):
    W, H, spp = 200, 100, 8
    a, _, _ = render(product_lib, W, H, spp, frames=2, aov="on")
    b, _, _ = render(product_lib, W, H, spp, frames=2, aov="on", **settings)
    for which in (0, 1):
        assert np.array_equal(a.read_aov(which), b.read_aov(which))


def test_feature_plane_transform_and_errors(product_lib):
    W, H = 96, 64
    ctx, sc, cam = render(product_lib, W, H, 2, aov="on")
    n0 = ctx.read_aov(1).copy()
    flip = np.diag([1.0, -1.0, 2.0])
    ctx.set_aov_transform(flip)
    ctx.render_frame(cam, R.RESET)
    n1 = ctx.read_aov(1)
    assert np.allclose(n1[..., :3], n0[..., :3] * np.array([1.0, -1.0, 2.0], np.float32), atol=1e-6)
    off, _, _ = render(product_lib, W, H, 2)
    with pytest.raises(R.Rfwb200Error, match="aov"):
        off.read_aov(0)
