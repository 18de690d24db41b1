"""GPU parity of the display pass (rfwb200_read_display / rfwb200_tone_map, the reference's assets/shaders/tone-map.frag as run
by system::render_frame(toneMap=true)) against the CPU restatement oracle/tonemap.py: the RGBA8 bytes must be IDENTICAL (the
kernel rounds once per operation in the shader's order, like the float32 restatement)."""
import sys
from pathlib import Path

import numpy as np
import pytest

import rfwb200 as R
import scenes as S

sys.path.insert(0, str(Path(R.REPO_DIR)))
from oracle.tonemap import tone_map, tone_map_spec  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("contrast,brightness", [(1.0, 0.05), (0.0, 0.0), (1.7, -0.2)])
def test_display_bytes_match_the_restatement(product_lib, contrast, brightness):
    W, H = 160, 96
    sc = S.cornell_box(unit_scale=True)
    g = R.RenderContext(product_lib)
    S.upload(g, sc, W, H)
    g.set_setting("spp", 2)
    g.render_frame(sc.camera(W, H), R.RESET)
    img = g.read_image()
    got = g.read_display(contrast, brightness).reshape(H, W, 4)
    want = tone_map(img, contrast, brightness)
    assert np.array_equal(got, want), int((got != want).sum())
    assert np.abs(got.astype(np.float64) - tone_map_spec(img, contrast, brightness)).max() <= 0.5 + 1e-3
    assert got[..., :3].std() > 10  # an image, not a constant
    # the framebuffer itself is untouched by the display pass
    assert np.array_equal(g.read_image(), img)


def test_display_of_a_sharded_frame_keeps_the_tile_major_order(product_lib):
    W, H = 128, 64
    sc = S.cornell_box(unit_scale=True)
    g = R.RenderContext(product_lib)
    g.set_shard(1, 2)
    S.upload(g, sc, W, H)
    g.render_frame(sc.camera(W, H), R.RESET)
    fb = g.read_framebuffer()
    got = g.read_display(1.0, 0.05)
    assert got.shape == (g.local_pixel_count(), 4)
    assert np.array_equal(got, tone_map(fb, 1.0, 0.05))
