"""The multi-GPU product path: tile shards folded straight into ONE display image by peer stores, no
collective.  All of it runs on a box with a single GPU — the ranks of a group may share a device, and CUDA IPC works
between two processes on one device — so the driver's 1-GPU test tier covers the same code the 8-GPU bench runs."""
import multiprocessing as mp
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import rfwb200 as R
import scenes as S

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent
BUILD = R.PKG_DIR / "host" / "_build"


def unit_cornell():
    return S.cornell_box(unit_scale=True)


def dying_wavefront():
    """the adapter driver's scene: no path continues after the first bounce, so the reference's loop leaves every sample
    before its depth-1 shadow rays are traced (CUDART/src/Context.cpp:109-120) — the ranks of a sharded frame have to agree
    on that (k_shard_sync), although each of them only sees a few tiles"""
    from test_adapter import _driver_scene

    return _driver_scene()


def _single(product_lib, scene_fn, W, H, spp, frames=1, **settings):
    sc = scene_fn()
    ctx = R.RenderContext(product_lib)
    S.upload(ctx, sc, W, H)
    ctx.set_setting("spp", spp)
    for k, v in settings.items():
        ctx.set_setting(k, v)
    cam = sc.camera(W, H)
    for f in range(frames):
        ctx.render_frame(cam, R.RESET if f == 0 else R.CONVERGE)
    img = ctx.read_image().copy()
    counters = ctx.get_frame_counters().as_dict()
    ctx.close()
    return img, counters


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("scene", [unit_cornell, S.feature_soup, dying_wavefront])
def test_device_group_frame_is_bit_identical_to_one_device(product_lib, scene, world):
    """rfwb200_create_group: one context, `world` ranks (here all on device 0), one render_frame call; every rank's fold
    kernel writes its tiles into rank 0's display image.  Frame, counters and probe equal the single-device context's."""
    W, H = 200, 100  # not a multiple of the 32x8 tile: padded edge tiles, ranks with unequal tile counts
    ref, ref_counters = _single(product_lib, scene, W, H, 4, frames=2)
    sc = scene()
    grp = R.RenderContext(product_lib, devices=[0] * world)
    S.upload(grp, sc, W, H)
    grp.set_setting("spp", 4)
    cam = sc.camera(W, H)
    grp.render_frame(cam, R.RESET)
    first = grp.read_image().copy()
    grp.render_frame(cam, R.CONVERGE)  # second frame: flow control (release / gate) of the shared image
    img = grp.read_image().copy()
    assert np.array_equal(img, ref)
    assert not np.array_equal(first, img)
    assert grp.get_frame_counters().as_dict() == ref_counters
    assert grp.local_pixel_count() == W * H
    grp.synchronize()
    grp.close()


def test_device_group_probe_stats_and_refit(product_lib):
    W, H = 128, 96
    sc = unit_cornell()
    one = R.RenderContext(product_lib)
    grp = R.RenderContext(product_lib, devices=[0, 0, 0, 0])
    for ctx in (one, grp):
        S.upload(ctx, sc, W, H)
        ctx.set_setting("spp", 2)
        ctx.set_probe_index(W // 2, H - 5)
    cam = sc.camera(W, H)
    one.render_frame(cam, R.RESET), grp.render_frame(cam, R.RESET)
    assert one.get_probe_results() == grp.get_probe_results()
    so, sg = one.get_stats(), grp.get_stats()
    assert (so.primary_count, so.secondary_count, so.deep_count, so.shadow_count) == (sg.primary_count, sg.secondary_count, sg.deep_count, sg.shadow_count)
    # move an instance: every rank refits its own copy on its device; frames stay identical
    T = np.eye(4, dtype=np.float32)
    T[:3, 3] = (0.2, 0.0, 0.1)
    for ctx in (one, grp):
        ctx.set_instance(len(sc.instances) - 1, sc.instances[-1][0], T @ np.asarray(sc.instances[-1][1], np.float32))
        ctx.update()
        ctx.render_frame(cam, R.RESET)
    assert np.array_equal(one.read_image(), grp.read_image())
    one.close(), grp.close()


def _ipc_rank(rank, world, W, H, spp, frames, q_handle, q_done, q_result):
    """one rank = one process, as under torchrun"""
    sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
    import rfwb200 as R2
    import scenes as S2

    lib = R2.load_product()
    sc = S2.cornell_box(unit_scale=True)
    ctx = R2.RenderContext(lib)
    ctx.set_shard(rank, world, 32, 8)
    S2.upload(ctx, sc, W, H)
    ctx.set_setting("spp", spp)
    if rank == 0:
        ctx.display_create()
        h = ctx.display_export()
        for _ in range(world - 1):
            q_handle.put(h)
    else:
        ctx.display_import(q_handle.get(timeout=120))
    cam = sc.camera(W, H)
    imgs = []
    for f in range(frames):
        ctx.render_frame(cam, R2.RESET if f == 0 else R2.CONVERGE)
        if rank == 0:
            ctx.display_wait()
            imgs.append(ctx.read_device(ctx.display_image(), W * H).reshape(H, W, 4).copy())
    ctx.synchronize()
    if rank == 0:
        q_result.put(imgs)
        for _ in range(world - 1):
            q_done.get(timeout=120)  # keep the allocation alive until every importer has closed it
    else:
        ctx.close()
        q_done.put(rank)


def test_display_image_across_processes_over_cuda_ipc(product_lib):
    """The torchrun layout: one process per rank; rank 0 exports the CUDA IPC handle of its display image, the other ranks
    map it and fold their tiles into it with peer stores; rank 0 waits for `world` arrivals per frame."""
    W, H, spp, world, frames = 200, 100, 2, 3, 3
    ref = []
    sc = unit_cornell()
    one = R.RenderContext(product_lib)
    S.upload(one, sc, W, H)
    one.set_setting("spp", spp)
    cam = sc.camera(W, H)
    for f in range(frames):
        one.render_frame(cam, R.RESET if f == 0 else R.CONVERGE)
        ref.append(one.read_image().copy())
    one.close()
    mpc = mp.get_context("spawn")
    qh, qd, qr = mpc.Queue(), mpc.Queue(), mpc.Queue()
    procs = [mpc.Process(target=_ipc_rank, args=(r, world, W, H, spp, frames, qh, qd, qr)) for r in range(world)]
    for p in procs:
        p.start()
    imgs = qr.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for f in range(frames):
        assert np.array_equal(imgs[f], ref[f]), f


needs_build = pytest.mark.skipif(not (BUILD / "adapter_driver").exists(), reason="adapter check build needs /root/reference (build container)")


@needs_build
def test_plugin_drives_a_device_group_and_fills_the_gl_texture(product_lib):
    """The reference-side plugin (B200RT.so, rfw::RenderContext vtable): RFWB200_DEVICES makes it own a device group, and
    with a texture target every render_frame ends with the frame uploaded into that texture (glTexSubImage2D, resolved from
    the host process — the driver exports stand-ins that record the upload), as EmbreeRT/src/Context.cpp:289-297 does."""
    import os

    exe, lib = str(BUILD / "adapter_driver"), str(BUILD / "B200RT.so")
    one = subprocess.run([exe, lib], capture_output=True, text=True)
    assert one.returncode == 0, one.stdout + one.stderr
    env = dict(os.environ, RFWB200_DEVICES="0,0,0")
    grp = subprocess.run([exe, lib, "--gl"], capture_output=True, text=True, env=env)
    assert grp.returncode == 0 and "adapter ok" in grp.stdout, grp.stdout + grp.stderr
    mean = lambda out: [l for l in out.splitlines() if l.startswith("mean")][0]
    assert mean(one.stdout) == mean(grp.stdout)  # mean radiance, probe and counts of the group equal one device's
    assert "gl uploads 2 tex 7 size 128x96 identical_to_read_pixels 1" in grp.stdout
    assert "targets 2" in grp.stdout  # BUFFER + OPENGL_TEXTURE: the GL entry points resolve in this host process
