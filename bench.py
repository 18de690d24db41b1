#!/usr/bin/env python3
"""bench.py — Msamples/s of the wavefront path tracer on BASELINE.json's configs, through the C ABI of librfwb200.so.

  python bench.py [--config C] --gpus N --steps K --warmup W     our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference [--config C] --gpus N ...     the reference's CPU path on all host cores, bounded sample

Configs (BASELINE.json `configs`); the default is 2, the one the metric is quoted on:
  1  Cornell box 512x512, 1 spp, E-mode (the image model of the reference's Embree backend)
  2  Sponza 1920x1080, 8 spp, PT-mode, MAX_PATH_LENGTH 2
  3  10 M triangles (Sponza instanced 38x, SURVEY.md §8d stand-in for San Miguel), 1920x1080, 4 spp
  4  animated skinned mesh (CesiumMan, 19 joints), per-frame pose -> refit -> render, 1920x1080, 1 spp
  5  Sponza 3840x2160, 16 spp, area + point + directional lights, tile-sharded over the GPUs

A step = one frame: render_frame(RESET) with the config's spp (config 4: set_mesh_pose + update + render_frame).  For N > 1
the frame is tile-sharded over the ranks (strong scaling: the frame is fixed); every rank's last launch writes its tiles
straight into ONE image in rank 0's memory (CUDA IPC mapping, peer stores over NVLink — rfwb200_display_*), rank 0 waits for
the N arrivals.  No collective is on the data path; torch.distributed only carries the 64-byte IPC handle and the barriers
of the timing protocol.  Prints ONE JSON line (contract in the task statement; roofline / cpu_baseline / e2e included).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO))

MAX_PATH = 2
CONFIGS = {
    1: dict(width=512, height=512, spp=1, mode="embree", scene="cornell", cpu_sample=(512, 512), label="configs[0]"),
    2: dict(width=1920, height=1080, spp=8, mode="pt", scene="sponza", cpu_sample=(1920, 1080), label="configs[1]"),
    3: dict(width=1920, height=1080, spp=4, mode="pt", scene="sponza_x38", cpu_sample=(960, 540), label="configs[2]"),
    4: dict(width=1920, height=1080, spp=1, mode="pt", scene="animated", cpu_sample=(1920, 1080), label="configs[3]"),
    5: dict(width=3840, height=2160, spp=16, mode="pt", scene="sponza_lights", cpu_sample=(960, 540), label="configs[4]"),
}
REF_KERNELS_SAMPLE = (960, 540)  # the reference's own kernels on the host: a quarter-resolution frame of config 2's camera
MIN_LOAD_SECONDS = 2.5           # the clock sampler watches at least this much of the workload


def build_scene(cfg):
    import scenes as S

    kind = cfg["scene"]
    skins = []
    if kind == "cornell":
        sc = S.cornell_box()
    elif kind == "sponza":
        sc = S.sponza_or_standin()
    elif kind == "sponza_x38":
        sc = S.sponza_instanced(38)
    elif kind == "sponza_lights":
        sc = S.add_config5_lights(S.sponza_or_standin())
    elif kind == "animated":
        sc, skins = S.animated_config4(1)
    else:
        raise ValueError(kind)
    return sc, skins


def workload_name(sc, cfg):
    mode = "E-mode" if cfg["mode"] == "embree" else f"PT-mode max_path_length={MAX_PATH}"
    extra = ", per-frame pose + device refit" if cfg["scene"] == "animated" else ""
    lights = ", + point + directional light" if cfg["scene"] == "sponza_lights" else ""
    return f"{sc.name}{lights} {cfg['width']}x{cfg['height']} {cfg['spp']}spp {mode}{extra} (BASELINE.json {cfg['label']})"


def data_note(sc):
    if "atrium" in sc.name or sc.name == "config4":
        return "synthetic (procedural stand-in: the baked reference asset is absent)"
    if "cornell" in sc.name:
        return "synthetic (the procedural Cornell box of SURVEY.md §8d config 1) + deterministic seeds"
    return "reference assets baked as shipped (geometry, textures, sky; config 4: the CesiumMan mesh, skin and animation) + deterministic seeds"


def base_config(sc, cfg, n, gpus):
    """`config` of the JSON line: the same dictionary in both arms (the GPU arm's own details go under the top-level key `gpu`)"""
    return {"workload": workload_name(sc, cfg), "config": n, "width": cfg["width"], "height": cfg["height"], "spp": cfg["spp"],
            "max_path_length": MAX_PATH if cfg["mode"] == "pt" else 0, "mode": cfg["mode"],
            "parallelism": "single" if gpus <= 1 or cfg["mode"] != "pt" else f"tile{gpus}",
            "l2": "inputs larger than L2: the wavefront state of a step (~3 GB at config 2) streams through HBM; the BVH and triangles "
                  "(~50 MB for Sponza) are meant to stay in the 126 MB L2; no flush between steps"}


def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the workload runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, l in self.lines:
            if t < t_begin or t > t_end:
                continue
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])), mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "window_s": round(t_end - t_begin, 2)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arms: the oracle port (and, for config 2, the reference's own Kernels.cu compiled for the host)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_frame_runner(cfg):
    """-> (one() -> seconds of one frame, samples per frame, threads, description, context, scene): the oracle's frame of
    this config at the bounded sample size, scene already uploaded."""
    import rfwb200 as R
    import scenes as S
    from oracle.oracle_lib import load_oracle  # the checker, timed here as the CPU baseline (never on the product path)

    lib = load_oracle(fast=True)
    sc, skins = build_scene(cfg)
    w, h = cfg["cpu_sample"]
    ctx = R.RenderContext(lib)
    S.upload(ctx, sc, w, h)
    ctx.set_setting("mode", cfg["mode"])
    ctx.set_setting("spp", cfg["spp"])
    if cfg["mode"] == "pt":
        ctx.set_setting("max_path_length", MAX_PATH)
    cam = sc.camera(w, h)
    threads = int(lib.fn("num_threads", ctypes.c_int, [])())
    frame = [0]
    if skins:
        from oracle import skinning as K  # numpy restatement of the reference's CPU skinning (gltf/mesh.cpp:18-48,428-449)

        def one():
            t0 = time.perf_counter()
            for sk in skins:  # the reference's route: skin on the CPU, re-send the mesh, refit, render
                m = sc.meshes[sk.mesh_index]
                v, n = K.set_pose(sk.base_vertices, sk.base_normals, sk.joints, sk.weights, sk.joint_matrices(frame[0]))
                ctx.set_mesh(sk.mesh_index, v, K.update_triangles(m.triangles, v, n, m.indices), m.indices)
            ctx.update()
            ctx.render_frame(cam, R.RESET)
            frame[0] += 1
            return time.perf_counter() - t0
    else:
        def one():
            t0 = time.perf_counter()
            ctx.render_frame(cam, R.RESET)
            return time.perf_counter() - t0
    what = f"{w}x{h} x {cfg['spp']} spp"
    full = (w, h) == (cfg["width"], cfg["height"])
    what += " (the whole frame of the workload)" if full else f" (the workload's scene and camera at reduced resolution; the frame is {cfg['width']}x{cfg['height']})"
    return one, w * h * cfg["spp"], threads, what, ctx, sc


def embree_probe():
    """BASELINE.md §3 "real Embree (conditional)": is Intel Embree (and TBB, which the reference's EmbreeRT backend needs beside it)
    installed on this box?  Reported, never assumed; when absent the CPU rows are the restatement's."""
    found = {"libembree": [], "libtbb": []}
    try:
        out = subprocess.run(["ldconfig", "-p"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout
        for l in out.splitlines():
            name = l.strip().split(" ")[0]
            if name.startswith("libembree"):
                found["libembree"].append(name)
            if name.startswith("libtbb"):
                found["libtbb"].append(name)
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"}
    usable = bool(found["libembree"]) and bool(found["libtbb"])
    return {"libembree": sorted(set(found["libembree"])), "libtbb": sorted(set(found["libtbb"])), "usable": usable,
            "note": "the reference's EmbreeRT backend could be timed directly" if usable else
                    "Intel Embree is not installed on this box: the CPU rows are the restatement's (a scalar port; real Embree would be faster)"}


def time_cpu(one, budget_s, warmup, steps):
    """`warmup` untimed frames, then up to `steps` timed ones, stopping early once `budget_s` is spent (at least one)."""
    for _ in range(warmup):
        one()
    times = []
    t_start = time.perf_counter()
    for _ in range(max(1, steps)):
        times.append(one())
        if time.perf_counter() - t_start > budget_s:
            break
    return sum(times) / len(times), len(times)


def reference_kernels_on_host(octx, sc, spp):
    """The reference's OWN wavefront kernels (CUDART/src/Kernels.cu compiled for the host by oracle/ref_build into
    oracle/_ref/librfwref_kernels.so, CUDA threads run one after the other) on the workload's scene and camera: the samples of
    the frame are dealt to one forked process per sample.  Reported beside the oracle port; absent when oracle/_ref is."""
    sys.path.insert(0, str(REPO / "tests"))
    try:
        import numpy as np
        import ref_pin_common as P
        if not P.REF_KERNELS_LIB.exists():
            return None
        w, h = REF_KERNELS_SAMPLE
        rs, keep = P.reference_kernels_scene(octx, sc)
        v = sc.camera(w, h).get_view()
        v14 = np.array(list(v.pos) + list(v.p1) + list(v.p2) + list(v.p3) + [v.aperture, v.spread_angle], np.float32)
        procs = max(1, min(os.cpu_count() or 1, spp))
        secs, mean = P.reference_kernels_timed(rs, v14, w, h, spp, procs)
        return {"value": w * h * spp / secs / 1e6, "unit": "Msamples/s", "cores": procs, "kind": "reference",
                "sample": f"{w}x{h} x {spp} spp (quarter-resolution frame of the workload's camera), {secs:.1f} s",
                "mean_radiance": mean,
                "note": "RFW/backends/CUDART/src/Kernels.cu itself, compiled for the host (oracle/ref_build/ref_kernels_shim.cpp), one "
                        "process per sample"}
    except Exception as e:  # the checker must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores, on the same config / metric /
    unit as our arm.  Neither of the reference's renderers can be built in this image (Embree, TBB, glm, the
    Rust BVH crate are absent), so the arm's value is the oracle port — the CPU restatement pinned on the reference's own
    kernels (`cpu_baseline.kind` = "port"); for config 2 the reference's own Kernels.cu compiled for the host is timed
    beside it and reported in `reference_kernels_on_host`."""
    import __graft_entry__ as g

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.oracle_lib import ORACLE_FAST_LIB

    if not ORACLE_FAST_LIB.exists():
        g.build()
    cfg = CONFIGS[args.config]
    one, samples, threads, what, octx, sc = cpu_frame_runner(cfg)
    budget = 60.0  # the whole arm ends within a few minutes whatever --steps asks for
    warmup = min(args.warmup, 1)
    dt, done = time_cpu(one, budget, warmup, args.steps)
    val = samples / dt / 1e6
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": done,
        "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": data_note(sc), "config": base_config(sc, cfg, args.config, args.gpus),
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": f"{what}, {done} frame(s) of {dt:.2f} s"},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "protocol_note": f"a CPU frame takes seconds: at most {budget:.0f} s of timed frames after one warm-up frame are run "
                         "(`steps` / `warmup` are what was run); the value is the mean frame, so the ratio compares per-frame throughput",
    }
    line["embree_probe"] = embree_probe()
    if args.config == 2:
        ref_kernels = reference_kernels_on_host(octx, sc, cfg["spp"])
        if ref_kernels:
            line["reference_kernels_on_host"] = ref_kernels
    emit(json.dumps(line))
    return 0


class OneLineStdout:
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL's version banner, compiler chatter of a
    first-time build), so for the lifetime of the run file descriptor 1 points at stderr and the line goes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


OUT = None


def emit(text: str):
    OUT.emit(text) if OUT else print(text)


def main():
    global OUT
    with OneLineStdout() as OUT:
        return run()


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed frames (default: enough for ~3 s)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--set", nargs="*", default=[], help="extra key=value settings (tuning experiments)")
    ap.add_argument("--inflight", type=int, default=2, help="N > 1: frames a rank may have enqueued ahead of the device")
    ap.add_argument("--no-check-image", action="store_true", help="N > 1: skip the comparison of the assembled frame with a single-rank render")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps is None:
            args.steps = 3
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import rfwb200 as R
    import scenes as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    cfg_no, cfg = args.config, CONFIGS[args.config]
    W, H, SPP = cfg["width"], cfg["height"], cfg["spp"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")  # these levels print a version banner on stdout; keep stdout to the ONE JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    warmup = max(3, args.warmup)

    lib = R.load_product()
    sc, skins = build_scene(cfg)
    ctx = R.RenderContext(lib, local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        ctx.set_shard(rank, world, 32, 8)
    t_up = time.perf_counter()
    S.upload(ctx, sc, W, H)
    upload_s = time.perf_counter() - t_up

    def apply_settings(c):
        c.set_setting("mode", cfg["mode"])
        c.set_setting("spp", SPP)
        if cfg["mode"] == "pt":
            c.set_setting("max_path_length", MAX_PATH)
        for kv in args.set:
            k, v = kv.split("=")
            c.set_setting(k, v)

    apply_settings(ctx)
    ctx.update()  # a setting of --set may ask for another form of the scene (builder, levels, bvh); a no-op otherwise
    for sk in skins:
        ctx.set_mesh_skin(sk.mesh_index, sk.base_vertices, sk.base_normals, sk.joints, sk.weights)
    cam = sc.camera(W, H)
    view = cam.get_view()
    pt_mode = cfg["mode"] == "pt"

    # ---- the display image of the sharded frame: rank 0 owns it, the other ranks map it over CUDA IPC -----------------------
    sharded = world > 1 and pt_mode
    if sharded:
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ctx.display_create()
            handle.copy_(torch.frombuffer(bytearray(ctx.display_export()), dtype=torch.uint8))
        dist.broadcast(handle, 0)
        if rank != 0:
            ctx.display_import(handle.cpu().numpy().tobytes())
    frame_no = [0]
    # at most INFLIGHT frames are enqueued ahead of the device (what a presenting application does: double buffering).  Without
    # the bound the host of each rank runs hundreds of frames ahead and the ranks' launch queues drift apart; the sharded frame
    # (ranks meet twice per bounce and once per image) then measured 2.53 ms at 8 GPUs in the free-running loop against 2.05 ms
    # in the e2e loop, which is paced by its read-back (profiles/r02/r02_bench8_c2.json).
    INFLIGHT = max(1, args.inflight)
    pace = [torch.cuda.Event() for _ in range(INFLIGHT)]
    paced = [0]

    def step():
        if skins:  # config 4: this frame's pose (64 B per joint) -> GPU skinning -> device refit
            for sk in skins:
                ctx.set_mesh_pose(sk.mesh_index, sk.joint_matrices(frame_no[0]))
            ctx.update()
            frame_no[0] += 1
        ctx.render_frame(view, R.RESET)
        if sharded and rank == 0:
            ctx.display_wait()  # rank 0's stream continues when all ranks' tiles of this frame are in its image
        if not sharded:
            return
        k = paced[0] % INFLIGHT
        if paced[0] >= INFLIGHT:
            pace[k].synchronize()
        pace[k].record(stream)
        paced[0] += 1

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    t0 = time.perf_counter()
    for _ in range(warmup):
        step()
    fence()
    warm_ms = (time.perf_counter() - t0) * 1e3 / warmup
    if args.steps is not None:
        steps = max(1, args.steps)
    else:  # enough frames for ~3 s (all ranks must agree on the count)
        n = torch.tensor([max(20, int(3000.0 / max(warm_ms, 0.05)))], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.broadcast(n, 0)
        steps = int(n.item())

    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    fence()  # all ranks start the timed region together (rank 0 has just spent 150 ms starting the sampler)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    fence()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / steps
    value = W * H * SPP / (ms_per_step * 1e-3) / 1e6
    # Keep the same workload running until the clock sampler has seen MIN_LOAD_SECONDS of it.  `value` comes from exactly the K
    # timed steps above; a 20-step run lasts a quarter of a second, which nvidia-smi cannot observe.
    n_soak = torch.tensor([0], dtype=torch.int64, device="cuda")
    if rank == 0:
        n_soak[0] = max(0, int((MIN_LOAD_SECONDS - (time.perf_counter() - t_begin)) / max(ms_per_step * 1e-3, 1e-5)) + 1)
    if world > 1:
        dist.broadcast(n_soak, 0)
    soak = int(n_soak.item())
    for _ in range(soak):
        step()
    fence()
    clocks = sampler.stop(t_begin, time.perf_counter()) if rank == 0 else None
    if clocks is not None:
        clocks["soak_steps_after_the_timed_region"] = soak

    # ---- e2e: the call a user makes, host buffers on both sides ---------------------------------------------------------
    # per step: the camera (and, config 4, the joint matrices) go host->device inside the calls, the finished HDR frame
    # comes back into pinned host memory (rank 0 reads the assembled image when sharded).  The read-back of frame i runs
    # beside the kernels of frame i+1 (rfwb200_read_framebuffer_async): the step costs max(render, copy), not their sum.
    reads = world == 1 or rank == 0
    pinned = [torch.empty((W * H if reads else 1, 4), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    e2e_steps = max(3, min(steps, 50))
    checksum = [0.0]

    def e2e_loop(n):
        for i in range(n):
            step()
            if reads:
                if i > 0:
                    ctx.read_wait()  # frame i-1 has landed in pinned[(i-1) % 2]; frame i is already rendering
                    checksum[0] += float(pinned[(i - 1) % 2][W * (H // 2) + W // 2, 0])
                ctx.read_framebuffer_async(pinned[i % 2].data_ptr(), W * H)
        if reads:
            ctx.read_wait()
            checksum[0] += float(pinned[(n - 1) % 2][W * (H // 2) + W // 2, 0])

    e2e_loop(2)
    fence()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    torch.cuda.synchronize()
    e2e_wall = (time.perf_counter() - t0) * 1e3 / e2e_steps
    fence()
    te = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e_value = W * H * SPP / (e2e_ms * 1e-3) / 1e6
    # the unpipelined form (render, then a blocking read): what round 1 reported
    nblock = min(e2e_steps, 10)
    t0 = time.perf_counter()
    for _ in range(nblock):
        step()
        if reads:
            ctx.read_framebuffer(pinned[0].numpy())
    fence()
    e2e_blocking_ms = (time.perf_counter() - t0) * 1e3 / nblock
    h2d_bytes = 72 + 16 + sum(64 * sk.n_joints for sk in skins)  # FrameParams + probe reset (+ joint matrices); the scene is resident
    d2h_bytes = W * H * 16

    # ---- the sharded frame equals a single-rank render of it (outside the timed region) -----------------------------------
    image_check = None
    if sharded and not args.no_check_image and not skins:
        step()
        fence()
        if rank == 0:
            a = ctx.read_framebuffer().reshape(H, W, 4).copy()
        fence()
        if rank == 0:
            try:
                full = R.RenderContext(lib, local_rank)
                S.upload(full, sc, W, H)
                apply_settings(full)
                full.render_frame(view, R.RESET)
                b = full.read_image()
                x = lambda im: int(np.bitwise_xor.reduce(im.view(np.uint32).reshape(-1)))
                image_check = {"identical_to_single_rank_frame": bool(np.array_equal(a, b)), "pixels_differing": int((np.abs(a - b).max(axis=-1) > 0).sum()),
                               "mean": float(a[..., :3].mean()), "xor_of_all_words": [x(a), x(b)]}
                full.close()
            except Exception as e:
                image_check = {"unavailable": f"{type(e).__name__}: {e}"}
        fence()

    # ---- per-stage device times (separate pass so the events do not perturb `value`) -------------------------------
    ctx.set_setting("timing", "on")
    stage = {"primary": 0.0, "trace": 0.0, "shade": 0.0, "sort": 0.0, "finalize": 0.0}
    reps = 3
    for _ in range(reps):
        step()
        st = ctx.get_stats()
        stage["primary"] += st.primary_time / reps
        stage["trace"] += (st.secondary_time + st.deep_time) / reps
        stage["shade"] += st.shade_time / reps
        stage["sort"] += st.animation_time / reps  # the re-ordering pass (no field of the reference's RenderStats fits it)
        stage["finalize"] += st.finalize_time / reps
    ctx.set_setting("timing", "off")
    fence()
    geo = ctx.get_geometry_stats() if skins else None
    fc = ctx.get_frame_counters()
    counters = fc.as_dict()
    prim = fc.pixels * fc.samples
    local_px = max(32, -(-W * H // world))
    nb = -(-SPP // max(1, min(SPP, 64, (1 << 24) // local_px))) if pt_mode else 1  # wavefronts per frame (rfwb200: batch_spp_for)
    if pt_mode:
        alg = {
            "primary": prim * 48,                                   # N_gen*32 (write O,D) + N_ext(primary)*16 (write hit)
            "trace": (fc.n_ext - prim) * 48 + fc.n_nee * 48,        # read O,D + write hit ; read connect entry
            "shade": fc.n_shade * 224 + fc.n_ext_out * 48 + fc.n_nee * 48,  # read state + triangle ; write ext ; write connect
            "finalize": fc.pixels * 32,
        }
        total_alg = fc.algorithmic_bytes()
        alg["acc_rmw"] = total_alg - sum(alg.values())
        launches_per_frame = {"primary": nb, "trace": nb * MAX_PATH, "shade": nb * (MAX_PATH + 1), "finalize": nb}
        kernel_names = {"primary": "k_wavefront_trace<PRIMARY=true>", "trace": "k_wavefront_trace<PRIMARY=false>", "shade": "k_shade"}
        dominant = max(("primary", "trace", "shade"), key=lambda k: stage[k])
    else:  # E-mode: one fused kernel per frame; compulsory traffic = the frame it writes
        alg = {"primary": fc.pixels * 16}
        total_alg = fc.pixels * 16
        launches_per_frame = {"primary": 1}
        kernel_names = {"primary": "k_emode"}
        dominant = "primary"
    peak, peak_src = load_peaks()
    dom_ms_per_launch = stage[dominant] / launches_per_frame[dominant]
    dom_bytes_per_launch = alg[dominant] / launches_per_frame[dominant]
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9 if dom_ms_per_launch > 0 else 0.0
    traffic, traffic_src = None, None
    prof = REPO / "profiles" / "ncu_summary.json"
    if prof.exists() and world == 1:  # a 1-GPU capture says nothing about a rank of a sharded frame
        try:
            entry = json.loads(prof.read_text()).get("r02", {}).get(f"config{cfg_no}", {}).get(dominant)
            if entry:
                traffic, traffic_src = entry.get("dram_bytes_per_launch"), entry.get("source")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": kernel_names[dominant],
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
        "launches_per_frame": launches_per_frame, "stage_ms_per_frame": stage, "stage_algorithmic_bytes_per_frame": alg,
        "scope": "this rank's share of the frame" if world > 1 else "the whole frame",
        "note": "traversal and shading are issue-bound, not HBM-bound (ncu, profiles/r02: trace issue-active 73-78 %, 15.5 of 32 threads "
                "per instruction, DRAM a few per cent of peak): BVH nodes, triangles, materials and textures are cache-resident and "
                "excluded from the algorithmic bytes by definition (SURVEY.md §8d); `traffic` is the ncu DRAM bytes per launch of the "
                "same kernel on this config (null when no capture of it is committed, and at N > 1)",
    }
    if world == 1:  # at N > 1 the counters are one rank's while the step is the whole job's
        frame_gbs = total_alg / (ms_per_step * 1e-3) / 1e9
        roofline["whole_frame"] = {"algorithmic_bytes": total_alg, "achieved_gbs": frame_gbs, "frac_of_measured": frame_gbs / peak,
                                   "frac_of_8TBs": frame_gbs / 8000.0, "bytes_per_sample": total_alg / max(prim, 1), "over": "ms_per_step"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        one, samples, threads, what, octx, osc = cpu_frame_runner(cfg)
        dt, done = time_cpu(one, 20.0, 1, 2)
        cpu = {"value": samples / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
               "sample": f"{what}: {done} frame(s) of {dt:.2f} s after one warm-up frame",
               "note": "CPU restatement of the reference's estimator (oracle/, pinned on the reference's own kernels); the reference's "
                       "renderers cannot be built here (DESIGN.md §4) and real Embree would be faster than a scalar restatement"}
        octx.close()

    # ---- E-mode of config 2 (the "Embree image" of the headline scene), outside the timed region ---------------
    emode = None
    if cfg_no == 2 and world == 1 and rank == 0 and not args.no_cpu_baseline:
        ctx.set_setting("mode", "embree")
        for _ in range(3):
            ctx.render_frame(view, R.RESET)
        torch.cuda.synchronize()
        n_e = 50
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(n_e):
            ctx.render_frame(view, R.RESET)
        a1.record(stream)
        torch.cuda.synchronize()
        e_ms = a0.elapsed_time(a1) / n_e
        ctx.set_setting("mode", "pt")
        e_cfg = dict(cfg, mode="embree", spp=1)
        one_e, samples_e, threads_e, what_e, octx_e, _ = cpu_frame_runner(e_cfg)
        dt_e, done_e = time_cpu(one_e, 10.0, 1, 3)
        octx_e.close()
        emode = {"what": "E-mode (primary visibility + direct light of the reference's EmbreeRT backend, one sample per frame) of the same scene and camera",
                 "ms_per_frame": e_ms, "mpixels_per_s": W * H / (e_ms * 1e-3) / 1e6,
                 "cpu_port": {"mpixels_per_s": samples_e / dt_e / 1e6, "cores": threads_e, "sample": f"{what_e}, {done_e} frame(s) of {dt_e:.2f} s"}}

    if rank == 0:
        bvh = ctx.get_bvh_info()
        conf = base_config(sc, cfg, cfg_no, world)
        gpu_conf =    {"tile": [32, 8], "wavefronts_per_frame": nb,
                       "assembly": "peer stores into rank 0's image over CUDA IPC (no collective)" if sharded else "none",
                       "triangles": bvh["triangles"], "bvh_nodes": bvh["nodes"], "bvh_build_ms": bvh["build_ms"], "scene_upload_s": upload_s,
                       "settings": args.set,
                       "frames_in_flight": INFLIGHT if sharded else "not bounded (one device, nothing to drift apart)"}
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": data_note(sc), "config": conf, "gpu": gpu_conf, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms, "steps": e2e_steps, "checksum": checksum[0],
                    "what": "render_frame + read_framebuffer_async into pinned host memory through the C ABI, every frame read back; the "
                            "copy of frame i overlaps the kernels of frame i+1 (wall clock over the loop including the last read)",
                    "blocking_ms_per_step": e2e_blocking_ms, "blocking_value": W * H * SPP / (e2e_blocking_ms * 1e-3) / 1e6},
            "roofline": roofline, "counters": counters,
        }
        if skins:
            line["config4"] = {"what": "`value` includes the per-frame pose upload, GPU skinning and the device refit (rfwb200_set_mesh_pose + rfwb200_update)",
                               "geometry_device_ms": geo.device_ms if geo else None, "refits": int(geo.refits) if geo else None,
                               "render_only_ms": sum(stage.values())}
        if image_check is not None:
            line["image_check"] = image_check
        if cpu:
            line["cpu_baseline"] = cpu
            line["embree_probe"] = embree_probe()
        if emode:
            line["emode_of_config2"] = emode
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        if rank != 0:
            ctx.close()  # importers unmap the display image before its owner frees it
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
