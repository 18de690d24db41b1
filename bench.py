#!/usr/bin/env python3
"""bench.py — Msamples/s of the wavefront path tracer on BASELINE.json's config (Sponza 1920x1080, 8 spp,
MAX_PATH_LENGTH 2), through the C ABI of librfwb200.so.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path on all host cores, bounded sample: the
                                                           oracle port of its wavefront estimator and, where oracle/_ref
                                                           holds it, the reference's own Kernels.cu compiled for the host;
                                                           the faster of the two is the value

A step = one render_frame(RESET) with spp = 8: every stage of the hot path for W*H*spp samples + finalize;
for N > 1 the frame is tile-sharded over the ranks (strong scaling: the frame is fixed) and assembled by one
NCCL all-gather + a de-tiling kernel inside the timed region.  Prints ONE JSON line (contract in the task
statement; roofline / cpu_baseline / e2e objects included).
"""
from __future__ import annotations

import argparse
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

WIDTH, HEIGHT, SPP, MAX_PATH = 1920, 1080, 8, 2
CPU_SAMPLE = (1920, 1080)  # cpu baseline: one full frame of the workload (same camera, same spp): ~8-10 s on 16 cores
REF_KERNELS_SAMPLE = (960, 540)  # the reference's own kernels on the host: a quarter-resolution frame of the same camera, same spp


def load_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.lines, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
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
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def reference_kernels_on_host(octx, sc):
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
        procs = max(1, min(os.cpu_count() or 1, SPP))
        secs, mean = P.reference_kernels_timed(rs, v14, w, h, SPP, procs)
        return {"value": w * h * SPP / secs / 1e6, "unit": "Msamples/s", "cores": procs, "kind": "reference",
                "sample": f"{w}x{h} x {SPP} spp (quarter-resolution frame of the workload's camera), {secs:.1f} s",
                "mean_radiance": mean,
                "note": "RFW/backends/CUDART/src/Kernels.cu itself, compiled for the host (oracle/ref_build/ref_kernels_shim.cpp), one "
                        "process per sample; the baseline value is the faster of this and the oracle port"}
    except Exception as e:  # the checker must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's wavefront estimator (oracle port; the reference's
    own backends cannot be built in this image — DESIGN.md) on all host cores, bounded sample of the workload."""
    import __graft_entry__ as g
    import rfwb200 as R
    import scenes as S

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.oracle_lib import ORACLE_FAST_LIB, load_oracle  # the reference arm is the one place bench.py times oracle/

    if not ORACLE_FAST_LIB.exists():
        g.build()
    lib = load_oracle(fast=True)
    sc = S.sponza_or_standin()
    w, h = CPU_SAMPLE
    ctx = R.RenderContext(lib)
    S.upload(ctx, sc, w, h)
    ctx.set_setting("spp", SPP)
    ctx.set_setting("max_path_length", MAX_PATH)
    cam = sc.camera(w, h)
    threads = int(lib.fn("num_threads", __import__("ctypes").c_int, [])())
    for _ in range(max(0, min(args.warmup, 1))):
        ctx.render_frame(cam, R.RESET)
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        ctx.render_frame(cam, R.RESET)
    dt = (time.perf_counter() - t0) / steps
    val = w * h * SPP / dt / 1e6
    sample = f"{w}x{h} x {SPP} spp (the whole frame of the workload), {steps} frame(s)"
    ref_kernels = reference_kernels_on_host(ctx, sc)
    line = {
        "impl": "reference", "metric": "Msamples/s", "value": val, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic" if "atrium" in sc.name else "reference asset (baked)",
        "config": {"workload": workload_name(sc), "width": WIDTH, "height": HEIGHT, "spp": SPP, "max_path_length": MAX_PATH},
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if ref_kernels:
        line["reference_kernels_on_host"] = ref_kernels
        line["oracle_port"] = dict(line["cpu_baseline"])
        if ref_kernels.get("value", 0.0) > val:  # the stronger CPU baseline is the arm's value
            rv = ref_kernels["value"]
            line["value"], line["ms_per_step"] = rv, WIDTH * HEIGHT * SPP / rv / 1e3
            line["cpu_baseline"] = {k: ref_kernels[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["e2e"]["value"] = rv
    emit(json.dumps(line))
    return 0


def workload_name(sc):
    return f"{sc.name} {WIDTH}x{HEIGHT} {SPP}spp PT-mode max_path_length={MAX_PATH} (BASELINE.json configs[1])"


class DevPtr:
    """zero-copy view of a device allocation of the library as a torch tensor (__cuda_array_interface__)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 2}


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
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--smem-nodes", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import rfwb200 as R
    import scenes as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")  # these levels print a version banner on stdout; keep stdout to the ONE JSON line (INFO etc. are respected)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)

    lib = R.load_product()
    sc = S.sponza_or_standin()
    ctx = R.RenderContext(lib, local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        ctx.set_shard(rank, world, 32, 8)
    t_up = time.perf_counter()
    S.upload(ctx, sc, WIDTH, HEIGHT)
    upload_s = time.perf_counter() - t_up
    ctx.set_setting("spp", SPP)
    ctx.set_setting("max_path_length", MAX_PATH)
    if args.smem_nodes is not None:
        ctx.set_setting("smem_nodes", args.smem_nodes)
    cam = sc.camera(WIDTH, HEIGHT)
    view = cam.get_view()

    n_local = ctx.local_pixel_count()
    stride = ctx.shard_stride()
    if world > 1:
        gathered = torch.empty(world * stride * 4, dtype=torch.float32, device="cuda")
        image = torch.zeros(WIDTH * HEIGHT * 4, dtype=torch.float32, device="cuda")
        local = torch.as_tensor(DevPtr(ctx.device_framebuffer(), stride * 4), device="cuda")

    def step():
        ctx.render_frame(view, R.RESET)
        if world > 1:
            dist.all_gather_into_tensor(gathered, local)  # the one collective of the path
            ctx.assemble_shards(gathered.data_ptr(), image.data_ptr())

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    fence()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    fence()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / steps
    value = WIDTH * HEIGHT * SPP / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: the call a user makes, host buffers on both sides -------------------------------------------------
    # per step: camera/frame parameters go host->device inside render_frame, the finished HDR framebuffer comes
    # back into pinned host memory (rank 0 reads the assembled image when sharded)
    pinned = torch.empty((WIDTH * HEIGHT if (world == 1 or rank == 0) else 1, 4), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(3, min(steps, 10))

    def e2e_step():
        step()
        if world == 1:
            ctx.read_framebuffer(pinned.numpy())
        elif rank == 0:
            pinned.copy_(image.view(-1, 4), non_blocking=False)

    e2e_step()
    fence()
    t0 = time.perf_counter()
    e2a, e2b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2a.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e2b.record(stream)
    fence()
    e2e_ms = max(e2a.elapsed_time(e2b), (time.perf_counter() - t0) * 1e3) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te.item())
    e2e_value = WIDTH * HEIGHT * SPP / (e2e_ms * 1e-3) / 1e6
    # the same through the display pass (tone-map to RGBA8 on the device, a quarter of the read-back): informational
    e2e_display = None
    if world == 1:
        pinned8 = torch.empty((WIDTH * HEIGHT, 4), dtype=torch.uint8, pin_memory=True)

        def display_step():
            step()
            ctx.read_display(1.0, 0.05, out=pinned8.numpy())  # rfw::Camera's default contrast / brightness

        display_step()
        fence()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            display_step()
        fence()
        d_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_display = {"value": WIDTH * HEIGHT * SPP / (d_ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": d_ms,
                       "d2h_bytes_per_step": WIDTH * HEIGHT * 4,
                       "what": "render_frame + read_display (ACES tone-map to RGBA8 on the device) into pinned host memory"}
    h2d_bytes = 72 + 16  # FrameParams + probe reset (the scene is resident: upload is excluded by the metric)
    d2h_bytes = WIDTH * HEIGHT * 16

    # ---- per-stage device times (separate pass so the events do not perturb `value`) -------------------------------
    ctx.set_setting("timing", "on")
    stage = {"primary": 0.0, "trace": 0.0, "shade": 0.0, "finalize": 0.0}
    reps = 3
    for _ in range(reps):
        ctx.render_frame(view, R.RESET)
        st = ctx.get_stats()
        stage["primary"] += st.primary_time / reps
        stage["trace"] += (st.secondary_time + st.deep_time) / reps
        stage["shade"] += st.shade_time / reps
        stage["finalize"] += st.finalize_time / reps
    ctx.set_setting("timing", "off")
    fc = ctx.get_frame_counters()
    counters = fc.as_dict()
    prim = fc.pixels * fc.samples
    alg = {
        "primary": prim * 48,                                   # N_gen*32 (write O,D) + N_ext(primary)*16 (write hit)
        "trace": (fc.n_ext - prim) * 48 + fc.n_nee * 48,        # read O,D + write hit ; read connect entry
        "shade": fc.n_shade * 224 + fc.n_ext_out * 48 + fc.n_nee * 48,  # read state + triangle ; write ext ; write connect
        "finalize": fc.pixels * 32,
    }
    # accumulator read-modify-writes are split between shade (terminations) and trace (unoccluded connects);
    # attribute them to the stage total so the sum equals the SURVEY formula
    total_alg = fc.algorithmic_bytes()
    alg_sum = sum(alg.values())
    alg["acc_rmw"] = total_alg - alg_sum
    nb = max(1, int(st.wavefronts)) if hasattr(st, "wavefronts") else 1  # wavefronts per frame (all samples travel in one when they fit)
    launches_per_frame = {"primary": nb, "trace": nb * MAX_PATH, "shade": nb * (MAX_PATH + 1), "finalize": nb}
    dominant = max(("primary", "trace", "shade"), key=lambda k: stage[k])
    peak, peak_src = load_peaks()
    dom_ms_per_launch = stage[dominant] / launches_per_frame[dominant]
    dom_bytes_per_launch = alg[dominant] / launches_per_frame[dominant]
    achieved = dom_bytes_per_launch / (dom_ms_per_launch * 1e-3) / 1e9 if dom_ms_per_launch > 0 else 0.0
    traffic = None
    prof = REPO / "profiles" / "ncu_summary.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get(dominant, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    frame_gbs = total_alg / (sum(stage.values()) * 1e-3) / 1e9 if sum(stage.values()) > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": {"primary": "k_wavefront_trace<PRIMARY=true>", "trace": "k_wavefront_trace<PRIMARY=false>", "shade": "k_shade"}[dominant],
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes_per_launch, "ms_per_launch": dom_ms_per_launch,
        "stage_ms_per_frame": stage, "stage_algorithmic_bytes_per_frame": alg,
        "whole_frame": {"algorithmic_bytes": total_alg, "achieved_gbs": frame_gbs, "frac_of_measured": frame_gbs / peak,
                        "frac_of_8TBs": frame_gbs / 8000.0, "bytes_per_sample": total_alg / max(prim, 1)},
        "note": "traversal is issue/latency-bound, not HBM-bound (ncu: issue active 76 %, L1 data pipe 49 %, DRAM 2 %; "
                "profiles/ncu_summary.json): BVH nodes, triangles, materials and textures are cache-resident and excluded from the "
                "algorithmic bytes by definition (SURVEY.md §8d); `traffic` is the ncu DRAM bytes per launch of the same kernel",
    }

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.oracle_lib import load_oracle  # cpu_baseline leg only, after the timed region

        olib = load_oracle(fast=True)
        octx = R.RenderContext(olib)
        w, h = CPU_SAMPLE
        osc = S.sponza_or_standin()
        S.upload(octx, osc, w, h)
        octx.set_setting("spp", SPP)
        octx.set_setting("max_path_length", MAX_PATH)
        ocam = sc.camera(w, h)
        t0 = time.perf_counter()
        octx.render_frame(ocam, R.RESET)
        dt = time.perf_counter() - t0
        import ctypes

        cpu = {"value": w * h * SPP / dt / 1e6, "unit": "Msamples/s", "cores": int(olib.fn("num_threads", ctypes.c_int, [])()),
               "kind": "port", "sample": f"{w}x{h} x {SPP} spp (one whole frame of the workload), {dt:.1f} s",
               "note": "CPU restatement of the reference's wavefront estimator (oracle/); the reference's Embree backend is not a "
                       "path tracer and cannot be built here (DESIGN.md); `--impl reference` also times the reference's own "
                       "Kernels.cu compiled for the host (no fork from this CUDA process)"}

    if rank == 0:
        bvh = ctx.get_bvh_info()
        line = {
            "metric": "Msamples/s", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic" if "atrium" in sc.name else "reference asset (baked sponza.obj + textures + sky_15.hdr)",
            "config": {"workload": workload_name(sc), "width": WIDTH, "height": HEIGHT, "spp": SPP, "max_path_length": MAX_PATH,
                       "parallelism": f"tile{world}" if world > 1 else "single", "tile": [32, 8],
                       "l2": "wavefront state per step (~400 MB) exceeds the 126 MB L2; the BVH/triangles (~50 MB) are meant to stay L2-resident",
                       "triangles": bvh["triangles"], "bvh_nodes": bvh["nodes"], "bvh_build_ms": bvh["build_ms"], "scene_upload_s": upload_s},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "ms_per_step": e2e_ms, "what": "render_frame + read_framebuffer into pinned host memory through the C ABI"},
            "roofline": roofline, "counters": counters,
        }
        if e2e_display:
            line["e2e_display"] = e2e_display
        if cpu:
            line["cpu_baseline"] = cpu
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
