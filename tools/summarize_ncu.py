#!/usr/bin/env python3
"""Turn the raw ncu outputs a gpurun call brought back into the small, tracked summaries under profiles/:

  summarize_ncu.py launches <launches.csv> <out.json>      per-kernel time shares of one `--metrics gpu__time_duration.sum` pass
  summarize_ncu.py full <report.ncu-rep> <out.json>        per-launch figures of a `--set full` capture (needs ncu here)
"""
import collections
import csv
import json
import subprocess
import sys


def kernel_label(name: str) -> str:
    base = name.split("(")[0].split("::")[-1].replace("void ", "").strip()
    return base


def launches(path, out):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hi], rows[hi + 1:]
    ik, iv, im, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in data:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)  # -> us
        k = kernel_label(r[ik])
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
    s = sum(tot.values())
    res = {"source": path, "launches": int(sum(cnt.values())), "total_us": s,
           "kernels": {k: {"launches": cnt[k], "total_us": round(v, 1), "avg_us": round(v / cnt[k], 2), "share": round(v / s, 4)} for k, v in tot.items()}}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res["kernels"], indent=1))


WANT = {
    "gpu__time_duration.sum": "gpu_time_us", "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst", "smsp__inst_executed.sum": "warp_instructions",
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_sm", "launch__registers_per_thread": "registers", "launch__grid_size": "grid",
    "smsp__sass_inst_executed_op_global_ld.sum": "global_load_instructions", "l1tex__data_pipe_lsu_wavefronts.sum": "l1_wavefronts",
}


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = {"kernel": kernel_label(r[h.index("Kernel Name")])}
        for m, label in WANT.items():
            if m in h:
                try:
                    d[label] = float(r[h.index(m)].replace(",", ""))
                except ValueError:
                    pass
        for m, label in (("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB")):
            if label in d:  # ncu picks a unit per column: normalise to MB
                unit = u[h.index(m)].lower()
                d[label] *= {"g": 1e3, "m": 1.0, "k": 1e-3, "b": 1e-6}.get(unit[:1], 1.0)
        if "gpu_time_us" in d:
            unit = u[h.index("gpu__time_duration.sum")].lower()
            d["gpu_time_us"] *= {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(unit.replace("econd", ""), 1.0)
        res.append(d)
    json.dump({"source": rep, "launches": res}, open(out, "w"), indent=1)
    for d in res:
        print(d)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
