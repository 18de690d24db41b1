#!/usr/bin/env python3
"""profiles/r02/sass_excerpt.txt: per-kernel SASS instruction counts and opcode histograms (cuobjdump -sass of the built library) plus
the ptxas -v resource lines of the same build — so a reader need not regenerate them to check what DESIGN.md says about the kernels
(FMNMX3 in the node test, MATCH.ANY in k_shade, UBLKCP only in the staged variants, LDG.E.128 everywhere, register / spill counts)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
SO = REPO / "rendering-fw_b200" / "librfwb200.so"
OUT = REPO / "profiles" / "r02" / "sass_excerpt.txt"
WANT = {
    "k_wavefront_traceILb1ELi1ELi2ELb1ELb0ELb0EE": "k_wavefront_trace<PRIMARY=true, LQ=1, LEAN=2, PACKED, !STAGE_P, !TL>  (camera rays, default)",
    "k_wavefront_traceILb0ELi1ELi2ELb1ELb0ELb0EE": "k_wavefront_trace<PRIMARY=false, LQ=1, LEAN=2, PACKED, !STAGE_P, !TL> (bounce rays, default; the dominant kernel)",
    "k_wavefront_traceILb0ELi1ELi2ELb1ELb0ELb1EE": "k_wavefront_trace<PRIMARY=false, ..., TL> (two-level scenes)",
    "k_wavefront_traceILb0ELi1ELi2ELb1ELb1ELb0EE": "k_wavefront_trace<PRIMARY=false, ..., STAGE_P> (TMA-staged packed prefix, variant 13)",
    "7k_shadeE": "k_shade (-use_fast_math)", "12k_shade_ieeeE": "k_shade_ieee", "11k_sort_moveE": "k_sort_move", "11k_sort_scanE": "k_sort_scan",
    "6k_foldE": "k_fold", "7k_emodeE": "k_emode",
}
KEYS = ["LDG.E.128", "STG.E.128", "FMNMX3", "FMNMX", "FFMA", "VIMNMX", "VIMNMX3", "MATCH.ANY", "VOTE", "SHFL", "ATOMG", "RED", "LDL", "STL", "UBLKCP",
        "SYNCS", "MUFU.RCP", "BRA", "BSSY", "BSYNC"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(SO)], capture_output=True, text=True, check=True).stdout
    out = ["SASS excerpt of rendering-fw_b200/librfwb200.so (cuobjdump -sass) — the default trace kernels, their two-level and TMA-staged\n"
           "variants, k_shade and the re-ordering kernels: instruction counts, histogram of the opcodes the design statements rest on, and the\n"
           "ptxas -v resource lines of the same build (rendering-fw_b200/csrc/*.ptxas.log).  Regenerate: python tools/sass_excerpt.py\n\n"]
    out.append("architectures in the fatbin: " + ", ".join(sorted(set(re.findall(r"arch = (sm_\d+a?)", sass)))) + "\n\n")
    for f in re.split(r"\n\s*Function : ", sass)[1:]:
        name = f.split("\n", 1)[0].strip()
        hit = [v for k, v in WANT.items() if k in name]
        if not hit:
            continue
        ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, flags=re.M)
        c = collections.Counter()
        for i in ins:
            for k in KEYS:
                if i == k or i.startswith(k + ".") or (k in ("LDG.E.128", "STG.E.128") and i.startswith(k[:3]) and ".128" in i):
                    c[k] += 1
        out.append(f"{hit[0]}\n  symbol {name}\n  {len(ins)} SASS instructions; " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]) + "\n")
    out.append("\nptxas -v (registers, spills, stack) of the same build:\n")
    for log in ("kernels_trace", "kernels_shade", "kernels_shade_ieee", "geometry"):
        txt = (REPO / "rendering-fw_b200" / "csrc" / f"{log}.ptxas.log").read_text()
        for name, spill, used in re.findall(r"Compiling entry function '([^']+)'.*?\n.*?\n\s+(\d+ bytes stack frame, \d+ bytes spill stores, \d+ bytes spill loads)\n"
                                            r"ptxas info\s+: (Used \d+ registers[^\n]*)", txt):
            short = [v for k, v in WANT.items() if k in name]
            if short or any(k in name for k in ("k_refit", "k_ploc", "k_lbvh_collapse", "k_flatten_shade", "k_pack_nodes", "k_skin")):
                out.append(f"  {short[0] if short else name[name.find('k_'):][:40]}: {used}; {spill}\n")
    OUT.write_text("".join(out))
    sys.stdout.write("".join(out))


if __name__ == "__main__":
    main()
