#!/usr/bin/env python3
"""Golden frames of the reference's own E-mode frame loop body (oracle/_ref/librfwref_eframe.so = EmbreeRT/src/Context.cpp:179-282
+ :417-476 compiled from /root/reference, Embree's two calls answered by the oracle's traversal): run here, where the reference
tree exists, and commit the frames as tests/golden/ref_eframe_vectors.npz.  tests/test_ref_pin.py compares the oracle's E-mode
frame with them everywhere and re-runs the live library where it is present."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(HERE.parent.parent))
import rfwb200 as R  # noqa: E402
import scenes as S  # noqa: E402
from oracle.oracle_lib import load_oracle  # noqa: E402
from ref_pin_common import EFRAME_CASES, ref_emode_frame  # noqa: E402

OUT = HERE / "ref_eframe_vectors.npz"


def main():
    lib = load_oracle()
    out = {}
    for name, (scene_fn, W, H, probe) in EFRAME_CASES.items():
        sc = scene_fn()
        o = R.RenderContext(lib)
        S.upload(o, sc, W, H)
        img, pr = ref_emode_frame(sc, o, W, H, probe=probe)
        out[f"{name}_image"] = img
        out[f"{name}_probe"] = np.array(pr, np.float64)
        print(name, img.shape, "mean", float(img[..., :3].mean()), "hit fraction", float((img[..., 3] > 0).mean()), "probe", pr)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
