#!/usr/bin/env python3
"""Run the pin scene through the REFERENCE's own CUDA wavefront kernels (RFW/backends/CUDART/src/Kernels.cu compiled for the
host by oracle/ref_build -> oracle/_ref/librfwref_kernels.so: generatePrimaryRay, intersect_rays x3, shade_rays, the counters
kernels, under the bounce loop of Context.cpp:83-159) and commit the accumulators, the camera-ray buffers and the queue
sizes as tests/golden/ref_kernels_vectors.npz.  tests/test_ref_pin.py checks the oracle's whole path-tracing pipeline against
them everywhere, and the stored vectors against the live library where it exists.  Runs only in the build container."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(REPO / "rendering-fw_b200" / "python"))
sys.path.insert(0, str(REPO / "tests"))
import rfwb200 as R  # noqa: E402
sys.path.insert(0, str(REPO))
from oracle.oracle_lib import load_oracle  # noqa: E402
import scenes as S  # noqa: E402
from ref_pin_common import pin_cases, pin_scene, pin_view14, reference_kernels_render  # noqa: E402

OUT = Path(__file__).resolve().parent / "ref_kernels_vectors.npz"


def main():
    out = {}
    for name, (w, h, first, count, aperture) in pin_cases().items():
        sc = pin_scene(rich=(name in ("rich", "lights")), lights=(name == "lights"))
        o = R.RenderContext(load_oracle())  # only to build and export the MBVHs in the reference's node layout
        S.upload(o, sc, w, h)
        ref = reference_kernels_render(o, sc, pin_view14(sc, w, h, aperture), w, h, first, count)
        out[name + "_acc"] = ref["acc"]
        out[name + "_origins"], out[name + "_directions"], out[name + "_states"] = ref["origins"], ref["directions"], ref["states"]
        out[name + "_counters"] = ref["counters"]
        print(name, "mean", ref["acc"][..., :3].mean() / count, "queues of the first sample", ref["counters"][0, :3].tolist())
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, OUT.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
