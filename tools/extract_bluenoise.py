#!/usr/bin/env python3
"""Extract the 256-spp blue-noise tables (Heitz et al. 2019 supplemental data) that the
reference ships as three uint64 arrays in
RFW/system/context/rfw/context/blue_noise.h:5,1646,4925 and write them as one
327,680-byte file: sobol[65536] | scrambling[131072] | ranking[131072], one byte per entry,
in the little-endian byte order `createBlueNoiseBuffer` (blue_noise.h:8204-8219) reads them.

Runs only where /root/reference exists (the build container); the output
rendering-fw_b200/data/bluenoise_256spp.bin is committed because the sampler must be
identical to the reference's for identical seeds.
"""
import re
import struct
import sys
from pathlib import Path

REF = Path("/root/reference/RFW/system/context/rfw/context/blue_noise.h")
OUT = Path(__file__).resolve().parent.parent / "rendering-fw_b200" / "data" / "bluenoise_256spp.bin"


def main() -> int:
    if not REF.exists():
        print("reference not present; keeping committed table", file=sys.stderr)
        return 0
    src = REF.read_text()
    out = bytearray()
    for name, n in (("sob256_64", 8192), ("scr256_64", 16384), ("rnk256_64", 16384)):
        m = re.search(name + r"\[\d+\]\s*=\s*\{(.*?)\};", src, re.S)
        vals = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", m.group(1))]
        assert len(vals) == n, (name, len(vals))
        out += struct.pack("<%dQ" % n, *vals)
    assert len(out) == 327680
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_bytes(bytes(out))
    print("wrote", OUT, len(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
