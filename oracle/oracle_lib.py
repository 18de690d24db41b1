"""Loader of the CPU oracle (test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
reference arm may import this).  The oracle exports the call shapes of include/rfwb200.h under the prefix ``rfworacle_``,
so the package's ctypes wrapper (rendering-fw_b200/python/rfwb200.py) drives it unchanged; the product package itself
holds no reference to anything under oracle/."""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REPO_DIR = ORACLE_DIR.parent
sys.path.insert(0, str(REPO_DIR / "rendering-fw_b200" / "python"))

import rfwb200 as R  # noqa: E402

ORACLE_LIB = ORACLE_DIR / "librfworacle.so"
ORACLE_FAST_LIB = ORACLE_DIR / "librfworacle_fast.so"


def _install_blue_noise(ctx) -> None:
    # the product embeds the table in its library; the oracle is handed the same bytes
    table = np.fromfile(R.BLUENOISE_BIN, dtype=np.uint8)
    ctx._check(ctx.L.fn("set_blue_noise", C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t])(ctx._h, table.ctypes.data, table.size))


def load_oracle(fast: bool = False) -> "R.Library":
    lib = R.Library(ORACLE_FAST_LIB if fast else ORACLE_LIB, "rfworacle_")
    lib.post_create = _install_blue_noise
    return lib
