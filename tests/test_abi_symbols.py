"""The C-ABI library loads and exports exactly what include/rfwb200.h declares (no compute calls: no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest

import rfwb200 as R

REPO = Path(__file__).resolve().parent.parent
HEADER = (REPO / "include" / "rfwb200.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"RFWB200_API\s+[\w\s\*]+?\b(rfwb200_\w+)\s*\(", HEADER)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    # every virtual of rfw::RenderContext (context.h:78-110) has a counterpart
    for required in ("create", "destroy", "init", "cleanup", "render_frame", "set_materials", "set_textures", "set_mesh",
                     "set_instance", "set_sky", "set_lights", "get_probe_results", "get_settings", "set_setting", "update",
                     "set_probe_index", "get_stats"):
        assert f"rfwb200_{required}" in names, required
    assert len(names) >= 30


def test_library_exports_every_declared_symbol(product_lib):
    lib = ctypes.CDLL(str(R.PRODUCT_LIB))
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing


def test_no_undeclared_exports(product_lib):
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", str(R.PRODUCT_LIB)], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    extra = {e for e in exported if not e.startswith("rfwb200_") and not e.startswith("_")}
    assert not extra, extra
    assert {e for e in exported if e.startswith("rfwb200_")} == set(declared_symbols())


def test_version_and_error_string_without_gpu(product_lib):
    lib = ctypes.CDLL(str(R.PRODUCT_LIB))
    lib.rfwb200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.rfwb200_version()
    lib.rfwb200_last_error.restype = ctypes.c_char_p
    assert lib.rfwb200_last_error() is not None


def test_create_fails_loudly_without_a_device(product_lib):
    """No CPU fallback: on a box without a B200 the product refuses to create a context."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.Rfwb200Error) as e:
        R.RenderContext(product_lib)
    assert "no CUDA device" in str(e.value) or "fallback" in str(e.value)


def test_product_does_not_link_or_reference_the_oracle():
    """The oracle is test infrastructure: nothing under rendering-fw_b200/ may include or load it."""
    pkg = REPO / "rendering-fw_b200"
    for f in list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")) + list(pkg.rglob("Makefile")):
        text = f.read_text(errors="ignore")
        assert "rfw_oracle" not in text and "librfworacle" not in text and "oracle/" not in text, f
    import subprocess

    needed = subprocess.run(["readelf", "-d", str(R.PRODUCT_LIB)], capture_output=True, text=True).stdout
    assert "oracle" not in needed


def test_wire_format_sizes():
    assert R.TRIANGLE_DTYPE.itemsize == 160 and R.MATERIAL_DTYPE.itemsize == 192
    assert R.TRIANGLE_DTYPE.fields["material"][1] == 28 and R.TRIANGLE_DTYPE.fields["vertex0"][1] == 112
    assert R.MATERIAL_DTYPE.fields["tex0"][1] == 32 and R.MATERIAL_DTYPE.fields["amap"][1] == 176
    assert R.AREA_LIGHT_DTYPE.fields["vertex0"][1] == 48 and R.AREA_LIGHT_DTYPE.fields["inst_idx"][1] == 76
