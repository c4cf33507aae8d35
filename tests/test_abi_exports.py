"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol include/swcgpu.h
declares; status codes map onto the Swift error enums; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from swcompression_b200 import _lib, errors


def test_library_builds_and_exports_every_declared_symbol():
    _lib.build()
    L = C.CDLL(_lib.SO_PATH)
    syms = _lib.declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_status_names_match_swift_enums():
    hdr = open(os.path.join(os.path.dirname(_lib.HEADER), "swc_status.h")).read()
    codes = {int(v): k for k, v in re.findall(r"(SWC_[A-Z0-9_]+)\s*=\s*(\d+)", hdr)}
    L = _lib.lib()
    for code in codes:
        name = L.swc_status_name(code).decode()
        assert name != "unknown", codes[code]
        if code >= 100:
            enum, case = name.split(".")
            exc = errors.error_for(code)
            assert type(exc).__name__ == enum and exc.case == case


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from swcompression_b200 import Deflate, EngineError
    with pytest.raises(EngineError) as e:
        Deflate.decompress(b"\x03\x00")
    assert e.value.case == "noDevice"


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", _lib.SO_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "swco_" not in out
    src_dir = os.path.join(os.path.dirname(_lib.SO_PATH), "csrc")
    for f in os.listdir(src_dir):
        if f.endswith((".cu", ".cuh", ".h", ".cpp")):
            txt = open(os.path.join(src_dir, f)).read()
            assert "swco" not in txt and "oracle/" not in txt, f
