"""The C-ABI library loads and exports every symbol include/audiort.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import pytest

from audio_raytracer_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    hdr = open(os.path.join(ROOT, "include", "audiort.h")).read()
    return sorted(set(re.findall(r"ART_API\s+[\w\s\*]+?\b(art_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(native.EXPORTS)


def test_library_exports_every_declared_symbol(art_lib):
    for name in _declared_functions():
        assert hasattr(art_lib, name), f"{name} is declared in include/audiort.h but not exported"


def test_wire_struct_sizes():
    from audio_raytracer_b200.layouts import AABB_DT, OBB_DT, SPHERE_DT, SETTINGS_DT
    assert (AABB_DT.itemsize, OBB_DT.itemsize, SPHERE_DT.itemsize, SETTINGS_DT.itemsize) == (20, 26, 16, 24)
    assert C.sizeof(native.ArtConfig) == 4 * (4 + native.ART_MAX_DEVICES + 1 + 3)
    assert C.sizeof(native.ArtOutputs) == 9 * C.sizeof(C.c_void_p)


def test_no_cpu_fallback_when_no_device(art_lib):
    """Without a CUDA device art_create must fail loudly (ART_E_NO_DEVICE), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    cfg = native.ArtConfig(abiVersion=native.ART_ABI_VERSION, device=0, flags=0)
    ctx = C.c_void_p()
    rc = art_lib.art_create(C.byref(cfg), C.byref(ctx))
    assert rc == native.ART_E_NO_DEVICE and not ctx.value
    assert b"no CPU fallback" in art_lib.art_last_error(None)
    with pytest.raises(native.ArtError):
        native.Context(0)


def test_bad_abi_version_rejected(art_lib):
    cfg = native.ArtConfig(abiVersion=999, device=0, flags=0)
    ctx = C.c_void_p()
    assert art_lib.art_create(C.byref(cfg), C.byref(ctx)) == native.ART_E_ARG


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (tier rule 3)."""
    pkg = os.path.join(ROOT, "audio-raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libaudiort_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
                assert not re.search(r'#include\s+[<"][^>"]*oracle', text), f
