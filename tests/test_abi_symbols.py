"""The C-ABI libraries load and export every symbol include/*.h declares (no compute calls, no GPU needed)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header: str, prefix: str):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    from woxel_b200 import _ffi
    names = declared_functions("woxel_b200.h", "wx_")
    assert len(names) >= 20
    lib = C.CDLL(_ffi.CUDA_LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/woxel_b200.h but not exported"
    assert sorted(_ffi.CUDA_API) == names, "ctypes table and header are out of sync"
    assert _ffi.cuda_lib().wx_abi_version() == 2


def test_host_library_exports_every_declared_symbol():
    from woxel_b200 import _ffi
    names = declared_functions("woxel_host.h", "wxh_")
    lib = _ffi.host_lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/woxel_host.h but not exported"
    assert sorted(_ffi.HOST_API) == names


def test_struct_layouts_match_the_reference_uniform():
    from woxel_b200 import _ffi
    s = _ffi.WxState
    # offsets of ComputeState (compute_state.rs:9-29 / raycast.comp.wgsl:1-23)
    assert (s.view_proj.offset, s.camera_to_world.offset, s.eye.offset, s.u.offset, s.mv.offset, s.wp.offset,
            s.render_mode.offset, s.show_345.offset, s.sun_dir.offset, s.sun_color.offset) == (0, 64, 128, 144, 160, 176, 192, 208, 224, 240)
    assert C.sizeof(s) == 256


def test_no_cpu_fallback():
    """Without a CUDA device the product refuses to run instead of silently rendering on the CPU."""
    import woxel_b200 as W
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(W.WxError) as e:
        W.Context()
    assert e.value.status == -2  # WX_ERR_NO_DEVICE
    # and nothing in the product package references the oracle
    pkg = os.path.join(ROOT, "woxel_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_ffi" not in text and "libwxo" not in text and "wxo_" not in text, f
                assert "emu_ffi" not in text and "libwx_emu" not in text and "wxe_" not in text, f


def test_product_libraries_carry_no_emulation_code():
    """tests/emu compiles wx_device.cuh for the host behind WX_HOST_EMU; the shipped libraries are built without it:
    no emulation or oracle symbol is defined or referenced by them."""
    import subprocess
    for so in ("libwoxel_b200.so", "libwoxel_host.so"):
        out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "woxel_b200", so)], capture_output=True, text=True, check=True).stdout
        und = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(ROOT, "woxel_b200", so)], capture_output=True, text=True, check=True).stdout
        for needle in ("wx_emu", "wxe_", "wxo_"):
            assert needle not in out and needle not in und, (so, needle)
