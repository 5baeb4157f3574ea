"""The boundary from plain C: tests/c/abi_smoke.c includes both public headers as C99 (-pedantic -Werror), checks the
ComputeState layout and drives set_voxel -> to_flat -> wx_tree_build -> wx_render -> wx_capture_srgb without Python."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_smoke.c")
OUT = os.path.join(ROOT, "tests", "c", "build", "abi_smoke")


def build_binary():
    import __graft_entry__ as g
    g.build()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    lib_dir = os.path.join(ROOT, "woxel_b200")
    cmd = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), SRC, "-o", OUT,
           "-L", lib_dir, "-lwoxel_host", "-lwoxel_b200", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return OUT


def run_binary():
    return subprocess.run([build_binary()], capture_output=True, text=True, timeout=300)


def test_headers_are_c99_and_refuse_to_run_without_a_gpu():
    r = run_binary()
    assert r.returncode == 0, r.stdout + r.stderr
    import torch
    if not torch.cuda.is_available():
        assert r.stdout.strip() == "no-device"  # no CPU fallback: wx_init says WX_ERR_NO_DEVICE
    else:
        assert r.stdout.startswith("rendered ")


@pytest.mark.gpu
def test_c_host_renders_through_the_abi():
    r = run_binary()
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("rendered 128x64"), r.stdout
