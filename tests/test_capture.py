"""Capture step after the raycast: RGBA8 (linear) -> RGB8 through the reference's linear_to_srgb (recorder.rs:20-37, :132-140)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_ffi as O
import scenes
import woxel_b200 as W
from woxel_b200 import _ffi


def oracle_lut():
    lib = O.lib()
    lib.wxo_linear_to_srgb.restype = C.c_uint8
    lib.wxo_linear_to_srgb.argtypes = [C.c_uint8]
    return np.array([lib.wxo_linear_to_srgb(v) for v in range(256)], np.uint8)


def test_oracle_linear_to_srgb_known_values():
    lut = oracle_lut()
    assert lut[0] == 0 and lut[255] == 255
    assert lut[1] == 13  # 12.92 * (1/255) * 255 = 12.92 -> 13 (linear branch; the branch point 0.0031308 lies between 0 and 1/255)
    assert (np.diff(lut.astype(int)) >= 0).all()
    # independent evaluation in float64: equal everywhere except where the f32 result sits within 1e-4 of a rounding tie
    c = np.arange(256) / 255.0
    s = np.where(c <= 0.0031308, 12.92 * c, 1.055 * np.power(c, 1 / 2.4) - 0.055) * 255.0
    near_tie = np.abs(s - np.floor(s) - 0.5) < 1e-4
    assert (lut[~near_tie] == np.floor(s[~near_tie] + 0.5).astype(np.uint8)).all()
    assert not near_tie.any() or near_tie.sum() < 3


def test_product_table_equals_oracle():
    t = np.zeros(256, np.uint8)
    assert _ffi.cuda_lib().wx_srgb_table(t.ctypes.data) == 0  # host arithmetic, no device
    assert np.array_equal(t, oracle_lut())


def test_write_ppm_roundtrip(tmp_path):
    rgb = (np.arange(5 * 7 * 3) % 251).astype(np.uint8).reshape(5, 7, 3)
    p = tmp_path / "f.ppm"
    W.write_ppm(str(p), rgb)
    raw = p.read_bytes()
    assert raw.startswith(b"P6\n7 5\n255\n") and raw[len(b"P6\n7 5\n255\n"):] == rgb.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(256, 128, 1), (70, 30, 1), (101, 37, 3)])
def test_capture_srgb_matches_oracle(gpu_ctx, shape):
    w, h, n = shape
    s = scenes.get_scene("cube")
    tree = gpu_ctx.upload(s.desc())
    try:
        eye, target = scenes.CAMERAS["oblique_a"]
        sts = [W.ComputeState.from_buffer_copy(bytes(scenes.state_for(eye, target, w, h, mode=m))) for m in (3, 4, 2)[:n]]
        rgba, _ = gpu_ctx.render(tree, sts, w, h)
        rgb = gpu_ctx.capture_srgb(n, w, h)
        ref = oracle_lut()[rgba[..., :3]]
        assert rgb.shape == (n, h, w, 3) and np.array_equal(rgb, ref)
        with pytest.raises(W.WxError):
            gpu_ctx.capture_srgb(n, w + 1, h)  # not the frame that was rendered last
    finally:
        tree.free()
