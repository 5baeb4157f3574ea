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


def test_write_png_decodes_to_the_same_pixels(tmp_path):
    """The host's PNG writer (frame_dump.cpp): checked chunk by chunk here and, when Pillow is importable, by a real decoder."""
    import struct
    import zlib
    rng = np.random.default_rng(3)
    for (h, w) in ((5, 7), (64, 129), (1, 1)):
        rgb = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        p = tmp_path / f"f{h}x{w}.png"
        W.write_png(str(p), rgb)
        raw = p.read_bytes()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(raw):
            n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
            data = raw[pos + 8:pos + 8 + n]
            (crc,) = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
            assert crc == zlib.crc32(typ + data)
            chunks.append((typ, data))
            pos += 12 + n
        assert [c[0] for c in chunks] == [b"IHDR", b"sRGB", b"IDAT", b"IEND"]
        assert struct.unpack(">IIBBBBB", chunks[0][1]) == (w, h, 8, 2, 0, 0, 0)
        rows = np.frombuffer(zlib.decompress(chunks[2][1]), np.uint8).reshape(h, 1 + 3 * w)
        assert not rows[:, 0].any() and np.array_equal(rows[:, 1:].reshape(h, w, 3), rgb)
        try:
            from PIL import Image
        except ImportError:
            continue
        assert np.array_equal(np.asarray(Image.open(str(p)).convert("RGB")), rgb)


def test_frame_dump_errors_are_statuses(tmp_path):
    rgb = np.zeros((4, 4, 3), np.uint8)
    with pytest.raises(W.WxError):
        W.write_png(str(tmp_path / "no_such_dir" / "f.png"), rgb)
    with pytest.raises(ValueError):
        W.write_ppm(str(tmp_path / "f.ppm"), np.zeros((4, 4), np.uint8))


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
