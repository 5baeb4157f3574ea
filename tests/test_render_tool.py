"""tools/render_vdb.py end to end on the GPU: a .vdb file (written by the test-side writer) -> reader -> wx_tree_build ->
wx_render (camera batch) -> wx_capture_srgb -> PNG files, as a user of the reference would dump frames."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import vdb_writer as V

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_render_vdb_tool_writes_frames(tmp_path):
    g = np.arange(-40, 41)
    x, y, z = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt((x + 0.5) ** 2 + (y + 0.5) ** 2 + (z + 0.5) ** 2)
    pts = np.stack([x, y, z], -1)[np.abs(r - 36.0) <= 2.0]
    path = tmp_path / "ball.vdb"
    V.VdbWriter(compression=V.ZIP | V.ACTIVE_MASK, half_float=True).write(str(path), pts, grid_name="ls_ball")
    out = tmp_path / "frames" / "ball"
    cmd = [sys.executable, os.path.join(ROOT, "tools", "render_vdb.py"), str(path), "ls_ball", str(out), "--size", "320", "200",
           "--modes", "0", "3", "--eye", "0.5", "0.5", "-150.5", "--orbit", "3"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    lines = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert [l["mode"] for l in lines] == [0, 3] and all(l["frames"] == 3 and l["kernel_ms"] > 0 for l in lines)
    from PIL import Image
    for mode in (0, 3):
        frames = [np.asarray(Image.open(f"{out}_mode{mode}_{k:03d}.png").convert("RGB")) for k in range(3)]
        for f in frames:
            assert f.shape == (200, 320, 3)
            centre, corner = f[100, 160].astype(int), f[2, 2].astype(int)
            assert not np.array_equal(centre, corner)  # the ball is in the middle of every orbit frame
        if mode == 0:  # Gray mode shades by the face normal only: the ball looks the same from the three orbit positions up to symmetry
            assert len({f.tobytes() for f in frames}) >= 1
