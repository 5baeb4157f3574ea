"""Render a .vdb grid to PNG files through the product path (reader -> wx_tree_build -> wx_render -> wx_capture_srgb ->
wxh_write_png): the reference's `cargo run` + screenshot, headless.  Needs a GPU.

  python tools/render_vdb.py assets/cube.vdb ls_cube out/cube [--size 1920 1080] [--modes 0 2 3 4] [--eye X Y Z] [--target X Y Z]
                             [--orbit N] [--grid]

Writes out/cube_mode3.png ... (with --orbit N: out/cube_mode3_000.png ... one frame per camera of an orbit around +y through
`eye`, rendered as ONE camera batch).  Prints one JSON line per mode with the kernel time."""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import woxel_b200 as W  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("path")
ap.add_argument("grid")
ap.add_argument("out")
ap.add_argument("--size", type=int, nargs=2, default=(1920, 1080))
ap.add_argument("--modes", type=int, nargs="+", default=[3])
ap.add_argument("--eye", type=float, nargs=3, default=(0.5, 0.5, -500.5))      # camera.rs:16-29
ap.add_argument("--target", type=float, nargs=3, default=(0.5, 0.5, 0.5))
ap.add_argument("--orbit", type=int, default=0)
ap.add_argument("--grid-lines", action="store_true", help="show_345: highlight N5 / N4 / leaf boundaries (the shader reads it in modes 0-2 only)")
a = ap.parse_args()
w, h = a.size

v = W.VdbReader(a.path).read_vdb345_grid(a.grid)
n5, n4, n3 = v.count_nodes()
ctx = W.Context()
tree = ctx.build(v.to_flat(narrow_leaves=False))  # compute_sdf on the GPU + device tables
os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)

eyes = [tuple(a.eye)]
if a.orbit > 1:
    c = a.target
    dx, dz = a.eye[0] - c[0], a.eye[2] - c[2]
    r, th0 = math.hypot(dx, dz), math.atan2(dx, -dz)
    eyes = [(c[0] + r * math.sin(th0 + 2 * math.pi * k / a.orbit), a.eye[1], c[2] - r * math.cos(th0 + 2 * math.pi * k / a.orbit))
            for k in range(a.orbit)]
for mode in a.modes:
    states = [W.ComputeState.build(W.Camera(eye=e, target=tuple(a.target), aspect=w / h), w, W.RenderMode(mode),
                                   show_grid=(a.grid_lines,) * 3) for e in eyes]
    ctx.render(tree, states, w, h)
    info = ctx.last_render_info()
    rgb = ctx.capture_srgb(len(states), w, h)
    names = []
    for k in range(len(states)):
        name = f"{a.out}_mode{mode}" + (f"_{k:03d}" if len(states) > 1 else "") + ".png"
        W.write_png(name, rgb[k])
        names.append(name)
    print(json.dumps({"grid": a.grid, "nodes": [n5, n4, n3], "mode": mode, "frames": len(states), "size": [w, h],
                      "kernel_ms": round(info.kernel_ms, 4), "sdf_gpu_ms": round(tree.sdf.device_ms, 2), "first": names[0]}), flush=True)
tree.free()
ctx.close()
