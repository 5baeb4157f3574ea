"""tools/warp_stats.py -- divergence statistics of the raycast kernel WITHOUT a GPU: the device code compiled for the host
(tests/emu, test infrastructure) replays the 32 lanes of every warp in lockstep and counts, per warp step, which table
levels the warp touches.  Used to size kernel changes before GPU time is spent on them.

  python tools/warp_stats.py [--scene sphere2048|sphere256|cube|icosahedron] [--width W --height H] [--warp-w 4]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="sphere2048")
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--warp-w", type=int, nargs="*", default=[4])
    ap.add_argument("--mode", type=int, default=0)
    a = ap.parse_args()
    import bench
    import emu_ffi as E
    import woxel_b200 as W
    bench.WIDTH, bench.HEIGHT = a.width, a.height
    t0 = time.time()
    _, flat, what, tm = bench.build_scene(a.scene)
    eye, target = bench.camera_for(a.scene, 0)
    st = W.ComputeState.build(W.Camera(eye=eye, target=target, aspect=a.width / a.height), a.width, W.RenderMode(a.mode))
    for ww in a.warp_w:
        t1 = time.time()
        _, aov, s = E.render(flat.desc, st, a.width, a.height, aov=True, warp_w=ww, stats=True)
        ws = s["warp_steps"]
        # warp slots a CTA of four warps side by side keeps busy: a slot is held until the CTA's slowest warp ends
        import numpy as np
        wh = 32 // ww
        it = aov["iters"][0].astype(np.int64) + (aov["state"][0] != 2)  # lookups per primary ray
        hh, wd = (a.height // wh) * wh, (a.width // (4 * ww)) * (4 * ww)
        per_warp = it[:hh, :wd].reshape(hh // wh, wh, wd // ww, ww).max(axis=(1, 3))
        per_cta = per_warp.reshape(per_warp.shape[0], -1, 4)
        cta_slot_use = float(per_cta.sum() / (4 * per_cta.max(-1).sum()))
        out = {
            "scene": what, "frame": [a.width, a.height], "mode": a.mode, "warp_footprint": [ww, 32 // ww],
            "rays": s["rays"], "warps": s["warps"], "lane_steps": s["lane_steps"], "warp_steps": ws,
            "steps_per_ray": round(s["lane_steps"] / s["rays"], 2), "warp_steps_per_warp": round(ws / s["warps"], 2),
            "lanes_active_per_warp_step": round(s["lane_steps"] / ws, 2),
            "per_warp_step": {k: round(s[k] / ws, 4) for k in ("root_blocks", "n5_blocks", "n4_blocks", "leaf_blocks", "generic_iters")},
            "level_sets_touched": {"".join(n for b, n in ((1, "N5 "), (2, "N4 "), (4, "leaf ")) if i & b).strip() or "none":
                                   round(s[f"combo{i}"] / ws, 4) for i in range(8)},
            "table_reads_per_lane_step": round(s["lane_table_reads"] / s["lane_steps"], 3),
            "warp_slot_use_inside_4_warp_cta": round(cta_slot_use, 4),
            "emu_s": round(time.time() - t1, 1), "scene_s": round(t1 - t0, 1),
        }
        print(json.dumps(out))


if __name__ == "__main__":
    main()
