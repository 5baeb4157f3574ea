"""The north-star acceptance bar as a function (used by the tolerance-mode tests, bench.py and tools): a frame + AOVs against
the oracle's -- hit voxel coordinate and leaf index equal on >= 99.9 % of the pixels with the mismatches listed, depth within
1e-4 relative, RGB within 1/255 (BASELINE.json north_star)."""
from __future__ import annotations

import numpy as np


def compare(rgba, aov, ref_rgba, ref_aov, max_list: int = 32) -> dict:
    """rgba [H,W,4] u8, aov dict of [H,W,...] arrays (state, voxel, leaf, depth); ref_* the oracle's.  Returns the figures."""
    st, rst = aov["state"], ref_aov["state"]
    h, w = st.shape
    # only the pixels the reference dispatches (wgpu_context.rs:281): the fringe is never written by either side
    disp = np.zeros((h, w), bool)
    disp[: h // 4 * 4, : w // 8 * 8] = True
    npix = int(disp.sum())
    state_eq = (st == rst) | ~disp
    hit = rst == 0
    both_hit = hit & (st == 0)
    vox_eq = (aov["voxel"] == ref_aov["voxel"]).all(-1)
    leaf_eq = aov["leaf"] == ref_aov["leaf"]
    # a pixel agrees when the ray ends in the same state and, for a hit, in the same voxel of the same leaf
    hit = hit & disp
    agree = state_eq & (~hit | (vox_eq & leaf_eq))
    d, rd = aov["depth"].astype(np.float64), ref_aov["depth"].astype(np.float64)
    rel = np.zeros_like(d)
    np.divide(np.abs(d - rd), np.abs(rd), out=rel, where=both_hit & (rd != 0))
    rgb_diff = np.abs(rgba[..., :3].astype(np.int16) - ref_rgba[..., :3].astype(np.int16)).max(-1) * disp
    # a pixel whose colour differs by more than 1/255 counts as a mismatch too (same voxel entered through another face: a DDA tie)
    agree = (agree & (rgb_diff <= 1)) | ~disp
    bad = np.argwhere(~agree)
    listed = [{"y": int(y), "x": int(x), "state": [int(st[y, x]), int(rst[y, x])], "voxel": [aov["voxel"][y, x].tolist(), ref_aov["voxel"][y, x].tolist()],
               "leaf": [int(aov["leaf"][y, x]), int(ref_aov["leaf"][y, x])], "rgb": [rgba[y, x, :3].tolist(), ref_rgba[y, x, :3].tolist()]}
              for y, x in bad[:max_list]]
    return {
        "pixels": npix,
        "voxel_leaf_agree_frac": float((agree & disp).sum()) / npix,
        "mismatch_pixels": int((~agree).sum()),
        "depth_rel_max_on_agreeing_hits": float(rel[agree & both_hit].max()) if (agree & both_hit).any() else 0.0,
        "depth_rel_over_1e-4_pixels": int(((rel > 1e-4) & agree).sum()),
        "rgb_within_1_frac": float(((rgb_diff <= 1) & disp).sum()) / npix,
        "rgb_over_1_pixels": int((rgb_diff > 1).sum()),
        "mismatches_listed": listed,
    }


def meets_bar(fig: dict) -> bool:
    return fig["voxel_leaf_agree_frac"] >= 0.999 and fig["depth_rel_over_1e-4_pixels"] == 0
