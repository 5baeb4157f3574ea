"""How much of a frame depends on float contraction?

WGSL leaves fused multiply-add contraction to the driver, and the reference shader cannot be executed here
(DESIGN.md section 2), so the arithmetic of the real wgpu path is not pinned.  The oracle is strict IEEE without
contraction -- and the CUDA path is bit-identical to it.  This test bounds what the other legal choice changes: the
same restatement compiled with -ffp-contract=fast -mfma (oracle/libwxo_fma.so) is rendered beside the strict build
and compared with the north-star bar (hit voxel + leaf index >= 99.9 % of pixels, depth within 1e-4 relative on the
agreeing hits, RGB within 1/255 on >= 99.9 % of pixels).  The mismatching pixels are listed; they are grazing-angle
DDA ties (of the primary ray, or of a shadow ray in the Diffuse mode).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_ffi as O
import scenes

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [("cube", "default", 0), ("cube", "oblique_a", 3), ("icosahedron", "default", 3), ("icosahedron", "oblique_b", 0)]
W_, H_ = 960, 540

_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import scenes
out = {}
for name, cam, mode in %r:
    s = scenes.get_scene(name)
    eye, target = scenes.CAMERAS[cam]
    rgba, aov, st = s.gpu.render(scenes.state_for(eye, target, %d, %d, mode=mode), %d, %d)
    for k in ("state", "voxel", "leaf", "depth", "iters"):
        out[f"{name}.{cam}.{mode}.{k}"] = aov[k]
    out[f"{name}.{cam}.{mode}.rgba"] = rgba
np.savez(sys.argv[2], **out)
""" % (CASES, W_, H_, W_, H_)


def _has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


@pytest.mark.skipif(not _has_fma(), reason="host CPU has no FMA unit")
def test_fma_contraction_stays_within_the_north_star_bar(tmp_path):
    out = tmp_path / "fma.npz"
    env = dict(os.environ, WXO_VARIANT="fma")
    r = subprocess.run([sys.executable, "-c", _CHILD, HERE, str(out)], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    fma = np.load(out)
    report = {}
    for name, cam, mode in CASES:
        s = scenes.get_scene(name)
        eye, target = scenes.CAMERAS[cam]
        rgba, aov, st = s.gpu.render(scenes.state_for(eye, target, W_, H_, mode=mode), W_, H_)
        key = f"{name}.{cam}.{mode}"
        # hit voxel + leaf index where the ray hits; for rays that leave the world only the outcome is compared (their
        # exit position after a 4000-voxel flight is not a hit voxel)
        hit = aov["state"] == 0
        same = (aov["state"] == fma[key + ".state"]) & (~hit | ((aov["voxel"] == fma[key + ".voxel"]).all(-1) & (aov["leaf"] == fma[key + ".leaf"])))
        d0, d1 = aov["depth"].astype(np.float64), fma[key + ".depth"].astype(np.float64)
        ok = same & hit & np.isfinite(d0) & np.isfinite(d1) & (d0 != 0)  # depth of the hit point
        rel = float((np.abs(d1[ok] - d0[ok]) / d0[ok]).max()) if ok.any() else 0.0
        # colour: a secondary (shadow / reflection) ray can flip its outcome at a grazing tie too, so RGB is held to the
        # same 99.9 % as the hit itself
        rgb_ok = np.abs(rgba.astype(int) - fma[key + ".rgba"].astype(int)).max(-1) <= 1
        rgb = float(rgb_ok.mean())
        bad = np.argwhere(~same)
        report[key] = {"agreement": float(same.mean()), "depth_rel_max": rel, "rgb_within_1_of_255": rgb,
                       "iteration_count_differs": float((aov["iters"] != fma[key + ".iters"]).mean()),
                       "mismatching_pixels_yx": bad[:16].tolist(), "n_mismatch": int(len(bad))}
        assert same.mean() >= 0.999, report[key]
        assert rel <= 1e-4, report[key]
        assert rgb >= 0.999, report[key]
    print(json.dumps(report, indent=1))


_CHILD_SPARSE = r"""
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import scenes
out = {}
for name in ("offcentre_sphere", "single_voxel"):
    s = scenes.get_scene(name)
    rgba, aov, _ = s.gpu.render(scenes.state_for(*scenes.CAMERAS["oblique_a"], 384, 216, mode=0), 384, 216)
    out[name + ".rgba"] = rgba
    for k in ("state", "voxel", "leaf", "depth"):
        out[name + "." + k] = aov[k]
np.savez(sys.argv[2], **out)
"""


@pytest.mark.skipif(not _has_fma(), reason="host CPU has no FMA unit")
def test_sparse_scenes_are_contraction_sensitive_in_the_reference_itself(tmp_path):
    """Round 2 finding.  In a tree that does not cover the world (missing N5s) the empty space is stepped through in 4096-voxel
    lattice cells whose planes COINCIDE with the +-4096 world boundary, where `any(4096 < |p|)` ends a ray: whether a ray ends on
    this step or the next, and with which axis mask (= which out-of-bounds shade), flips with one ulp of p.  The ORACLE ITSELF,
    compiled with FMA contraction, differs from its strict build on 4-7 % of the pixels of such scenes (0.003 % on the assets,
    the test above) -- so does any contracting WGSL compiler, and so does the library's tolerance mode (WX_OPT_MARCH >= 1).  The
    hits agree; what differs is the shade of out-of-bounds pixels.  This is a property of the reference, recorded here."""
    import agreement
    out_fma, out_strict = tmp_path / "fma.npz", tmp_path / "strict.npz"
    for variant, out in (("fma", out_fma), ("", out_strict)):
        env = dict(os.environ)
        if variant:
            env["WXO_VARIANT"] = variant
        else:
            env.pop("WXO_VARIANT", None)
        r = subprocess.run([sys.executable, "-c", _CHILD_SPARSE, HERE, str(out)], env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
    a, b = np.load(out_fma), np.load(out_strict)
    for name in ("offcentre_sphere", "single_voxel"):
        fig = agreement.compare(a[name + ".rgba"], {k: a[f"{name}.{k}"] for k in ("state", "voxel", "leaf", "depth")},
                                b[name + ".rgba"], {k: b[f"{name}.{k}"] for k in ("state", "voxel", "leaf", "depth")}, max_list=0)
        assert 0.85 < fig["voxel_leaf_agree_frac"] < 0.99, (name, fig)       # far below the 99.9 % bar ...
        assert (a[name + ".state"] == b[name + ".state"]).mean() > 0.9999  # ... yet every ray ends in the same state
        hit = b[name + ".state"] == 0
        if hit.any():
            assert (a[name + ".voxel"][hit] == b[name + ".voxel"][hit]).all(-1).mean() >= 0.999
