"""compute_sdf on the 2048^3 sphere: GPU (wx_compute_sdf) vs the product's host sweep; full-size value check."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import woxel_b200 as W

scene = sys.argv[1] if len(sys.argv) > 1 else "sphere2048"
v = W.VDB345.sphere(half=1024, radius=1000.0, band=3.0) if scene == "sphere2048" else W.VDB345.torus()
ctx = W.Context()
flat = v.to_flat(narrow_leaves=False)
t0 = time.time(); info = flat.compute_sdf_gpu(ctx); gpu_wall = time.time() - t0
t0 = time.time(); info2 = flat.compute_sdf_gpu(ctx); gpu_wall2 = time.time() - t0
g5, g4, g3 = flat.tab5.copy(), flat.tab4.copy(), flat.tab3.copy()
t0 = time.time(); tb = ctx.build(flat); build_s = time.time() - t0; build_ms = tb.sdf.device_ms; tb.free()
t0 = time.time(); tu = ctx.upload(flat); upload_s = time.time() - t0; tu.free()
t0 = time.time(); v.compute_sdf(); host_s = time.time() - t0
f = v.to_flat(narrow_leaves=False)
bits = np.unpackbits(np.ascontiguousarray(f.vals3).view(np.uint8), bitorder="little").reshape(f.n3, 512).astype(bool)
same = bool(np.array_equal(g5, f.tab5) and np.array_equal(g4, f.tab4) and np.array_equal(g3[~bits], f.tab3[~bits].astype(g3.dtype)))
print(json.dumps({"scene": scene, "n5": f.n5, "n4": f.n4, "n3": f.n3, "gpu_device_ms": round(info2.device_ms, 2), "gpu_call_s_first": round(gpu_wall, 3),
                  "gpu_call_s": round(gpu_wall2, 3), "host_sweep_s": round(host_s, 2), "tree_build_call_s": round(build_s, 3), "tree_build_device_ms": round(build_ms, 2),
                  "tree_upload_call_s": round(upload_s, 3), "identical_to_host_sweep": same,
                  "max_dist": list(info2.max_dist), "leaf_elem_bytes": int(flat.desc.tab3_elem_bytes)}))
