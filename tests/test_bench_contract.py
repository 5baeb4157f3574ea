"""bench.py's reference arm runs on the CPU alone (it is what the driver times beside the GPU arm): one JSON line with
the contract's keys.  The GPU arm's line is checked on the GPU box by the driver; here only its argument handling."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # the debug-size scene keeps this to a few seconds; the driver runs the default (2048^3 sphere, ~1 s per step)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--scene", "sphere256"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] == 3  # W >= 3 as in the GPU arm, never capped: the driver's warmup_match
    # the two arms print the same `config` (the driver's same_config)
    import bench
    assert d["config"] == bench.workload_config(d["config"]["workload"].split(", 3840x2160")[0])


def test_traffic_json_carries_the_kernel_hash():
    """profiles/traffic.json describes ONE build of the raycast kernel: it stores that kernel's SASS md5, bench.py compares it
    with the loaded library's and says `traffic_stale` when they differ.  Here: the committed figures belong to the committed kernel."""
    import json
    import shutil
    import pytest
    import bench
    prof = json.load(open(os.path.join(bench.ROOT, "profiles", "traffic.json")))["sphere2048"]
    assert len(prof["sass_md5"]) == 32 and prof["warp_instructions_per_launch"] > 1e8
    if shutil.which("cuobjdump") is None:
        pytest.skip("no cuobjdump")
    assert bench.sass_md5_of_loaded_kernel() == prof["sass_md5"], "profiles/traffic.json is stale: re-capture (tools/update_traffic.py)"


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_measured_bytes_per_ray_helper():
    """The reporting extra of the roofline object: computed from profiles/traffic.json, None (never an exception) on odd input."""
    import json
    import os
    import bench
    prof = json.load(open(os.path.join(bench.ROOT, "profiles", "traffic.json")))["sphere2048"]
    m = bench.measured_bytes_per_ray(prof, 3840 * 2160, 6541.8)
    assert m["l1"] > m["l2"] > m["dram"] > 0 and m["rays_per_s_roofline_Grays"]["l2"] > 100 and m["rays_per_s_roofline_Grays"]["dram"] > 100
    for bad in (None, {}, {"l2_bytes_per_launch": "x"}, {"l2_bytes_per_launch": 5}):
        assert bench.measured_bytes_per_ray(bad, 10, 6541.8) is None or isinstance(bench.measured_bytes_per_ray(bad, 10, 6541.8), dict)
    assert bench.measured_bytes_per_ray(prof, 0, 6541.8) is None
