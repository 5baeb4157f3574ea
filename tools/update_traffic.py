"""tools/update_traffic.py -- refresh profiles/traffic.json from one `ncu --set full` capture of the raycast kernel.

  python tools/update_traffic.py gpurun_out/prof_X.ncu-rep [--scene sphere2048] [--lib woxel_b200/libwoxel_b200.so] [--note "..."]

traffic.json holds the per-launch figures bench.py cannot measure itself (it never runs under a profiler): DRAM / L2 / L1 bytes
and warp instructions of wx::raycast_kernel<0,false>.  They are constants of ONE build, so the file also stores the md5 of that
kernel's SASS (tools/sass_diff.py); bench.py recomputes it for the library it has loaded and reports `traffic_stale: true` when
they differ -- the numbers then describe an older kernel."""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import sass_diff  # noqa: E402

KERNEL = "_ZN2wx14raycast_kernelILi0ELb0EEEvNS_12RenderParamsE"


def kernel_md5(lib: str):
    k = sass_diff.kernels(lib).get(KERNEL)
    return k[0] if k else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--scene", default="sphere2048")
    ap.add_argument("--lib", default=os.path.join(ROOT, "woxel_b200", "libwoxel_b200.so"))
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u, v = rows[0], rows[1], rows[2]
    val = {n: (float(v[i]), u[i]) for i, n in enumerate(h) if v[i].replace(".", "", 1).replace("-", "", 1).isdigit()}

    def mbytes(name):
        x, unit = val[name]
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]

    path = os.path.join(ROOT, "profiles", "traffic.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    old = doc.get(a.scene, {})
    doc[a.scene] = {
        "kernel": v[h.index("Kernel Name")],
        "sass_md5": kernel_md5(a.lib),
        "capture": os.path.basename(a.rep),
        "gpu_time_us": val["gpu__time_duration.sum"][0],
        "dram_bytes_per_launch": int(mbytes("dram__bytes_read.sum") + mbytes("dram__bytes_write.sum")),
        "warp_instructions_per_launch": int(val["smsp__inst_executed.sum"][0]),
        "l2_bytes_per_launch": int(val["lts__t_sectors.sum"][0] * 32),
        "l1_bytes_per_launch": int(val["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"][0] * 32),
        "l2_peak_gbs": old.get("l2_peak_gbs"),
        "l2_peak_source": old.get("l2_peak_source"),
        "source": f"one launch under `ncu --set full --clock-control none` ({os.path.basename(a.rep)}): dram__bytes_read.sum + dram__bytes_write.sum, "
                  f"lts__t_sectors.sum x 32 B, l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum x 32 B, smsp__inst_executed.sum" + (". " + a.note if a.note else ""),
    }
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps(doc[a.scene], indent=1))


if __name__ == "__main__":
    main()
