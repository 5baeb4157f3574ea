"""Summarise an .ncu-rep (one kernel launch, --set full --import-source on) into a small JSON for profiles/.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/rN_name.json [--top 25]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    raw = ncu_csv(rep, "raw")
    h, u, v = raw[0], raw[1], raw[2]
    summary = {"report": rep, "kernel": v[h.index("Kernel Name")]}
    for i, n in enumerate(h):
        if n in KEYS:
            summary[n] = {"unit": u[i], "value": v[i]}
        if "average_warps_issue_stalled" in n and n.endswith("per_issue_active.ratio"):
            try:
                if float(v[i]) >= 0.1:
                    summary.setdefault("stalls_per_issue", {})[n.split("issue_stalled_")[1].split("_per_issue")[0]] = round(float(v[i]), 3)
            except ValueError:
                pass
    try:
        src = ncu_csv(rep, "source")
        hh = src[1]
        col = {n: i for i, n in enumerate(hh)}
        rows = src[2:]
        tot_i = sum(int(r[col["Instructions Executed"]]) for r in rows)
        tot_s = sum(int(r[col["# Samples"]]) for r in rows)
        hot = sorted(rows, key=lambda r: -int(r[col["# Samples"]]))[:top]
        summary["sass_instructions"] = len(rows)
        summary["warp_instructions_executed"] = tot_i
        summary["hot_instructions_by_samples"] = [
            {"sass": r[col["Source"]].strip(), "inst_M": round(int(r[col["Instructions Executed"]]) / 1e6, 2),
             "avg_threads": r[col["Avg. Threads Executed"]], "samples_pct": round(100 * int(r[col["# Samples"]]) / max(tot_s, 1), 2)}
            for r in hot]
    except Exception as e:  # source page needs -lineinfo / --import-source
        summary["source_page"] = f"unavailable: {e}"
    json.dump(summary, open(dst, "w"), indent=1)
    print("wrote", dst)


if __name__ == "__main__":
    main()
