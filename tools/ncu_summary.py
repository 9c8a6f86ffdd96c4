"""Summarise an .ncu-rep (one kernel launch, `ncu --set full`) as JSON: the raw-page metrics the judge reads
(duration, dram bytes, issue / tensor / warps-active percentages, occupancy limiters, stall ratios) and - when the
capture carries source (`--import-source on`) and the kernel is the tc kNN kernel - the per-region split of
tools/ncu_regions.py.

    python tools/ncu_summary.py gpurun_out/X.ncu-rep "description" > profiles/r02_X_ncu.json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__ops_path_tensor_src_bf16_dst_fp32.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")


def main():
    rep, desc = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    metrics = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            metrics[h] = {"unit": u, "value": v}
    out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "", "capture": desc, "metrics": metrics}
    if "--regions" in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        open("/tmp/_src.csv", "w").write(src)
        tool = "tools/ncu_regions_tc4.py" if "knn_tc4_kernel" in out["kernel"] else "tools/ncu_regions.py"
        reg = subprocess.run([sys.executable, tool, "/tmp/_src.csv"], capture_output=True, text=True).stdout
        out["regions_by_sass_landmarks"] = [l for l in reg.splitlines() if l.strip()]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
