"""Print the judged counters of every kernel in an .ncu-rep (run where ncu is installed):
    python profiles/ncu_raw_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>_raw.txt
"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__maximum_warps_per_active_cycle_pct", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("==", r[idx["Kernel Name"]], "(launch id", r[idx["ID"]] + ")")
        for w in WANT:
            if w in idx:
                print(f"   {w:64s} {r[idx[w]]:>16s} {units[idx[w]]}")


if __name__ == "__main__":
    main(sys.argv[1])
