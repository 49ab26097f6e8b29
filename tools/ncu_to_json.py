"""Per-launch counters of the classify kernel from an `ncu --set full` report into
profiles/ncu_kernels.json, keyed by workload (run here, no GPU needed):

    python tools/ncu_to_json.py gpurun_out/r2_c3_mini.ncu-rep c3 <windows> [profiles/<summary>.txt]

bench.py reads the file for `roofline.traffic` (dram__bytes_read + dram__bytes_write per launch)
and for the integer-ALU roofline (ALU-pipe warp instructions x 32 / windows)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "": 1.0, "%": 1.0,
        "inst": 1.0, "sector": 1.0, "register/thread": 1.0, "cycle": 1.0, "warp": 1.0, "request": 1.0}


def main(rep, workload, windows, summary):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, row = rows[0], rows[1], rows[-1]

    def get(name):
        i = hdr.index(name)
        return float(row[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    windows = int(windows)
    # ncu reports the pipe's instruction count only as a share of its peak issue rate, which is
    # 2 warp instructions per SM and cycle (four sub-partitions, 16 lanes each)
    sm_cycles = get("sm__cycles_active.sum")
    alu = get("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100.0 * 2.0 * sm_cycles
    fma = get("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") / 100.0 * 2.0 * sm_cycles
    entry = {
        "kernel": row[hdr.index("Kernel Name")],
        "windows": windows,
        "duration_ms": get("gpu__time_duration.sum"),
        "dram_bytes_per_launch": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
        "dram_pct_of_peak": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "alu_warp_insts_per_launch": alu,
        "alu_thread_insts_per_window": alu * 32.0 / windows,
        "fma_warp_insts_per_launch": fma,
        "warp_insts_per_launch": get("smsp__inst_executed.sum"),
        "alu_pipe_pct": get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "l1_sector_hit_pct": get("l1tex__t_sector_hit_rate.pct"),
        "l1_sector_misses": get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum"),
        "l2_sector_hit_pct": get("lts__t_sector_hit_rate.pct"),
        "registers": get("launch__registers_per_thread"),
        "source": f"{summary or os.path.basename(rep)} (ncu --set full --clock-control none, {workload} workload)",
    }
    path = os.path.join(ROOT, "profiles", "ncu_kernels.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[workload] = entry
    with open(path, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
