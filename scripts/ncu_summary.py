"""Summarise an `ncu -i X.ncu-rep --page raw --csv` export: one block per captured launch with the
metrics that decide the roofline position (duration, DRAM/L2/L1 traffic and hit rates, pipe utilisation,
occupancy, top stall reasons).

usage: python scripts/ncu_summary.py gpurun_out/X_raw.csv > profiles/rNN_ncu_summary.txt
"""
import csv
import re
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %peak"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "LSU wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64 inst"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_fma.sum", "fma inst"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_lsu.sum", "lsu inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp inst"),
]


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    col = {h: i for i, h in enumerate(header)}
    stall_cols = [h for h in header if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in header if "warp_issue_stalled" in h and h.endswith("pct")]
    for row in rd:
        name = re.sub(r"\(.*", "", row[col["Kernel Name"]])
        print(f"== {name}  id={row[col['ID']]}")
        for k, label in KEYS:
            if k in col:
                print(f"   {label:16s} {row[col[k]]:>16s} {units[col[k]]}")
        st = []
        for h in stall_cols:
            try:
                st.append((float(row[col[h]].replace(",", "")), h))
            except ValueError:
                pass
        st.sort(reverse=True)
        for v, h in st[:5]:
            short = re.sub(r"smsp__average_warps_issue_stalled_|_per_issue_active.ratio|smsp__pcsamp_|warp_issue_stalled_", "", h)
            print(f"   stall {short:28s} {v:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
