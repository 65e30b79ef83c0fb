"""Summarise an .ncu-rep (read here, no GPU needed): one line block per captured launch with the metrics the
roofline discussion uses.  Usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.txt"""
import csv, io, subprocess, sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "sm cycles"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC (elapsed)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu (MUFU) pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "membar", "math_pipe_throttle",
               "mio_throttle", "branch_resolving", "no_instruction", "not_selected", "lg_throttle", "sleeping"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in body:
        print("=" * 100)
        print(r[col["Kernel Name"]][:140])
        for key, label in KEYS:
            if key in col:
                print(f"  {label:28s} {r[col[key]]:>18s} {units[col[key]]}")
        st = []
        for n in STALL_NAMES:
            k = STALLS % n
            if k in col:
                st.append((float(r[col[k]] or 0), n))
        st.sort(reverse=True)
        print("  stall cycles per issued instr: " + ", ".join(f"{n}={v:.2f}" for v, n in st[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
