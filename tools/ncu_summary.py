"""Summarise an .ncu-rep (ncu --set full) into a small markdown table for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.md "title" """
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (DMMA)"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA inst % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "stall samples: math pipe throttle"),
    ("smsp__pcsamp_warps_issue_stalled_wait", "stall samples: wait (fixed latency)"),
    ("smsp__pcsamp_warps_issue_stalled_not_selected", "stall samples: not selected"),
    ("smsp__pcsamp_warps_issue_stalled_barrier", "stall samples: barrier"),
    ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall samples: short scoreboard (LDS)"),
    ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall samples: long scoreboard"),
    ("smsp__pcsamp_warps_issue_stalled_selected", "samples: selected (issuing)"),
]
with open(out, "w") as f:
    f.write(f"# {title}\n\nSource: `{rep}` (ncu --set full --clock-control none), read with `ncu -i ... --page raw --csv`.\n\n")
    kernels = rows[2:]
    f.write("| metric | " + " | ".join(r[idx["Kernel Name"]].split("(")[0] for r in kernels) + " |\n")
    f.write("|---|" + "---|" * len(kernels) + "\n")
    for m, label in want:
        if m not in idx:
            continue
        f.write(f"| {label} ({units[idx[m]]}) | " + " | ".join(r[idx[m]] for r in kernels) + " |\n")
print(open(out).read())
