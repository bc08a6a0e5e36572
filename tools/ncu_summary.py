"""Markdown table of the counters the roofline discussion uses, from an `ncu --set full` report (reads it with `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep [more reports ...] > profiles/<round>_ncu_<what>_summary.md
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread",
    "launch__grid_size",
    "launch__block_size",
    "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def table(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    names = []
    for r in body:
        n = r[col["Kernel Name"]]
        n = n.split("(")[0].replace("hmmb200::", "").replace("void ", "")
        names.append(n)
    out = [f"Report `{path}` ({len(body)} launch(es)); values per launch.", "",
           "| metric | unit | " + " | ".join(names) + " |", "|---|---|" + "---|" * len(names)]
    for m in METRICS:
        if m not in col:
            continue
        out.append(f"| {m} | {units[col[m]]} | " + " | ".join(r[col[m]] for r in body) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(table(p))
        print()
