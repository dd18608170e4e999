"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_name.txt [--traffic-key kernel_name]
"""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]
UNIT_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tkey = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of {Path(rep).name}", ""]
    traffic = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines.append(f"## {d.get('Kernel Name', '?')[:110]}")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:85s} {d[k]:>16s} {u[k]}")
        try:
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * UNIT_BYTES[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * UNIT_BYTES[u["dram__bytes_write.sum"]]
            lines.append(f"{'dram traffic (read + write), bytes per launch':85s} {rd + wr:16.0f}")
            traffic.append(rd + wr)
        except (KeyError, ValueError):
            pass
        lines.append("")
    Path(out).write_text("\n".join(lines))
    if tkey and traffic:
        tj = Path(out).parent / "traffic.json"
        cur = json.loads(tj.read_text()) if tj.exists() else {}
        # bench.py reads the keys "msm_accum_points_kernel" (one launch) and "sc_round_kernel" (WHOLE sumcheck: scale the captured
        # launches to 8.59 GB by hand, see the note in traffic.json); every other capture is stored under a "capture:" key
        if tkey in ("msm_accum_points_kernel", "sc_round_kernel"):
            tkey = "capture:" + tkey
        cur[tkey] = sum(traffic) if "sc_" in tkey else traffic[0]
        cur[tkey + "_note"] = f"from {Path(rep).name}: " + ("sum over captured launches" if "sc_" in tkey else "one launch")
        tj.write_text(json.dumps(cur, indent=1))
    print("\n".join(lines))


if __name__ == "__main__":
    main()
