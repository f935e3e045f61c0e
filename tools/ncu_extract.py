#!/usr/bin/env python3
"""Extract the judged metrics from an .ncu-rep (ncu --set full) into a small CSV + markdown table.

    python tools/ncu_extract.py gpurun_out/prof_r01x.ncu-rep profiles/r01x
writes profiles/r01x_kernels.csv, profiles/r01x_kernels.md and updates profiles/dram_traffic.json.
"""
import csv
import json
import os
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def to_bytes(val, unit):
    v = float(val)
    u = unit.lower()
    for k, m in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("byte", 1.0)):
        if u.startswith(k):
            return v * m
    return v


def main(rep, out_prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = ["Kernel Name"] + [m for m in METRICS if m in idx]
    os.makedirs(os.path.dirname(out_prefix) or ".", exist_ok=True)
    with open(out_prefix + "_kernels.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([""] + [units[idx[m]] for m in cols[1:]])
        for r in rows[2:]:
            w.writerow([r[idx[c]] for c in cols])
    traffic = {}
    with open(out_prefix + "_kernels.md", "w") as f:
        f.write("| kernel | time | DRAM read | DRAM write | DRAM %% of ncu peak | achieved read | issue active | regs | grid x block |\n")
        f.write("|---|---|---|---|---|---|---|---|---|\n")
        for r in rows[2:]:
            g = lambda m: r[idx[m]] if m in idx else "?"
            u = lambda m: units[idx[m]] if m in idx else ""
            name = r[idx["Kernel Name"]].split("(")[0]
            f.write("| `%s` | %s %s | %s %s | %s %s | %s | %s %s | %s %% | %s | %s x %s |\n" % (
                name, g("gpu__time_duration.sum"), u("gpu__time_duration.sum"),
                g("dram__bytes_read.sum"), u("dram__bytes_read.sum"), g("dram__bytes_write.sum"), u("dram__bytes_write.sum"),
                g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                g("dram__bytes_read.sum.per_second"), u("dram__bytes_read.sum.per_second"),
                g("smsp__issue_active.avg.pct_of_peak_sustained_active"), g("launch__registers_per_thread"),
                g("launch__grid_size"), g("launch__block_size")))
            tot = to_bytes(g("dram__bytes_read.sum"), u("dram__bytes_read.sum")) + \
                to_bytes(g("dram__bytes_write.sum"), u("dram__bytes_write.sum"))
            traffic.setdefault(name, tot)
    tpath = os.path.join(os.path.dirname(out_prefix) or ".", "dram_traffic.json")
    cur = json.load(open(tpath)) if os.path.isfile(tpath) else {}
    cur.pop("pool_bwd_bytes_per_launch", None)          # (stale key of an early round-1 build: the backward is bwd_finish_kernel)
    for name, tot in traffic.items():
        if "pool_fwd_ldg_kernel<4" in name or "pool_fwd_tma_kernel<4>" in name:
            cur["pool_fwd_bytes_per_launch"] = tot
        if "pool_bwd_kernel" in name or "bwd_finish_kernel" in name:
            # both gradient maps in one launch; ncu sees only what reached DRAM inside the kernel's window (~60 MB of the
            # 268 MB are still dirty in L2 when it ends and drain during the next step)
            cur["bwd_bytes_per_launch_inside_kernel_window"] = tot
        if "retrify" in name:
            cur["retrify_weights_bytes_per_launch"] = tot
        if "pool_finish_cons" in name:
            cur["pool_finish_cons_bytes_per_launch"] = tot
        if "disc_fused" in name:
            cur["disc_fused_bytes_per_launch"] = tot
        if "mc_stats" in name:
            cur["mc_stats_bytes_per_launch"] = tot
    cur["source"] = os.path.basename(rep) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
    json.dump(cur, open(tpath, "w"), indent=1)
    print(open(out_prefix + "_kernels.md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
