#!/usr/bin/env python
"""Summarise an `ncu --set full` capture (read on the CPU box with `ncu -i ... --page raw --csv`) into the small CSV
kept under profiles/, and record the measured DRAM traffic per launch of a kernel in profiles/traffic.json, which
bench.py reports as roofline.traffic.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x_ncu_full_summary.csv [--traffic-key cfg2_n256]
"""
import argparse
import csv
import io
import json
import os
import subprocess

KEEP = (
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "sass__inst_executed_local_loads",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
)
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--traffic-key", default=None)
    ap.add_argument("--kernel", default=None, help="substring of the kernel name (default: every captured launch)")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    name_col = ix.get("Kernel Name")
    with open(args.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{k} [{units[ix[k]]}]" for k in KEEP if k in ix])
        for r in data:
            if args.kernel and args.kernel not in r[name_col]:
                continue
            w.writerow([r[name_col][:90]] + [r[ix[k]] for k in KEEP if k in ix])
    if args.traffic_key:
        sel = [r for r in data if not args.kernel or args.kernel in r[name_col]]
        r = sel[-1]
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[ix[k]]) * UNIT_SCALE[units[ix[k]]]
        path = os.path.join(os.path.dirname(os.path.abspath(args.out)), "traffic.json")
        d = json.load(open(path)) if os.path.exists(path) else {}
        d[args.traffic_key] = {"dram_bytes_per_launch": tot, "kernel": r[name_col][:90],
                               "source": os.path.relpath(args.out, os.path.dirname(os.path.dirname(os.path.abspath(args.out))))
                                         + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch)"}
        json.dump(d, open(path, "w"), indent=1, sort_keys=True)
        print(f"{args.traffic_key}: {tot / 1e9:.3f} GB per launch -> {path}")


if __name__ == "__main__":
    main()
