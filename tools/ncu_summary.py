#!/usr/bin/env python
"""Condense an .ncu-rep (read here, no GPU needed) into the few numbers the roofline argument uses.

    python tools/ncu_summary.py gpurun_out/onesweep_v0.ncu-rep [more.ncu-rep ...] > profiles/rNN_x.md
Also writes the items as JSON (--json FILE; --variant NAME stamps the shape the capture was taken on)."""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smem atom wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "smem load wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "smem store wavefronts"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "ADU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math pipe"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "stall membar"),
]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        item = {"file": path, "kernel": d.get("Kernel Name", "?")}
        for k, label in KEYS:
            if k in d and d[k] != "":
                item[label] = f"{d[k]} {u[k]}".strip()
        try:
            rd = float(d["dram__bytes_read.sum"]) * TO_BYTES[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * TO_BYTES[u["dram__bytes_write.sum"]]
            item["dram_bytes_total"] = rd + wr
        except Exception:
            pass
        out.append(item)
    return out


def main():
    args = sys.argv[1:]
    json_out = None
    if "--json" in args:
        i = args.index("--json")
        json_out = args[i + 1]
        del args[i:i + 2]
    variant = None          # --variant NAME: the compiled shape the capture was taken on (bench.py checks it)
    if "--variant" in args:
        i = args.index("--variant")
        variant = args[i + 1]
        del args[i:i + 2]
    allitems = []
    for p in args:
        for item in summarize(p):
            if variant is not None:
                item["variant"] = variant
            allitems.append(item)
            print(f"### {item['kernel'][:120]}\n\nsource: `{item['file']}`\n")
            print("| metric | value |\n|---|---|")
            for k, v in item.items():
                if k not in ("file", "kernel"):
                    print(f"| {k} | {v} |")
            print()
    if json_out:
        with open(json_out, "w") as f:
            json.dump(allitems, f, indent=1)


if __name__ == "__main__":
    main()
