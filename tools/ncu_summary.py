#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): one line per profiled launch with the metrics the roofline uses.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [> profiles/x.summary.txt]"""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("sm__cycles_elapsed.avg", "cycles"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__t_bytes.sum", "l2_bytes"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "xu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp_insts")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        parts = [name.split("(")[0][-40:]]
        for k, short in KEYS:
            cands = [i for i, h in enumerate(hdr) if h == k or h.endswith("." + k)]
            if cands:
                i = cands[0]
                parts.append("%s=%s%s" % (short, r[i], (" " + units[i]) if units[i] else ""))
        print("  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
