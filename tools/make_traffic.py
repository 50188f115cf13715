#!/usr/bin/env python
"""DRAM bytes per site per kernel from an `ncu --set full` capture of tools/kbench.py -> profiles/rNN_traffic.json
(read by bench.py for roofline.traffic).  usage: python tools/make_traffic.py gpurun_out/x.ncu-rep SITES_PER_LAUNCH out.json"""
import csv
import json
import subprocess
import sys

KERNELS = {"prep_tiles48": "prep_tiles", "lstm_seq_x2": "lstm_seq_x2", "lstm_seq": "lstm_seq1", "l3l4_fused": "l3l4_fused",
           "heads_tc": "heads_tc", "decide_sites": "decide_sites", "create_tensors": "create_tensors"}
path, sites, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
col = lambda k: [i for i, h in enumerate(hdr) if h == k][0]
rd, wr, nm = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), hdr.index("Kernel Name")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
acc = {}
for r in rows[2:]:
    base = r[nm].replace("void ", "").split("(")[0].split("<")[0].split("::")[-1].strip()
    key = KERNELS.get(base)
    if key is None:
        continue
    b = float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]]
    acc.setdefault(key, []).append(b)
res = {"source": "%s (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch on a %d-site chunk, mean over the captured launches)" % (path, sites),
       "sites_per_launch": sites, "dram_bytes_per_site": {k: sum(v) / len(v) / sites for k, v in acc.items()},
       "launches_captured": {k: len(v) for k, v in acc.items()}}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
