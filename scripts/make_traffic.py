#!/usr/bin/env python
"""profiles/r02_traffic.json from the --set full captures of one visit (gpurun_out/prof_c3.ncu-rep, prof_c2.ncu-rep):
DRAM bytes read + written per launch of the kernels bench.py's `roofline` covers. usage: make_traffic.py <visit tag>"""
import csv, json, subprocess, sys


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    res = []
    for r in rows[2:]:
        g = lambda k: float(r[h.index(k)].replace(",", ""))
        unit = lambda k: rows[1][h.index(k)]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = g("dram__bytes_read.sum") * scale[unit("dram__bytes_read.sum")]
        wr = g("dram__bytes_write.sum") * scale[unit("dram__bytes_write.sum")]
        res.append((r[h.index("Kernel Name")], rd, wr, g("gpu__time_duration.sum")))
    return res


tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = {}
for wl, rep, want in (("C3", "gpurun_out/prof_c3.ncu-rep", ("k_match_pt2pt", "k_plane_fit", "k_compact_pt2pl")), ("C2", "gpurun_out/prof_c2.ncu-rep", ("k_iterate_nn1_horn",))):
    ks = kernels(rep)
    rd = wr = 0.0
    parts = {}
    for w in want:
        m = [k for k in ks if w in k[0]]
        if m:
            rd += m[0][1]
            wr += m[0][2]
            parts[w] = {"dram_bytes_read": m[0][1], "dram_bytes_write": m[0][2], "ncu_time_us": m[0][3]}
    out[wl] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "kernels": parts,
               "source": f"ncu --set full --clock-control none, {tag}, one launch of each kernel of the timed step function (profiles/{tag.split()[0]}_ncu_{wl.lower()}_kernels.txt)"}
json.dump(out, open("profiles/r02_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
