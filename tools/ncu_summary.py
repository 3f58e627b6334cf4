#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [n_hot_lines] [--traffic KEY] -- key metrics + stall breakdown + hottest SASS lines
(reads with `ncu -i`).  With --traffic KEY the DRAM bytes of the first kernel of the report are also recorded in
profiles/ncu_traffic.json under KEY (e.g. "k_firpfbch2_analysis_fused@2^28"), which is where bench.py takes
`roofline.traffic` from."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
traffic_key = None
if "--traffic" in sys.argv:
    i = sys.argv.index("--traffic")
    traffic_key = sys.argv[i + 1]
    del sys.argv[i:i + 2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex.sum"]
def _bytes(val, unit):
    return float(val) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


if traffic_key and len(rows) > 2:
    import json
    import os
    r = rows[2]
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    try:
        db = json.load(open(path))
    except Exception:
        db = {}
    db[traffic_key] = {"dram_bytes_read": _bytes(r[ir], units[ir]), "dram_bytes_write": _bytes(r[iw], units[iw]),
                       "kernel": r[hdr.index("Kernel Name")][:120], "source": "ncu --set full, " + os.path.basename(rep)}
    json.dump(db, open(path, "w"), indent=1, sort_keys=True)
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")][:90])
    for w in want:
        if w in hdr:
            print("  %-70s %-14s %s" % (w, units[hdr.index(w)], r[hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
cols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = {h[i]: 0 for i in cols}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    try:
        n = int(r[4])
    except ValueError:
        continue
    for i in cols:
        try:
            tot[h[i]] += int(r[i])
        except ValueError:
            pass
    data.append((n, r[1][:80], {h[i][6:]: r[i] for i in cols if r[i] not in ("0", "")}))
S = sum(tot.values()) or 1
print("stall breakdown (% of samples):")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]:
    print("  %6.2f%% %s" % (100.0 * v / S, k))
print("hottest SASS lines:")
for n, s, st in sorted(data, key=lambda t: -t[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print("  %6d  %-80s %s" % (n, s, st))
