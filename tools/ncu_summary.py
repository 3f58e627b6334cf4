#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep -- key metrics + stall breakdown + hottest SASS lines (reads with `ncu -i`)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
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
