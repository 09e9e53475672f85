"""Summarise an ncu CSV with gpu__time_duration.sum + dram bytes per launch: per-kernel time, DRAM GB and achieved GB/s (last step)."""
import csv, sys, collections
path = sys.argv[1]; nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = list(csv.reader(open(path, errors="replace")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
idc, kn, mn, mv, mu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
L = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    d = L.setdefault(r[idc], {"k": r[kn]})
    u = r[mu]
    if r[mn].startswith("gpu__time"):
        d["ms"] = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else (v if u in ("ms", "msecond") else v * 1e3))
    else:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        d[r[mn]] = v * scale
L = list(L.values())
n = len(L) // nsteps
last = L[-n:]
tot = collections.defaultdict(lambda: [0, 0.0, 0.0])
for d in last:
    k = d["k"].split("(")[0][:60]
    t = tot[k]; t[0] += 1; t[1] += d.get("ms", 0); t[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
total = sum(v[1] for v in tot.values())
print("last step: %d launches, %.3f ms summed" % (n, total))
for k, (c, ms, by) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print("%-62s %4d %8.3f ms %5.1f%%  %8.1f MB  %7.0f GB/s" % (k, c, ms, 100 * ms / total, by / 1e6, by / ms / 1e6 if ms else 0))
