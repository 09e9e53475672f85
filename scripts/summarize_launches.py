"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals for the LAST step.
usage: python scripts/summarize_launches.py launches.csv [n_steps_in_file] [--list PATTERN]"""
import csv, sys, collections
path = sys.argv[1]
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 2
pat = sys.argv[sys.argv.index("--list") + 1] if "--list" in sys.argv else None
rows = list(csv.reader(open(path, errors="replace")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
L = []
for r in rows[hdr + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    u = r[mu]
    ms = v / 1e6 if u in ("ns", "nsecond") else (v / 1e3 if u in ("us", "usecond") else v)
    L.append((r[kn], ms))
n = len(L) // nsteps
last = L[-n:]
tot = collections.defaultdict(lambda: [0, 0.0])
for k, ms in last:
    k = k.split("(")[0][:70]
    tot[k][0] += 1; tot[k][1] += ms
total = sum(v[1] for v in tot.values())
print("last step: %d launches, %.3f ms summed (serialised, cold-cache)" % (n, total))
for k, (c, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-72s %5d %9.3f ms %5.1f%%" % (k, c, ms, 100 * ms / total))
if pat:
    print("--- launches matching", pat)
    for i, (k, ms) in enumerate(last):
        if pat in k: print(i, k[:60], "%.3f" % ms)
