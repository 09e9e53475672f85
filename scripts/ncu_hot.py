"""Top stall-sample SASS lines of an .ncu-rep captured with --import-source on: python scripts/ncu_hot.py rep [N]"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[1]
si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stalls = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
rows = []
for idx, x in enumerate(r[2:]):
    try: rows.append((int(x[si]), idx, x))
    except Exception: pass
tot = sum(a for a, _, _ in rows)
print("total samples", tot, " instructions", len(rows))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for a, idx, x in sorted(rows, reverse=True)[:N]:
    st = sorted(((int(x[i]), h[i][6:]) for i in stalls if x[i].isdigit() and int(x[i]) > 0), reverse=True)[:2]
    print("%6d %5.1f%% #%-5d exec %-8s %-70s %s" % (a, 100.0 * a / tot, idx, x[ie], x[so].strip()[:70], st))
