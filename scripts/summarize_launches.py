"""Summarise an `ncu --metrics gpu__time_duration.sum` launch list of `bench.py`: the list is cut into training steps at every
to_ndhwc_kernel launch (the first kernel of a forward pass) and one step (default: the third from last, a timed graph replay)
is tabulated per kernel.  ncu times every launch alone and cold, so compare SHARES of the step, not absolute times."""
import csv, sys, collections
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else -3
rows = list(csv.reader(open(path, errors="replace")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}
L = [(r[kn].split("(")[0][:62], float(r[mv].replace(",", "")) * scale.get(r[mu].replace("second", ""), 1e-6))
     for r in rows[hdr + 1:] if len(r) > mv]
starts = [i for i, (n, _) in enumerate(L) if n.startswith("to_ndhwc")]
segs = [(starts[i], starts[i + 1] if i + 1 < len(starts) else len(L)) for i in range(len(starts))]
print("%d launches in the run, %d steps: launches per step %s" % (len(L), len(segs), [b - a for a, b in segs]))
a, b = segs[which]
tot = collections.defaultdict(lambda: [0, 0.0])
for n, t in L[a:b]:
    tot[n][0] += 1; tot[n][1] += t
total = sum(v[1] for v in tot.values())
print("step %d: %d launches, %.3f ms summed (each launch timed alone)" % (which, b - a, total))
fam = collections.defaultdict(float)
for n, (c, t) in tot.items():
    f = ("conv fprop/dgrad (zs/igemm)" if ("zs_kernel" in n or "igemm_kernel" in n) else
         "weight gradients (wg2/wgp/wgrad+finalize)" if ("wg2" in n or "wgp" in n or "wgrad" in n) else
         "GroupNorm" if n.startswith("void gn_") or n.startswith("gn_") else
         "attention gate" if "gate_" in n or "channel_sum" in n or "add_channel" in n else
         "loss" if "loss_" in n else
         "heads / pool / layout" if any(s in n for s in ("head", "lerp", "trilinear", "pool", "ndhwc", "ncdhw")) else
         "optimizer + weight re-pack" if ("multi_tensor" in n or "pack_" in n or n.strip() == "void at::native::" or "at::<unnamed>" in n) else
         "other (torch fills, memsets, reductions)")
    fam[f] += t
for f, t in sorted(fam.items(), key=lambda kv: -kv[1]):
    print("  %-46s %8.3f ms %5.1f%%" % (f, t, 100 * t / total))
print()
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 60]:
    print("%-64s %4d %8.3f ms %5.1f%%" % (n, c, t, 100 * t / total))
