"""Regenerate profiles/ncu_traffic.json from `ncu --set full` captures (.ncu-rep files under gpurun_out/ or profiles/).

    python scripts/ncu_traffic.py <kernel label>=<report.ncu-rep>:<algorithmic bytes>:<description> ... [--roofline <label>]

For every report: dram__bytes_read.sum + dram__bytes_write.sum of the FIRST captured launch (per launch, like roofline.achieved),
its duration, and the algorithmic bytes given on the command line.  bench.py copies the entry named by --roofline into the
`roofline.traffic` field of its JSON line, so the number in the bench line is the committed measurement of THIS kernel set."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    def get(name):
        i = h.index(name)
        val = float(v[i].replace(",", ""))
        unit = u[i]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "%": 1.0}.get(unit, 1.0)
        return val * scale
    name = v[h.index("Kernel Name")]
    d = {"kernel_name": name, "dram_bytes": get("dram__bytes_read.sum") + get("dram__bytes_write.sum"),
         "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
         "duration_s": get("gpu__time_duration.sum"), "registers_per_thread": get("launch__registers_per_thread")}
    for opt in ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tensor.sum"):
        if opt in h:
            d[opt] = get(opt)
    return d


def main():
    args = sys.argv[1:]
    roof = None
    if "--roofline" in args:
        i = args.index("--roofline")
        roof = args[i + 1]
        del args[i:i + 2]
    kernels = {}
    for a in args:
        label, rest = a.split("=", 1)
        rep, alg, desc = rest.split(":", 2)
        m = metrics(rep)
        m["algorithmic_bytes"] = float(alg)
        m["desc"] = desc
        m["report"] = os.path.basename(rep)
        kernels[label] = m
        print("%s: %.1f MB DRAM (%.1f read + %.1f written) vs %.1f MB algorithmic, %.1f us" % (
            label, m["dram_bytes"] / 1e6, m["dram_bytes_read"] / 1e6, m["dram_bytes_write"] / 1e6, float(alg) / 1e6, m["duration_s"] * 1e6))
    out = {"source": "ncu --set full --clock-control none, scripts/ncu_traffic.py", "roofline_kernel": roof or next(iter(kernels)),
           "kernels": kernels}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
