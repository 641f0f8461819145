"""DRAM traffic of a workload from an `ncu --set full` capture -> profiles/ncu_traffic.json (read by bench.py).

    python tools/ncu_traffic.py <rep.ncu-rep> <workload> <units> [--launches-per-step N] [--skip K]

Sums dram__bytes_read.sum + dram__bytes_write.sum over the captured launches (after skipping K), divides by the number
of steps the capture covers (launches / launches-per-step) and records it as the per-step traffic of `workload` at
`units` work items (points / rows / samples), with the per-kernel breakdown."""
import argparse, csv, json, os, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_traffic.json")
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"), ap.add_argument("workload"), ap.add_argument("units", type=float)
    ap.add_argument("--launches-per-step", type=int, default=0, help="0 = the whole capture is one step")
    ap.add_argument("--skip", type=int, default=0)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2 + a.skip:]
    ni, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    per_kernel, total = {}, 0.0
    for r in data:
        b = float(r[ri]) * SCALE[units[ri]] + float(r[wi]) * SCALE[units[wi]]
        k = r[ni].split("(")[0].replace("void ", "")
        e = per_kernel.setdefault(k, {"launches": 0, "dram_bytes": 0.0})
        e["launches"] += 1
        e["dram_bytes"] += b
        total += b
    n_steps = len(data) / a.launches_per_step if a.launches_per_step else 1.0
    doc = {"workloads": {}}
    if os.path.exists(OUT):
        doc = json.load(open(OUT))
    doc["workloads"][a.workload] = {
        "units": a.units, "dram_bytes": total / n_steps, "steps_captured": n_steps,
        "source": os.path.basename(a.rep) + " (ncu --set full --clock-control none, B200)",
        "kernels": {k: {"launches_per_step": v["launches"] / n_steps, "dram_bytes_per_step": v["dram_bytes"] / n_steps}
                    for k, v in per_kernel.items()}}
    json.dump(doc, open(OUT, "w"), indent=1)
    print(json.dumps(doc["workloads"][a.workload], indent=1))


if __name__ == "__main__":
    main()
