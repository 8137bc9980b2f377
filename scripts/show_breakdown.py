"""Print the per-layer timing table of one or more bench.py JSON lines side by side (ms per step / TFLOP/s)."""
import json
import sys

runs = []
for fn in sys.argv[1:]:
    line = [l for l in open(fn) if l.startswith("{")][-1]
    runs.append((fn, json.loads(line)))
keys = []
for _, d in runs:
    for k in d.get("kernel_breakdown", {}):
        if k not in keys:
            keys.append(k)
print("%-30s" % "layer" + "".join("%22s" % fn.split("/")[-1][-20:] for fn, _ in runs))
for k in keys:
    row = "%-30s" % k
    for _, d in runs:
        v = d["kernel_breakdown"].get(k)
        row += "%12.2f ms %6s" % (v["ms_per_step"], "" if not v or not v["tflops"] else "%.0f" % v["tflops"]) if v else " " * 22
    print(row)
print("%-30s" % "lines/s" + "".join("%22.1f" % d["value"] for _, d in runs))
print("%-30s" % "e2e lines/s" + "".join("%22.1f" % d["e2e"]["value"] for _, d in runs))
