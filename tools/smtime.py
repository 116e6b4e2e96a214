"""Per-launch SM-time table from the CSV tools/gpu_smtime.sh writes: duration, mean SM-active time (cycles / clock), share."""
import csv
import re
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = OrderedDict()
for r in csv.DictReader(lines):
    k = int(r["ID"])
    d = rows.setdefault(k, {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("frcnn::", ""), "grid": r["Grid Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    d[r["Metric Name"]] = (v, r["Metric Unit"])
ids = [k for k, d in rows.items() if "conv_first" in d["name"]]
step = [rows[k] for k in rows if ids[-2] <= k < ids[-1]] if len(ids) >= 2 else list(rows.values())
step = [d for d in step if "at::" not in d["name"]]
tot_d = tot_a = 0.0
print("%-44s %9s %9s %9s %8s  %s" % ("kernel", "dur us", "active us", "tensor us", "act/dur", "grid"))
for d in step:
    dur, u = d["gpu__time_duration.sum"]
    dur = dur / 1e3 if u in ("ns", "nsecond") else dur
    el = d["sm__cycles_elapsed.max"][0]
    act = d["sm__cycles_active.avg"][0] / el * dur
    ten = d.get("sm__pipe_tensor_subpipe_hmma_cycles_active.avg", (0, ""))[0] / el * dur
    tot_d += dur
    tot_a += act
    print("%-44s %9.1f %9.1f %9.1f %8.2f  %s" % (d["name"][:44], dur, act, ten, act / dur, d["grid"]))
print("total duration %.1f us, total mean-SM-active %.1f us" % (tot_d, tot_a))
