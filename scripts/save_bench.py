"""Split bench.py's JSON line into profiles/bench_<round>_<workload>.json (the headline with its verify block, and one
file per record of its `configs` list):   python bench.py | tail -1 | python scripts/save_bench.py r02"""
import json, os, sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles")
configs = d.pop("configs", None) or []
suffix = "" if d.get("n_gpus", 1) == 1 else "_n%d" % d["n_gpus"]
with open(os.path.join(root, "bench_%s_%s%s.json" % (rnd, d["config"]["workload"], suffix)), "w") as f:
    json.dump(d, f, indent=1)
    f.write("\n")
for c in configs:
    with open(os.path.join(root, "bench_%s_%s.json" % (rnd, c["config"]["workload"])), "w") as f:
        json.dump(c, f, indent=1)
        f.write("\n")
print("saved", d["config"]["workload"], [c["config"]["workload"] for c in configs])
