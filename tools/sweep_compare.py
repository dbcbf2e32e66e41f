"""Compare gpurun_out/sweep_*.json against sweep_base.json: per op, the best variant and its gain."""
import glob, json, sys
base = json.load(open("gpurun_out/sweep_base.json"))
bops = dict(base["ops"])
vs = {}
for p in sorted(glob.glob("gpurun_out/sweep_*.json")):
    d = json.load(open(p))
    if d["name"] != "base":
        vs[d["name"]] = dict(d["ops"])
    print(f"{d['name']:14s} step {d['step_ms']:.3f} ms  sum {sum(dict(d['ops']).values()):.3f}")
tot = 0.0
for op, b in bops.items():
    best, bn = b, "base"
    for n, o in vs.items():
        if op in o and o[op] < best:
            best, bn = o[op], n
    if b - best > 0.004:
        print(f"{op:40s} base {b:.3f} best {best:.3f} ({bn}) gain {b-best:.3f}  | " + " ".join(f"{n}={o.get(op, 0):.3f}" for n, o in vs.items()))
        tot += b - best
print("total gain if picking best per op: %.3f ms" % tot)
