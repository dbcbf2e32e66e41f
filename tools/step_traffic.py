"""profiles/rNN_step_dram_traffic.json from the ncu launch list of one bench.py step:
    python tools/step_traffic.py gpurun_out/launches.csv > profiles/r01_step_dram_traffic.json
(metrics pass: gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active; see tools/gpu_round.sh)."""
import collections, csv, json, re, sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(r[ix["ID"]], {"kernel": re.sub(r"^void (vsb::)?", "", r[ix["Kernel Name"]].split("(")[0])})
    d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["kernel"], {"kernel": d["kernel"], "launches": 0, "ms": 0.0, "dram_read_bytes": 0.0,
                                     "dram_write_bytes": 0.0, "_t": 0.0})
    ns = d.get("gpu__time_duration.sum", 0.0)
    a["launches"] += 1
    a["ms"] += ns / 1e6
    a["dram_read_bytes"] += d.get("dram__bytes_read.sum", 0.0)
    a["dram_write_bytes"] += d.get("dram__bytes_write.sum", 0.0)
    a["_t"] += ns * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
ks = []
for a in sorted(agg.values(), key=lambda a: -a["ms"]):
    a["tensor_active_pct"] = a.pop("_t") / (a["ms"] * 1e6) if a["ms"] else 0.0
    ks.append(a)
tot = sum(a["ms"] for a in ks)
json.dump({"what": "ncu per-launch metrics of one SF50 batch-64 step (bench.py --profile-range), grouped by kernel; "
                   "per-launch times are serialised and cold-cache, so shares (not absolutes) compare with bench.py",
           "total_ms": tot, "launches": sum(a["launches"] for a in ks),
           "dram_bytes": sum(a["dram_read_bytes"] + a["dram_write_bytes"] for a in ks),
           "conv_share_of_step": sum(a["ms"] for a in ks if a["kernel"].startswith("conv_")) / tot,
           "kernels": ks}, sys.stdout, indent=1)
