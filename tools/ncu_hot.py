"""Top stalled SASS instructions of one launch in an ncu report (needs -lineinfo + --import-source on).
    python tools/ncu_hot.py gpurun_out/prof.ncu-rep LAUNCH_INDEX [TOP_N]"""
import csv, io, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:120])
hdr = rows[1]
data = []
for r in rows[2:]:
    if not r or not r[0].startswith('0x'):
        break
    data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print("stall mix:", ", ".join(f"{h[6:]}={100*v/tot:.1f}%" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:topn]
for i in order:
    r = data[i]
    s = int(r[ix['# Samples']])
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print(f"{i:5d} {s:6d} {100*s/tot:5.1f}% exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:64]:64s} {st}")
