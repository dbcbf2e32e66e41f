import json, sys, collections
ops=json.load(open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/per_op.json'))
tot=sum(o['ms'] for o in ops)
print("total ms", round(tot,3))
for o in sorted(ops,key=lambda o:-o['ms'])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    tf = o['gflop']/o['ms'] if o['ms']>0 else 0
    print(f"{o['op']:45s} {o['ms']:8.3f} ms  {o['gflop']:9.2f} GF  {tf:8.1f} TF/s")
g=collections.defaultdict(float)
for o in ops:
    n=o['op']
    if 'stem' in n: k='stem'
    elif 'fuse' in n: k='fuse'
    elif 'pathway1' in n: k='fast '+n.split('.')[0]
    elif 'pathway0' in n: k='slow '+n.split('.')[0]
    else: k=n
    g[k]+=o['ms']
for k,v in sorted(g.items(), key=lambda kv:-kv[1]): print(f"{k:30s} {v:8.3f}")
