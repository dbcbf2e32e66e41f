import json, sys, collections
ops=json.load(open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/per_op.json'))
PEAK_TF, HBM = 1383.6, 6529.7
tot=sum(o['ms'] for o in ops)
ideal_tot=0
for o in ops:
    o['ideal']=max(o['gflop']/PEAK_TF, o.get('mbytes',0)/HBM/1e3*1e3/1e3*1e3) if False else max(o['gflop']/PEAK_TF, o.get('mbytes',0)/(HBM*1e3)*1e3)
    ideal_tot+=o['ideal']
print("total ms", round(tot,3), "ideal(per-op roofline) ms", round(ideal_tot,3))
for o in sorted(ops,key=lambda o:-(o['ms']-o['ideal']))[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    tf = o['gflop']/o['ms'] if o['ms']>0 else 0
    gbs = o.get('mbytes',0)/o['ms'] if o['ms']>0 else 0
    print(f"{o['op']:42s} {o['ms']:7.3f} ms ideal {o['ideal']:6.3f} gap {o['ms']-o['ideal']:6.3f} {o['gflop']:8.2f} GF {tf:7.1f} TF/s {o.get('mbytes',0):8.1f} MB {gbs:7.1f} GB/s")
g=collections.defaultdict(lambda:[0.0,0.0])
for o in ops:
    n=o['op']
    if 'stem' in n: k='stem'
    elif 'fuse' in n: k='fuse'
    elif 'pathway1' in n: k='fast '+n.split('.')[0]
    elif 'pathway0' in n: k='slow '+n.split('.')[0]
    else: k=n
    g[k][0]+=o['ms']; g[k][1]+=o['ideal']
for k,v in sorted(g.items(), key=lambda kv:-kv[1][0]): print(f"{k:30s} {v[0]:8.3f}  ideal {v[1]:7.3f}")
