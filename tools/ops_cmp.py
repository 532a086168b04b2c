import json,sys
tag=sys.argv[1]; n=int(sys.argv[2]); pat=sys.argv[3].split(',')
rows={}
for i in range(n):
    d=json.load(open(f"gpurun_out/{tag}_ops_{i}.json"))
    ops=d if isinstance(d,list) else d.get("ops",d)
    for o in ops:
        nm=o['name']
        if any(p in nm for p in pat): rows.setdefault(nm,[None]*n)[i]=round(o['ms']*1e3,1)
    rows.setdefault('TOTAL',[None]*n)[i]=round(sum(o['ms'] for o in ops)*1e3)
for k,v in rows.items(): print(f"{k:28s}",v)
