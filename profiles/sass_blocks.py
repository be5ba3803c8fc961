import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; ia=hdr.index("Address"); isrc=hdr.index("Source"); iex=hdr.index("Instructions Executed"); ismp=hdr.index("# Samples")
blocks=[]; cur=None; tot=0; tots=0
def new(): return {'n':0,'ex':0,'smp':0,'start':None,'shfl':0,'imadw':0,'calls':0,'ldl':0,'stl':0,'stg':0,'ldg':0,'bar':0}
cur=new()
for r in rows[2:]:
    if len(r)<=iex: continue
    src=r[isrc]; ex=int(r[iex] or 0); smp=int(r[ismp] or 0)
    if cur['start'] is None: cur['start']=r[ia]
    cur['n']+=1; cur['ex']+=ex; cur['smp']+=smp; tot+=ex; tots+=smp
    for k,pat in (('shfl','SHFL'),('imadw','IMAD.WIDE'),('calls','CALL'),('bar','BAR')):
        if pat in src: cur[k]+=1
    for k,pat in (('stl','STL'),('ldl','LDL'),('stg','STG'),('ldg','LDG')):
        if pat in src: cur[k]+=ex
    if src.strip().startswith('RET') or 'EXIT' in src:
        blocks.append(cur); cur=new()
print("total instr executed", tot, "samples", tots, "blocks", len(blocks))
blocks.sort(key=lambda b:-b['smp'])
for b in blocks[:16]:
    print("%s n=%5d ex=%10d (%4.1f%%) smp=%6d (%4.1f%%) shfl=%d imadw=%d calls=%d ldl=%d stl=%d ldg=%d stg=%d"%(b['start'][-6:],b['n'],b['ex'],100*b['ex']/tot,b['smp'],100*b['smp']/tots,b['shfl'],b['imadw'],b['calls'],b['ldl'],b['stl'],b['ldg'],b['stg']))
