import csv,sys,re
from collections import Counter
sass=sys.argv[1]; section=sys.argv[2]; ncucsv=sys.argv[3]
lines=open(sass).read().split('\n')
# locate section
start=None
for i,l in enumerate(lines):
    if l.startswith('\t.section\t.text.'+section) or ('.section' in l and '.text.' in l and section in l):
        start=i; break
ins=[]  # (offset, innermost(file,line), chain list, func label)
cur_chain=[]; pending=[]; func=section
i=start+1
while i<len(lines):
    l=lines[i]
    if l.startswith('\t.section'): break
    m=re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?',l)
    if m:
        pending.append((m.group(1).split('/')[-1],int(m.group(2))))
    else:
        m2=re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);',l)
        if m2:
            if pending: cur_chain=pending; pending=[]
            ins.append((int(m2.group(1),16),cur_chain,func,m2.group(2).strip()))
        else:
            m3=re.match(r'^(\$?[\w$.]+):\s*$',l)
            if m3 and not l.startswith('.L'): func=m3.group(1)
    i+=1
print('sass instructions',len(ins))
rows=list(csv.reader(open(ncucsv)))
hdr=rows[1]; ia=hdr.index("Address"); iex=hdr.index("Instructions Executed"); ismp=hdr.index("# Samples"); isrc=hdr.index("Source")
data=[r for r in rows[2:] if len(r)>iex]
print('ncu rows',len(data))
n=min(len(ins),len(data))
inner=Counter(); outer=Counter(); fn=Counter(); inner_ex=Counter(); outer_ex=Counter(); fn_ex=Counter()
tot=0; totex=0
for k in range(n):
    smp=int(data[k][ismp] or 0); ex=int(data[k][iex] or 0)
    off,chain,func,txt=ins[k]
    tot+=smp; totex+=ex
    a=chain[0] if chain else ('?',0); b=chain[-1] if chain else ('?',0)
    inner[a]+=smp; outer[b]+=smp; fn[func]+=smp
    inner_ex[a]+=ex; outer_ex[b]+=ex; fn_ex[func]+=ex
print('total samples',tot,'instr',totex)
print('--- by function label'); 
for k,v in fn.most_common(25): print('%6d %5.1f%%  ex %10d %5.1f%%  %s'%(v,100*v/tot,fn_ex[k],100*fn_ex[k]/totex,k[:150]))
print('--- by outermost line')
for k,v in outer.most_common(45): print('%6d %5.1f%%  ex %10d %5.1f%%  %s:%d'%(v,100*v/tot,outer_ex[k],100*outer_ex[k]/totex,k[0],k[1]))
print('--- by innermost line')
for k,v in inner.most_common(30): print('%6d %5.1f%%  ex %10d %5.1f%%  %s:%d'%(v,100*v/tot,inner_ex[k],100*inner_ex[k]/totex,k[0],k[1]))
