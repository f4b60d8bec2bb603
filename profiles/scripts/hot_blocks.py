import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
sidx=[hdr.index(h) for h in stalls]
sec=[]
for r in rows[2:]:
    if len(r)<10 or r[0]=="Kernel Name": break
    sec.append(r)
tot=sum(int(r[iE]) for r in sec); totS=sum(int(r[iSm]) for r in sec)
print("total inst", tot, "samples", totS)
blocks=[]; cur=None
for i,r in enumerate(sec):
    e=int(r[iE])
    if cur is None or e!=cur[0]: cur=[e,i,i]; blocks.append(cur)
    else: cur[2]=i
big=sorted(blocks,key=lambda b:-b[0]*(b[2]-b[1]+1))[:int(sys.argv[2]) if len(sys.argv)>2 else 14]
for b in sorted(big,key=lambda b:b[1]):
    n=b[2]-b[1]+1; ops={}; smp=0; st={}
    for r in sec[b[1]:b[2]+1]:
        t=r[iS].split(); op=t[1] if t[0].startswith('@') else t[0]
        op=op.split('.')[0]; ops[op]=ops.get(op,0)+1; smp+=int(r[iSm])
        for h,ix in zip(stalls,sidx):
            v=int(r[ix] or 0)
            if v: st[h]=st.get(h,0)+v
    print(f"lines {b[1]}-{b[2]} n={n} exec={b[0]} inst%={100*b[0]*n/tot:.1f} samp%={100*smp/totS:.1f}", sorted(ops.items(),key=lambda x:-x[1])[:8], sorted(st.items(),key=lambda x:-x[1])[:5])
