#!/usr/bin/env python3
"""Group the SASS dump of scripts/ncu_sass_dump.py into runs of instructions with the same execution count (loops,
branches): instructions, share of all warp-instructions and of all stall samples per run.  Usage:
    python scripts/ncu_sass_dump.py x.ncu-rep > dump.txt; python scripts/ncu_segments.py dump.txt [min share]"""
import re,sys
rows=[]
for l in open(sys.argv[1]):
    m=re.match(r'(\S+)\s+(\d+)\s+(\d+)\s+(\d+)\s+(.*)',l)
    if m: rows.append((m.group(1),int(m.group(2)),int(m.group(3)),int(m.group(4)),m.group(5)))
tot=sum(r[1] for r in rows); ts=sum(r[2] for r in rows)
print(len(rows),tot,ts)
prev=None; seg_start=0; acc=0; accs=0
segs=[]
for i,r in enumerate(rows):
    e=r[1]
    if prev is not None and (e>prev*1.15 or e<prev/1.15):
        segs.append((seg_start,i-1,acc,accs)); seg_start=i; acc=0; accs=0
    acc+=e; accs+=r[2]; prev=e if e>0 else prev
segs.append((seg_start,len(rows)-1,acc,accs))
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.003
for s in segs:
    if s[2]>tot*thr: print(s[0],s[1],rows[s[0]][0],'n=%d'%(s[1]-s[0]+1),'exec/inst=%d'%(s[2]/(s[1]-s[0]+1)),'Minstr=%.1f'%(s[2]/1e6),'share=%.1f%%'%(100*s[2]/tot),'samples=%.1f%%'%(100*s[3]/ts))
