"""Stall reasons per SASS segment (segments = address ranges between BAR.SYNC / EXIT) of an ncu report: python scripts/ncu_stalls.py rep.ncu-rep"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == 'Address')
cols=[c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
idx={c:hdr.index(c) for c in cols}
iexe, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
sass=[]
for r in rows:
    if r and r[0].startswith('0x') and len(r)>isamp:
        sass.append((int(r[0],16), r[isrc], int(r[iexe] or 0), int(r[isamp] or 0), {c:int(r[idx[c]] or 0) for c in cols}))
sass.sort(key=lambda x:x[0]); base=sass[0][0]
# segments by BAR
seg=collections.Counter(); segn=0; sege=0; start=0; tot=sum(x[3] for x in sass)
k=0
for a,s,e,sm,st in sass:
    for c,v in st.items(): seg[c]+=v
    segn+=sm; sege+=e
    if 'BAR.SYNC' in s or 'EXIT' in s and segn>50:
        top=', '.join(f"{c[6:]} {100*v/max(segn,1):.0f}%" for c,v in seg.most_common(6))
        print(f"segment {k} (ends at +{a-base:x} {s.strip().split()[0]:8s}) samples {100*segn/tot:5.1f}%  : {top}")
        seg=collections.Counter(); segn=0; sege=0; k+=1
