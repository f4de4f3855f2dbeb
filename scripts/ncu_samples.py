"""Stall-sample (time) distribution per source line of an ncu report: python scripts/ncu_samples.py rep [N]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == 'Line No')
isamp, iexe = hdr.index('# Samples'), hdr.index('Instructions Executed')
cur = None; samp = collections.Counter(); exe = collections.Counter(); text = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) <= isamp or not r[0].isdigit(): continue
    try:
        k = (cur, int(r[0])); samp[k] += int(r[isamp]); exe[k] += int(r[iexe]); text[k] = r[1].strip()[:95]
    except ValueError: pass
tot, te = sum(samp.values()), sum(exe.values())
print("total samples", tot, "instructions", te)
for k, n in samp.most_common(N):
    print(f"{100*n/tot:5.2f}% time {100*exe[k]/te:5.2f}% inst  {k[0]}:{k[1]}  {text[k]}")
b = collections.Counter(); be = collections.Counter()
for k, n in samp.items(): b[(k[0], k[1] // 20 * 20)] += n; be[(k[0], k[1] // 20 * 20)] += exe[k]
print("by 20-line bucket (time%, inst%):")
for k in sorted(b):
    if b[k] * 100 > tot: print(f"  {k[0]}:{k[1]:4d}  {100*b[k]/tot:5.1f}%  {100*be[k]/te:5.1f}%")
