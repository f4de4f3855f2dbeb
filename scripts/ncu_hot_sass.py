"""The N SASS instructions with the most stall samples of an ncu report, with their dominant stall reason: python scripts/ncu_hot_sass.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == 'Address')
cols = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
idx = {c: hdr.index(c) for c in cols}
iexe, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
sass = []
for r in rows:
    if r and r[0].startswith('0x') and len(r) > isamp:
        st = {c: int(r[idx[c]] or 0) for c in cols}
        sass.append((int(r[0], 16), r[isrc].strip(), int(r[iexe] or 0), int(r[isamp] or 0), max(st, key=st.get)[6:]))
base = min(x[0] for x in sass); tot = sum(x[3] for x in sass)
print(f"# the {N} SASS instructions with the most stall samples (of {tot})")
for a, s, e, sm, why in sorted(sass, key=lambda x: -x[3])[:N]:
    print(f"  {a - base:6x} exe={e:8d} samp={sm:4d} ({100 * sm / tot:4.1f}%) {why:12s} {s[:60]}")
