"""Executed instructions / stall samples per SASS segment (address ranges between BAR.SYNC / EXIT): python scripts/ncu_segments.py rep.ncu-rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == 'Address')
iexe, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
sass = [(int(r[0],16), r[isrc], int(r[iexe] or 0), int(r[isamp] or 0)) for r in rows if r and r[0].startswith('0x') and len(r) > isamp]
sass.sort()
te = sum(x[2] for x in sass); ts = sum(x[3] for x in sass)
seg_e = seg_s = seg_n = 0; start = sass[0][0]
base = sass[0][0]
for a, s, e, sm in sass:
    seg_e += e; seg_s += sm; seg_n += 1
    if 'BAR.SYNC' in s or 'EXIT' in s and seg_n > 20:
        print(f"{start-base:6x}-{a-base:6x} n={seg_n:5d} exe={seg_e:10d} ({100*seg_e/te:5.1f}%) samp={seg_s:6d} ({100*seg_s/ts:5.1f}%)  [{s.strip()[:30]}] barexe={e}")
        seg_e = seg_s = seg_n = 0; start = a
print(f"rest n={seg_n} exe={seg_e} samp={seg_s}; total {te} {ts}")
