"""Per-phase executed-instruction and stall-sample shares of the fused kernel from an ncu report.
python scripts/ncu_phases.py rep.ncu-rep   (phase boundaries are read from the '// -- Pn' markers of mdn_fused.cuh;
lines of mdn_common.cuh / headers are attributed by SASS address ORDER: an instruction belongs to the phase of the nearest
preceding mdn_fused.cuh-attributed instruction)"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
src = open('mdn_sfm_b200/csrc/mdn_fused.cuh').read().split('\n')
marks = []
for i, l in enumerate(src, 1):
    m = re.search(r'// -+ (P\d|tail|per \(target|P0)', l)
    if m: marks.append((i, m.group(1)))
KSTART = next(i for i, l in enumerate(src, 1) if 'fused_tile_kernel(const __grid_constant__' in l)
def phase_of(line):
    name = 'prologue'
    for i, n in marks:
        if line >= i: name = n
    return name
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == 'Address')
iexe, isamp, isrc = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
sass = [(r[0], r[isrc], int(r[iexe] or 0), int(r[isamp] or 0)) for r in rows if r and r[0].startswith('0x') and len(r) > isamp]
# address -> source line via cuda,sass view
out2 = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows2 = list(csv.reader(io.StringIO(out2)))
addr2line = {}
cur_file = None; cur_line = None
for r in rows2:
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if not r: continue
    if r[0].isdigit(): cur_line = int(r[0]); continue
    if r[0] == '' and len(r) > 2 and r[2].startswith('0x'): addr2line[r[2]] = (cur_file, cur_line)
ph_exe = collections.Counter(); ph_samp = collections.Counter(); ph_static = collections.Counter()
cur = 'prologue'
for a, s, e, sm in sorted(sass, key=lambda x: int(x[0], 16)):
    fl = addr2line.get(a)
    if fl and fl[0] == 'mdn_fused.cuh' and fl[1] >= KSTART: cur = phase_of(fl[1])
    ph_exe[cur] += e; ph_samp[cur] += sm; ph_static[cur] += 1
te, ts = sum(ph_exe.values()), sum(ph_samp.values())
print(f"total executed {te}  samples {ts}")
for k in ph_exe:
    print(f"  {k:14s} static {ph_static[k]:5d}  executed {ph_exe[k]:>11d} ({100*ph_exe[k]/te:5.1f}%)  time {100*ph_samp[k]/max(ts,1):5.1f}%")
