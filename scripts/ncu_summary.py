"""Summarises an ncu report of the fused kernel: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [--lines N]

Prints duration, instruction totals, issue utilisation, pipe utilisation, stall reasons, DRAM bytes, and the
executed warp-instructions per source line (top N) and per opcode.  Reads the report with `ncu -i` (no GPU needed).
"""
import collections
import csv
import io
import re
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        m = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print("kernel:", m.get("Kernel Name"), "grid", m.get("launch__grid_size"), "block", m.get("launch__block_size"))
        keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
                "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
                "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
                "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
                "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_shared_st.sum",
                "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
                "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
        for k in keys:
            if k in m:
                print(f"  {k:75s} {m[k]:>16s} {u[k]}")
        print("  stalls (warps per issue):")
        st = [(float(v), k) for k, v in m.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
        for v, k in sorted(st, reverse=True)[:9]:
            print(f"    {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:28s} {v:6.2f}")
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]))))
    cur = None
    by_line = collections.Counter()
    by_op = collections.Counter()
    text = {}
    seen_addr = set()
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8:
            continue
        if r[0].isdigit():
            try:
                by_line[(cur, int(r[0]))] += int(r[7])
                text[(cur, int(r[0]))] = r[1].strip()[:100]
            except ValueError:
                pass
        elif r[0] == "" and r[2].startswith("0x"):
            if r[2] in seen_addr:      # an inlined instruction is listed under every source line of its call chain: count it once
                continue
            seen_addr.add(r[2])
            try:
                n = int(r[7])
            except ValueError:
                continue
            parts = r[3].split()
            if parts and parts[0].startswith("@"):
                parts = parts[1:]
            if parts:
                by_op[parts[0].split(".")[0]] += n
    tot = sum(by_line.values())
    print("executed warp-instructions summed over source lines (inlined code counts once per line of its call chain):", tot)
    print("top source lines:")
    for (f, ln), n in by_line.most_common(nlines):
        print(f"  {100 * n / tot:5.2f}% {n:>10d} {f}:{ln}  {text[(f, ln)]}")
    print("opcodes:")
    tot_op = sum(by_op.values())
    for op, n in by_op.most_common(30):
        print(f"  {op:10s} {n:>11d} {100 * n / max(tot_op, 1):5.1f}%")
    # phase ranges from '// -- Pn' style markers are file specific; print a coarse histogram by 25-line buckets instead
    buckets = collections.Counter()
    for (f, ln), n in by_line.items():
        buckets[(f, ln // 25 * 25)] += n
    print("by 25-line bucket:")
    for (f, b0), n in sorted(buckets.items()):
        if n * 200 > tot:
            print(f"  {f}:{b0:4d}-{b0 + 24:4d} {100 * n / tot:5.1f}%")


if __name__ == "__main__":
    main()
