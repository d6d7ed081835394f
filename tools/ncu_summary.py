#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, stall reasons, per-opcode dynamic
instruction counts and per-region stall samples.  usage: tools_ncu_summary.py rep.ncu-rep frames_per_launch"""
import collections, csv, io, subprocess, sys
rep, frames = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'sm__inst_executed.avg.per_cycle_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic']
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for i, h in enumerate(hdr):
    if h in keep:
        print(f"  {h:75s} {vals[i]} {units[i]}")
st = sorted(((float(vals[i]), h) for i, h in enumerate(hdr)
             if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')), reverse=True)
print("stall reasons (warps per issue-active cycle):")
for v, h in st[:9]:
    print(f"  {v:6.3f} {h.split('stalled_')[1].replace('_per_issue_active.ratio', '')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; ix = {h: i for i, h in enumerate(h2)}; data = rows[2:]
by_op, samp_op, tot, tots = collections.Counter(), collections.Counter(), 0, 0
for r in data:
    try:
        n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except Exception:
        continue
    w = r[ix["Source"]].split()
    op = (w[1] if w[0].startswith('@') else w[0]).split('.')[0]
    by_op[op] += n; samp_op[op] += s; tot += n; tots += s
print(f"dynamic warp-instructions per frame: {tot / frames:.1f} (static {len(data)})")
for op, n in by_op.most_common(22):
    print(f"  {op:10s} {n / frames:8.1f} /frame  {100 * samp_op[op] / max(tots, 1):5.1f}% of stall samples")
stall_cols = [h for h in h2 if h.startswith('stall_') and 'Not Issued' not in h]
chunk = 400
print("regions (SASS order):")
for s0 in range(0, len(data), chunk):
    seg = data[s0:s0 + chunk]
    ni = sum(int(r[ix["Instructions Executed"]] or 0) for r in seg)
    ns = sum(int(r[ix["# Samples"]] or 0) for r in seg)
    if ni / frames < 1 and ns < 100:
        continue
    ops, stc = collections.Counter(), collections.Counter()
    for r in seg:
        w = r[ix["Source"]].split()
        ops[(w[1] if w[0].startswith('@') else w[0]).split('.')[0]] += int(r[ix["Instructions Executed"]] or 0)
        for c in stall_cols:
            if r[ix[c]]:
                stc[c[6:]] += int(r[ix[c]])
    print(f"  [{s0:5d}] {ni / frames:7.1f} inst/frame {100 * ns / max(tots, 1):5.1f}% samples | "
          + ", ".join(f"{k}:{v / frames:.0f}" for k, v in ops.most_common(5)) + " | "
          + ", ".join(f"{k}:{100 * v / max(tots, 1):.1f}%" for k, v in stc.most_common(3)))
