#!/usr/bin/env python
"""Summarise an .ncu-rep: per-kernel headline metrics, and (--stalls N) the N hottest SASS lines
with their dominant stall reasons for one kernel instance (--index i)."""
import argparse, csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'sm__cycles_elapsed.max', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active']

def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout

def I(x):
    try: return int(float(x.replace(',', '')))
    except Exception: return 0

ap = argparse.ArgumentParser()
ap.add_argument('rep')
ap.add_argument('--stalls', type=int, default=0)
ap.add_argument('--index', type=int, default=0)
a = ap.parse_args()
rows = list(csv.reader(io.StringIO(run(['-i', a.rep, '--page', 'raw', '--csv']))))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
for n, d in enumerate(data):
    print(f"---- [{n}] {d[idx['Kernel Name']][:90]}")
    for w in WANT:
        if w in idx:
            print(f"   {w:72s} {d[idx[w]]:>16s} {units[idx[w]]}")
if a.stalls:
    name = data[a.index][idx['Kernel Name']]
    out = run(['-i', a.rep, '--page', 'source', '--csv', '--launch-skip', str(a.index), '--launch-count', '1'])
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) >= len(hdr) - 2]
    half = len(data) // 2 if len(data) > 1 and data[0][1] == data[len(data)//2][1] else len(data)
    data = data[:half]
    ix = {h: i for i, h in enumerate(hdr)}
    si = ix['# Samples']
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(I(d[si]) for d in data)
    agg = {s: sum(I(d[ix[s]]) for d in data) for s in stalls}
    print(f"\n==== stalls for [{a.index}] {name[:80]}: {tot} samples")
    print("   by reason:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for n, d in enumerate(data): d.append(n)
    for d in sorted(data, key=lambda d: -I(d[si]))[:a.stalls]:
        st = {s: I(d[ix[s]]) for s in stalls if I(d[ix[s]]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{d[-1]:5d} {I(d[si]):6d} {d[ix['Source']].strip()[:66]:66s} {st}")
