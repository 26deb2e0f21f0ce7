#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS source page) : key metrics, samples per warp role / barrier.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second',
        'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__cycles_elapsed.avg.per_second', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum']
print("kernel:", vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h} [{units[i]}] = {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr): continue
    if r == hdr: break          # next captured launch: keep the first table only
    data.append(r)
iS = hdr.index('# Samples'); iSrc = hdr.index('Source')
stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print("total samples", tot)
c = collections.Counter()
for r in data:
    for k in stall:
        if r[k] not in ('', '0'):
            c[hdr[k][6:]] += int(r[k])
print("stall mix:", c.most_common(8))
print("top instructions:")
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]):
    r = data[i]
    st = {hdr[k][6:]: int(r[k]) for k in stall if r[k] not in ('', '0')}
    print(f"  {i:5d} {int(r[iS]):7d} {r[iSrc].strip()[:80]:80s} {sorted(st.items(), key=lambda kv: -kv[1])[:2]}")

# ---- mbarrier wait attribution (pass "name=hexoffset,..." as argv[3]): samples on the try_wait + following 2 rows
if len(sys.argv) > 3:
    names = sorted(((int(v, 16), k) for k, v in (kv.split('=') for kv in sys.argv[3].split(','))))
    def nm(off):
        best = '?'
        for o, k in names:
            if off >= o: best = k
        return best
    agg = collections.Counter()
    for i, r in enumerate(data):
        m = re.search(r'SYNCS\.PHASECHK\.TRANS64(\.TRYWAIT)? (\w+), \[(\w+)\+URZ(\+0x([0-9a-f]+))?\]', r[iSrc])
        if m:
            off = int(m.group(5), 16) if m.group(5) else 0
            agg[nm(off)] += sum(int(data[j][iS]) for j in range(i, min(i + 3, len(data))))
    print("samples at mbarrier waits:", dict(agg))
