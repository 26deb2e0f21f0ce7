#!/usr/bin/env python
"""One line per captured launch of an .ncu-rep: duration, tensor-pipe %, issue %, DRAM bytes, L2 hit rate, L2->SM bytes, occupancy.
usage: python tools/ncu_table.py report.ncu-rep [> profiles/summary.md]"""
import csv, io, re, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
C = {k: col(v) for k, v in dict(
    name="Kernel Name", grid="Grid Size", dur="gpu__time_duration.sum", tensor="sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    issue="smsp__issue_active.avg.pct_of_peak_sustained_active", dr="dram__bytes_read.sum", dw="dram__bytes_write.sum",
    l2hit="lts__t_sector_hit_rate.pct", l2tp="lts__throughput.avg.pct_of_peak_sustained_elapsed",
    xbar="l1tex__m_xbar2l1tex_read_bytes.sum", warps="sm__warps_active.avg.pct_of_peak_sustained_active",
    regs="launch__registers_per_thread", smem="launch__shared_mem_per_block_dynamic", inst="smsp__inst_executed.sum").items()}
def num(r, k):
    i = C[k]
    if i is None: return float("nan")
    try: return float(r[i].replace(",", ""))
    except ValueError: return float("nan")
def scale(r, k):            # bytes columns may be in Kbyte / Mbyte
    i = C[k]
    if i is None: return float("nan")
    u = units[i].lower()
    return num(r, k) * (1e3 if u.startswith("k") else 1e6 if u.startswith("m") else 1e9 if u.startswith("g") else 1)
print("| kernel | grid | us | tensor pipe % | issue % | L2 throughput % | L2->SM MB | DRAM rd+wr MB | L2 hit % | warps % | regs | dyn smem KB |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    if len(r) < len(hdr): continue
    name = re.sub(r"\(.*", "", r[C["name"]]).replace("void ", "").replace("<unnamed>::", "").replace("(int)", "").replace("(bool)", "")
    du = units[C["dur"]].lower(); dur = num(r, "dur") * (1e-3 if du.startswith("n") else 1e3 if du.startswith("m") else 1)
    print(f"| {name} | {r[C['grid']].replace(' ', '')} | {dur:.1f} | {num(r, 'tensor'):.1f} | {num(r, 'issue'):.1f} | {num(r, 'l2tp'):.1f} | "
          f"{scale(r, 'xbar') / 1e6:.1f} | {(scale(r, 'dr') + scale(r, 'dw')) / 1e6:.1f} | {num(r, 'l2hit'):.1f} | {num(r, 'warps'):.1f} | "
          f"{num(r, 'regs'):.0f} | {scale(r, 'smem') / 1e3:.0f} |")
