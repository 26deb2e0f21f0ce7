#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific SASS instructions in librdg_b200.so (B200_PROFILING.md: tcgen05.mma = UTC*MMA,
tcgen05.ld / st = LDTM / STTM, TMA = UTMALDG / UBLKCP, cp.async = LDGSTS, tcgen05.commit = UTCBAR).
usage: python tools/sass_summary.py > profiles/rN_sass_instruction_counts.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "pr-disagg-radar-gan_b200", "rdg_b200", "librdg_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ("UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "LDGSTS", "HMMA")
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for op in OPS:
            if re.search(r"\b" + op + r"[.\s]", line):
                cnt[cur][op] += 1
names = subprocess.run(["cu++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
rows = []
for mangled, name in zip(cnt, names):
    c = cnt[mangled]
    if any(c[k] for k in ("UTCHMMA", "UTCQMMA", "UTCMMA", "UTMALDG", "LDTM", "STTM", "UBLKCP")):
        name = re.sub(r"\((?:const |<unnamed>|CUtensorMap|TcConvArgs|rdg).*", "", name).replace("<unnamed>::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
        rows.append((name, c))
rows.sort(key=lambda r: (-(r[1]["UTCHMMA"] + r[1]["UTCQMMA"] + r[1]["UTCMMA"]), r[0]))
print("# SASS instruction counts per kernel of librdg_b200.so (cuobjdump -sass; tools/sass_summary.py)")
print(f"{'kernel':84s} {'UTC*MMA':>8s} {'LDTM':>5s} {'STTM':>5s} {'UTMALDG':>8s} {'UBLKCP':>7s} {'LDGSTS':>7s} {'UTCBAR':>7s} {'HMMA':>5s}")
tot = collections.Counter()
for name, c in rows:
    mma = c["UTCHMMA"] + c["UTCQMMA"] + c["UTCMMA"]
    print(f"{name[:84]:84s} {mma:8d} {c['LDTM']:5d} {c['STTM']:5d} {c['UTMALDG']:8d} {c['UBLKCP']:7d} {c['LDGSTS']:7d} {c['UTCBAR']:7d} {c['HMMA']:5d}")
    tot.update(c)
print(f"{'TOTAL (' + str(len(rows)) + ' kernels)':84s} {tot['UTCHMMA'] + tot['UTCQMMA'] + tot['UTCMMA']:8d} {tot['LDTM']:5d} {tot['STTM']:5d} {tot['UTMALDG']:8d} {tot['UBLKCP']:7d} {tot['LDGSTS']:7d} {tot['UTCBAR']:7d} {tot['HMMA']:5d}")
