"""Executed warp-instructions per source line (ncu source page joined with nvdisasm line info).
usage: python profiles/ncu_by_line.py rep so file.cuh first_line last_line"""
import collections, csv, io, os, re, subprocess, sys

# which instantiation of the fused kernel the capture holds (mangled-name fragment): <FAST=1, TEXHIST=0> by default
KERNEL = os.environ.get("V5_NCU_KERNEL", "ela_fused_kernelILb1ELb0")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_function import line_map

rep, so, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, ie, ins = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Source")
lm = line_map(so)
base = int(rows[2][ia], 16)
cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) <= ie: continue
    loc = lm.get(int(r[ia], 16) - base, (None, ""))[0]
    if not loc or loc[0] != fname or not (lo <= loc[1] <= hi): continue
    e = int(r[ie] or 0)
    cnt[loc[1]] += e
    t = r[ins].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[loc[1]][op.split(".")[0]] += e
text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(so))), "csrc", fname)).read().split("\n")
tot = sum(cnt.values())
print("total in range:", tot)
for ln in sorted(cnt):
    if cnt[ln] * 200 < tot: continue
    print(f"{ln:4d} {cnt[ln]/1e6:9.1f}M {cnt[ln]*100/tot:5.1f}%  {', '.join(f'{o}:{c*100//cnt[ln]}' for o,c in ops[ln].most_common(4)):40s} | {text[ln-1].strip()[:90]}")
