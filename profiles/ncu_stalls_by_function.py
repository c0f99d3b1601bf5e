"""Stall samples per source function and reason (ncu source page joined with nvdisasm line info).
usage: V5_NCU_KERNEL=<mangled fragment> python profiles/ncu_stalls_by_function.py <rep.ncu-rep> <libv5ela.so> [lines]
With `lines` the table is per source line (file:line) instead of per function."""
import collections, csv, io, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_by_function import line_map, functions

rep, so = sys.argv[1], sys.argv[2]
per_line = len(sys.argv) > 3 and sys.argv[3] == "lines"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
reasons = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
lm = line_map(so)
csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fake-video-detection-engine_b200", "csrc")
fn_of = {}
def func(loc):
    if loc is None:
        return "?"
    f, l = loc
    if per_line:
        return f"{f}:{l}"
    if f not in fn_of:
        p = os.path.join(csrc, f)
        fn_of[f] = functions(p) if os.path.exists(p) else []
    name = f
    for start, nm in fn_of[f]:
        if start <= l:
            name = nm
        else:
            break
    return name
agg = collections.defaultdict(lambda: collections.Counter())
base = None
for r in rows[2:]:
    try:
        a = int(r[ia], 16)
    except ValueError:
        continue
    if base is None:
        base = a
    loc = lm.get(a - base, (None, ""))[0]
    k = func(loc)
    agg[k]["inst"] += int(r[ie] or 0)
    agg[k]["samples"] += int(r[isamp] or 0)
    for i, nm in reasons:
        agg[k][nm] += int(r[i] or 0)
tot = sum(v["samples"] for v in agg.values())
tin = sum(v["inst"] for v in agg.values())
names = [nm for _, nm in reasons]
keep = [n for n in names if sum(v[n] for v in agg.values()) > 0.01 * tot]
print(f"samples {tot}  executed warp-instructions {tin}")
print(f"{'':28s} {'inst%':>6s} {'smp%':>6s}  " + " ".join(f"{n[:9]:>9s}" for n in keep))
allv = collections.Counter()
for v in agg.values():
    allv.update(v)
print(f"{'ALL':28s} {100.0:6.1f} {100.0:6.1f}  " + " ".join(f"{100 * allv[n] / tot:9.1f}" for n in keep))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:45]:
    print(f"{k[:28]:28s} {100 * v['inst'] / tin:6.1f} {100 * v['samples'] / tot:6.1f}  " + " ".join(f"{100 * v[n] / tot:9.2f}" for n in keep))
