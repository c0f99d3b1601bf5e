"""SASS mnemonics that prove which hardware paths the shipped kernels use (B200_PROFILING.md: UBLKCP = bulk-copy TMA, SYNCS = mbarrier,
IMMA = int8 tensor-core MMA), counted per kernel of the built library with cuobjdump. usage: python profiles/sass_evidence.py [libv5ela.so]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "fake-video-detection-engine_b200", "v5ela", "libv5ela.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ("UBLKCP", "SYNCS", "IMMA", "IDP", "ATOMS", "BAR.SYNC", "LDS", "STS", "LDG", "STG", "MUFU", "DFMA")
name, counts, order = None, collections.defaultdict(collections.Counter), []
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        order.append(name)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and name:
        op = m.group(1)
        counts[name]["total"] += 1
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                counts[name][w] += 1
print(f"# cuobjdump -sass {os.path.basename(so)}: instructions per kernel, and how many of them are the listed mnemonics")
print(f"{'kernel':64s} {'total':>6s} " + " ".join(f"{w:>8s}" for w in WATCH))
for n in order:
    c = counts[n]
    print(f"{n[:64]:64s} {c['total']:6d} " + " ".join(f"{c[w]:8d}" for w in WATCH))
