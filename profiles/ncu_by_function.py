"""Join the SASS-level ncu source page with nvdisasm line info: executed warp-instructions and stall samples per source function.
usage: python profiles/ncu_by_function.py <rep.ncu-rep> <libv5ela.so> [pixels_per_launch]"""
import bisect, collections, csv, io, os, re, subprocess, sys, tempfile

# which instantiation of the fused kernel the capture holds (mangled-name fragment): <FAST=1, TEXHIST=0> by default
KERNEL = os.environ.get("V5_NCU_KERNEL", "ela_fused_kernelILb1ELb0")

def line_map(so):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    dis = ""                                                 # one cubin per translation unit: the kernel is in one of them
    for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    m, cur, infn = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            infn = KERNEL in ln
        if not infn:
            continue
        mm = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        mm = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S.*?);", ln)
        if mm:
            m[int(mm.group(1), 16)] = (cur, mm.group(2))
    return m

def functions(path):
    out = []
    for i, t in enumerate(open(path).read().split("\n"), 1):
        mm = re.match(r"^(?:V5_DEV|V5_HOSTDEV|inline|__global__|__device__ __forceinline__)\s+[\w\s\*&:<>]*?\b(\w+)\(", t)
        if mm:
            out.append((i, mm.group(1)))
    return out

def main(rep, so, px=None):
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, ie, isamp, ins = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    lm = line_map(so)
    base = int(rows[2][ia], 16)
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(so))), "csrc")
    fn = {f: functions(os.path.join(csrc, f)) for f in ("v5ela_device.cuh", "v5ela_workitem.cuh", "v5ela_dctmma.cuh")}
    ex, sm, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    tot_e = tot_s = 0
    for r in rows[2:]:
        if len(r) <= ie:
            continue
        off = int(r[ia], 16) - base
        e, s = int(r[ie] or 0), int(r[isamp] or 0)
        loc = lm.get(off, (None, ""))[0]
        name = "?"
        if loc and loc[0] in fn:
            starts = [a for a, _ in fn[loc[0]]]
            j = bisect.bisect_right(starts, loc[1]) - 1
            name = fn[loc[0]][j][1] if j >= 0 else loc[0]
        elif loc:
            name = loc[0]
        ex[name] += e; sm[name] += s; tot_e += e; tot_s += s
        op = r[ins].split()[0] if not r[ins].strip().startswith("@") else r[ins].split()[1]
        ops[name][op.split(".")[0]] += e
    print(f"total executed warp-instructions {tot_e}  samples {tot_s}" + (f"  = {tot_e/px:.3f} warp-inst/px" if px else ""))
    for k, v in ex.most_common(24):
        top = ", ".join(f"{o}:{c*100//max(v,1)}%" for o, c in ops[k].most_common(6))
        print(f"{v:12d} {v*100/tot_e:5.1f}%  samples {sm[k]*100/max(tot_s,1):5.1f}%  {k:22s} {top}")
    allops = collections.Counter()
    for k in ops: allops.update(ops[k])
    print("opcode mix:", ", ".join(f"{o}:{c*100/tot_e:.1f}%" for o, c in allops.most_common(16)))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else None)
