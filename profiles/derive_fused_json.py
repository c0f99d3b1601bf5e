"""Derive profiles/fused_kernel_dram.json and fused_kernel_issue.json (what bench.py quotes as `roofline.traffic` and the
issue-slot figures) from an `ncu --set full` capture of one 256 x 1080p launch of the fused kernel.
usage: python profiles/derive_fused_json.py <file.ncu-rep> <source label> [extra output directory]
Run it on the tree the capture was taken from: both files are stamped with the hash of csrc/ (bench.py's source_hash) so that the
bench line can say whether its ncu figures describe the build it timed."""
import json, os, shutil, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_summary import raw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N, H, W = 256, 1080, 1920


def main(rep, label, extra=None):
    hdr, units, rows = raw(rep)
    row = [r for r in rows if "ela_fused_kernel" in dict(zip(hdr, r)).get("Kernel Name", "")][0]
    d = dict(zip(hdr, row))

    def val(k):
        v, u = float(d[k].replace(",", "")), units[hdr.index(k)].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    from source_hash import source_hash as _sh                      # the same function bench.py uses (comments and white space do not count)

    def source_hash():
        return _sh(os.path.join(ROOT, "fake-video-detection-engine_b200", "csrc"))

    sh = os.environ.get("V5_PROFILE_SOURCE_HASH") or source_hash()
    dram = {"kernel": "v5::ela_fused_kernel", "source": label, "source_hash": sh, "frames_per_launch": N, "height": H, "width": W,
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
            "dram_bytes_per_frame": (rd + wr) / N, "algorithmic_bytes_per_frame": 3 * H * W + 3144}
    issue = {"source": label, "source_hash": sh, "kernel_ms": val("gpu__time_duration.sum") * (1e-6 if units[hdr.index("gpu__time_duration.sum")].lower().startswith("ns") else (1e-3 if units[hdr.index("gpu__time_duration.sum")].lower().startswith("us") else 1.0)),
             "warp_instructions_per_pixel": val("smsp__inst_executed.sum") / (N * H * W),
             "issue_slots_busy_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "alu_pipe_pct": val("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
             "fma_pipe_pct_of_heavy_plus_lite": val("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
             "lsu_pipe_pct": val("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
             "dram_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
    for name, obj in (("fused_kernel_dram.json", dram), ("fused_kernel_issue.json", issue)):
        path = os.path.join(ROOT, "profiles", name)
        with open(path, "w") as f:
            json.dump(obj, f, indent=1)
        if extra:
            shutil.copy(path, os.path.join(extra, name))
    print(json.dumps({"dram": dram, "issue": issue}))


if __name__ == "__main__":
    main(*sys.argv[1:4])
