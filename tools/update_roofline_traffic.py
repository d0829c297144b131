"""Writes profiles/roofline_traffic.json from an ncu report of the centre kernel (read here, no GPU needed):
DRAM bytes per launch, pipe figures, and the identity of the kernel sources / commit the capture belongs to
(bench.py reports `roofline.traffic` only while the sources are still the captured ones).
usage: python tools/update_roofline_traffic.py gpurun_out/X.ncu-rep "workload text" profiles/SUMMARY.md"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_sha)
csv.field_size_limit(10 ** 9)
rep, workload, src = sys.argv[1], sys.argv[2], sys.argv[3]
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units, row = raw[0], raw[1], raw[2]
def val(name, scale=1.0):
    i = hdr.index(name); u = units[i].lower(); v = float(row[i])
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return v * mult * scale
out = {
    "kernel": row[hdr.index("Kernel Name")],
    "workload": workload,
    "dram_bytes_read": int(val("dram__bytes_read.sum")),
    "dram_bytes_write": int(val("dram__bytes_write.sum")),
    "fp64_pipe_active_pct": round(val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
    "issue_slots_busy_pct": round(val("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
    "warp_instructions": int(val("smsp__inst_executed.sum")),
    "duration_ms_under_ncu": round(val("gpu__time_duration.sum") / ({"nsecond": 1e6, "ns": 1e6, "usecond": 1e3, "us": 1e3, "msecond": 1, "ms": 1, "second": 1e-3, "s": 1e-3}[units[hdr.index("gpu__time_duration.sum")].lower()]), 4),
    "kernel_sha": bench.kernel_sha(),
    "commit": subprocess.run(["git", "-C", ROOT, "rev-parse", "--short=12", "HEAD"], capture_output=True, text=True).stdout.strip(),
    "source": src,
}
json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=2)
print(json.dumps(out, indent=2))
