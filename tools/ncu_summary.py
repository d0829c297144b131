"""Summarise an ncu report (read here, no GPU needed) into profiles/<name>.md:
per-kernel headline metrics + instruction share per code region of desc.cu.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/NAME.md [launches.csv]"""
import csv
import io
import subprocess
import sys

csv.field_size_limit(10 ** 9)
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread-instr"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread-instr"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread-instr"),
    ("smsp__inst_executed_op_shared_atom.sum", "shared atomics (warp instr)"),
    ("smsp__sass_inst_executed_op_shared_ld.sum", "shared loads (warp instr)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instr"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard (smem/MUFU)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, rows = raw[0], raw[1], raw[2:]
    ki = hdr.index("Kernel Name")
    lines = ["# ncu summary of `%s`" % rep.split("/")[-1], "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun; "
             "read here with `ncu -i ... --page raw --csv` (tools/ncu_summary.py).", ""]
    lines.append("| metric | " + " | ".join("`%s`" % r[ki][:48] for r in rows) + " |")
    lines.append("|---|" + "---|" * len(rows))
    for m, label in METRICS:
        if m in hdr:
            i = hdr.index(m)
            lines.append("| %s (%s) | " % (label, units[i]) + " | ".join(r[i] for r in rows) + " |")
    # instruction share per source line (top 25) for each kernel
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    cur_file = cur_fn = None
    agg = {}
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            cur_fn = r[1]
        elif r[0].isdigit():
            try:
                agg[(cur_fn, cur_file, int(r[0]))] = (float(r[7]), float(r[6]), r[1].strip())
            except (ValueError, IndexError):
                pass
    for fn in sorted(set(k[0] for k in agg)):
        tot = sum(v[0] for k, v in agg.items() if k[0] == fn) or 1
        tots = sum(v[1] for k, v in agg.items() if k[0] == fn) or 1
        lines += ["", "## hottest source lines of `%s`" % fn[:80], "",
                  "total %.1f M warp instructions, %d stall samples" % (tot / 1e6, tots), "",
                  "| % instr | % samples | where | source |", "|---|---|---|---|"]
        top = sorted([(v[0], v[1], k, v[2]) for k, v in agg.items() if k[0] == fn], reverse=True)[:25]
        for inst, samp, k, s in top:
            lines.append("| %.1f | %.1f | %s:%d | `%s` |" % (100 * inst / tot, 100 * samp / tots, k[1], k[2], s[:90].replace("|", "\\|")))
    if len(sys.argv) > 3:
        lines += ["", "## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised)", "",
                  "| kernel | ns |", "|---|---|"]
        rr = list(csv.reader(l for l in open(sys.argv[3]) if l.startswith('"')))
        h = rr[0]
        a, b = h.index("Kernel Name"), h.index("Metric Value")
        for r in rr[1:]:
            lines.append("| `%s` | %s |" % (r[a][:70], r[b]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
