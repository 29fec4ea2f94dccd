"""Summarises an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [unit]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_average_branch_targets_threads_uniform.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "sm__sass_inst_executed_op_shared_ld.sum", "sm__sass_inst_executed_op_shared_st.sum"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w} = {vals[i]} {rows[1][i]}")
st = []
for i, h in enumerate(hdr):
    if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
        try:
            st.append((float(vals[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
print("stalls per issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ci, si, sm = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
ins = []
for r in rows[2:]:
    try:
        ins.append((r[si].strip(), int(r[ci]), int(r[sm])))
    except (ValueError, IndexError):
        pass
tot = sum(n for _, n, _ in ins)
cnt = collections.Counter(n for _, n, _ in ins if n > tot / 1e5)
unit = float(sys.argv[2]) if len(sys.argv) > 2 else cnt.most_common(1)[0][0]
print(f"SASS lines {len(ins)}, warp-instructions {tot}, per-step unit {unit:.0f} -> {tot / unit:.1f} instructions per warp-step")
op, smp = collections.Counter(), collections.Counter()
for s, n, k in ins:
    o = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
    op[o] += n
    smp[o] += k
ts = sum(smp.values())
for o, n in op.most_common(22):
    print(f"  {o:<8} {n / unit:7.1f}/step {100 * n / tot:5.1f}% inst {100 * smp[o] / max(ts, 1):5.1f}% samples")
