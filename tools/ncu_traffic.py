"""Writes profiles/r02_traffic.json (the ncu-derived fields bench.py quotes in its roofline block) from
the two --set full captures made by tools/gpu_final.sh:

    python tools/ncu_traffic.py gpurun_out/prof_rollout_full.ncu-rep gpurun_out/prof_pointwise_full.ncu-rep [commit]
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, v = rows[0], rows[2]
    return lambda k: v[h.index(k)].replace(",", "")


def block(rep, population, warp_steps=None):
    g = raw(rep)
    f = lambda k: float(g(k))
    rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
    unit = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
    # the raw page reports bytes in a per-column unit (second header row)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    u = dict(zip(rows[0], rows[1]))
    rd *= unit.get(u["dram__bytes_read.sum"], 1.0)
    wr *= unit.get(u["dram__bytes_write.sum"], 1.0)
    inst = f("smsp__inst_executed.sum")
    fl = (f("smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed") +
          f("smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed") +
          2 * f("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed"))
    b = {
        "kernel": g("Kernel Name"),
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_individual": (rd + wr) / population, "population": population,
        "duration_under_ncu": g("gpu__time_duration.sum") + " " + u["gpu__time_duration.sum"],
        "registers_per_thread": int(f("launch__registers_per_thread")),
        "achieved_occupancy_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions": inst,
        "branch_targets_uniform_pct": f("smsp__sass_average_branch_targets_threads_uniform.pct"),
        "executed_fp32_flops_per_cycle": fl,
        "executed_fp32_frac_of_fma_peak": fl / (148 * 128 * 2),
        "pipe_fma_pct": f("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "source": os.path.basename(rep) + " (ncu --set full --clock-control none, one launch at the bench size)",
    }
    if warp_steps:
        # two denominators, both reported: (a) warp-steps the kernel actually executed (a warp steps while ANY of its
        # lanes still integrates) = the execution count of the step loop's own instructions in the source page;
        # (b) lane-steps / 32, which charges the lanes that ride along dead (9-10 % on C5) to the instruction count
        b["warp_instructions_per_32_lane_steps"] = inst / warp_steps
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(src)))
        ci = rows[1].index("Instructions Executed")
        counts = []
        for r in rows[2:]:
            try:
                counts.append(int(r[ci]))
            except (ValueError, IndexError):
                pass
        import collections
        tot = sum(counts)
        executed_warp_steps = collections.Counter(c for c in counts if c > tot / 1e5).most_common(1)[0][0]
        b["executed_warp_steps"] = executed_warp_steps
        b["warp_instructions_per_warp_step"] = inst / executed_warp_steps
        b["lane_waste"] = 1.0 - warp_steps / executed_warp_steps
    return b


if __name__ == "__main__":
    roll, pw = sys.argv[1:3]
    # warp-steps of the C5 launch: executed rollout-steps / 32 lanes, from the bench line of the same build
    with open(os.path.join(ROOT, "gpurun_out", "bench.json")) as fh:
        line = json.load(fh)
    warp_steps = line["config"]["executed_rollout_steps_per_step"] / 32.0
    out = {"k_rollout": block(roll, 65536, warp_steps), "k_pointwise": block(pw, 100000)}
    # the commit the CAPTURED build was made from (argv[3] when the capture is processed after later commits)
    commit = sys.argv[3] if len(sys.argv) > 3 else \
        subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    dirty = subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "multitreegp_b200"], capture_output=True,
                           text=True).stdout.strip()
    for b in out.values():
        b["commit"] = commit + ("+uncommitted" if dirty else "")
        b["dram_bytes_per_launch"] = b["dram_bytes_read"] + b["dram_bytes_write"]
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))
