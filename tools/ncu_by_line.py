"""Attributes executed warp-instructions of a profiled kernel to CUDA source lines.

    python tools/ncu_by_line.py <report.ncu-rep> <cubin> <kernel-mangled-substring> [unit]

Joins `ncu --page source --csv` (per-SASS-instruction executed counts) with `nvdisasm -g`
(per-instruction file/line) by instruction order.
"""
import collections, csv, io, re, subprocess, sys

rep, cubin, kname = sys.argv[1:4]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l and l.rstrip().endswith(":"))
loc, lines = ("?", 0), []
for l in dis[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), loc, m.group(2).strip()))
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
print(f"nvdisasm instructions {len(lines)}, ncu instructions {len(ins)}")
n = min(len(lines), len(ins))
tot = sum(c for _, c, _ in ins)
cnt = collections.Counter(c for _, c, _ in ins if c > tot / 1e5)
unit = float(sys.argv[4]) if len(sys.argv) > 4 else cnt.most_common(1)[0][0]
by_line, by_file, smp_line = collections.Counter(), collections.Counter(), collections.Counter()
for (off, loc, text), (s, c, k) in zip(lines[:n], ins[:n]):
    by_line[loc] += c
    by_file[loc[0]] += c
    smp_line[loc] += k
rest = sum(c for _, c, _ in ins[n:])
print(f"total {tot} ({tot / unit:.1f}/step), attributed {tot - rest}, in sub-functions {rest} ({rest / unit:.1f}/step)")
for f, c in by_file.most_common():
    print(f"  {f:<24} {c / unit:8.1f}/step {100 * c / tot:5.1f}%")
ts = sum(smp_line.values())
print("top lines:")
for loc, c in by_line.most_common(45):
    print(f"  {loc[0]:<22}:{loc[1]:<5} {c / unit:7.1f}/step {100 * c / tot:5.1f}%  samples {100 * smp_line[loc] / max(ts, 1):5.1f}%")
if len(sys.argv) > 5 and sys.argv[5] == "all":  # every line with >= 0.3 instructions per step, in source order
    print("all lines:")
    for loc in sorted(by_line):
        if by_line[loc] / unit >= 0.3:
            print(f"  {loc[0]:<22}:{loc[1]:<5} {by_line[loc] / unit:7.1f}/step  samples {100 * smp_line[loc] / max(ts, 1):5.1f}%")
if len(sys.argv) > 5 and sys.argv[5] == "sass":  # the executed SASS in address order with per-step counts
    for (off, loc, text), (s, c, k) in zip(lines[:n], ins[:n]):
        if c / unit >= 0.05:
            print(f"{off:6x} {c / unit:6.2f} {100 * k / max(ts, 1):5.2f}%  {loc[0]}:{loc[1]:<4} {text}")
