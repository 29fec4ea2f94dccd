"""Stall-reason samples per source file / per SASS instruction of a profiled kernel.
    python tools/ncu_stalls.py <report.ncu-rep> <cubin> <kernel-mangled-substring> [top]
"""
import collections, csv, io, re, subprocess, sys

rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
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
names = ["stall_branch_resolving", "stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_not_selected",
         "stall_selected", "stall_math", "stall_dispatch", "stall_mio", "stall_lg", "stall_barrier", "stall_misc"]
idx = {n: h.index(n) for n in names if n in h}
ci = h.index("Instructions Executed")
recs = []
for r in rows[2:]:
    try:
        recs.append((int(r[ci]), {n: int(r[i] or 0) for n, i in idx.items()}))
    except (ValueError, IndexError):
        pass
n = min(len(lines), len(recs))
by_file = collections.defaultdict(collections.Counter)
tot = collections.Counter()
per = []
for (off, loc, text), (c, st) in zip(lines[:n], recs[:n]):
    for k, v in st.items():
        by_file[loc[0]][k] += v
        tot[k] += v
    per.append((sum(v for k, v in st.items() if k != "stall_selected"), off, loc, text[:60], st))
allsum = sum(tot.values())
print("total samples", allsum, {k: f"{100 * v / allsum:.1f}%" for k, v in tot.most_common()})
for f, cnt in sorted(by_file.items(), key=lambda kv: -sum(kv[1].values())):
    s = sum(cnt.values())
    print(f"{f:<24} {100 * s / allsum:5.1f}%  " + " ".join(f"{k[6:]}={100 * v / allsum:.1f}" for k, v in cnt.most_common(6)))
print("top stalled instructions (not counting `selected`):")
for s, off, loc, text, st in sorted(per, reverse=True)[:top]:
    print(f"  {off:6x} {100 * s / allsum:5.2f}%  {loc[0]}:{loc[1]:<4} {text:<60} " +
          " ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v))
