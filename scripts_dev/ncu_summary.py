"""Summarise an .ncu-rep (read here, without a GPU): headline metrics + stall samples per CUDA source line
(joined through nvdisasm line info of the in-tree libdge.so).  Usage: ncu_summary.py rep kernel_cu_file out.md"""
import collections, csv, io, os, re, subprocess, sys

rep, cu, out = sys.argv[1], sys.argv[2], sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = [f"# ncu summary: {os.path.basename(rep)}", "", "(`ncu --set full --clock-control none --import-source on`; one launch; cold-cache, serialised -- use for shares and stall reasons, not for absolute time)", ""]
for r in rows[2:]:
    lines.append("| metric | value | unit |"); lines.append("|---|---|---|")
    for k in keys:
        if k in h:
            i = h.index(k)
            lines.append(f"| {k} | {r[i]} | {u[i]} |")
    lines.append("")
# per source line
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(srows) if "# Samples" in r)
sh = srows[hi]; ci = {c: i for i, c in enumerate(sh)}
cub = "/tmp/_ncu_sum_cub"; os.makedirs(cub, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "drl_graph_exploration_b200", "libdge.so")], cwd=cub, capture_output=True)
base = os.path.basename(cu).replace(".cu", "")
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(cub, base + ".sm_100a.cubin")], capture_output=True, text=True).stdout
kname = rows[2][h.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0]
addr2line, cur, infn = {}, None, False
for l in dis.split("\n"):
    if ".text." in l and l.strip().startswith(".section"):
        infn = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur and infn:
        addr2line[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0]); b0 = None
for r in srows[hi + 1:]:
    if len(r) < len(sh):
        continue
    try:
        a = int(r[0], 16) if not r[0].isdigit() else int(r[0]); s = int(r[ci["# Samples"]])
    except ValueError:
        continue
    b0 = a if b0 is None else b0
    g = agg[addr2line.get(a - b0, ("?", 0))]; g[0] += s
    for j, c in enumerate(["stall_barrier", "stall_long_sb", "stall_wait", "stall_short_sb"]):
        try:
            g[j + 1] += int(r[ci[c]])
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
text = open(cu).read().split("\n")
lines += ["## warp-stall samples by CUDA source line (top 25)", "", "| samples | % | barrier | long_sb | wait | short_sb | line |", "|---|---|---|---|---|---|---|"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    code = text[k[1] - 1].strip()[:100].replace("|", "\\|") if k[0] == os.path.basename(cu) and 0 < k[1] <= len(text) else ""
    lines.append(f"| {v[0]} | {100 * v[0] / tot:.1f} | {v[1]} | {v[2]} | {v[3]} | {v[4]} | `{k[0]}:{k[1]}` {code} |")
rng = os.environ.get("LINES")
if rng:   # LINES=a-b : every sampled line of that source range, in source order
    lo, hi_ = (int(v) for v in rng.split("-"))
    lines += ["", f"## lines {lo}-{hi_} in source order", "", "| samples | barrier | long_sb | wait | short_sb | line |", "|---|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: kv[0][1]):
        if k[0] == os.path.basename(cu) and lo <= k[1] <= hi_:
            lines.append(f"| {v[0]} | {v[1]} | {v[2]} | {v[3]} | {v[4]} | `{k[1]}` {text[k[1] - 1].strip()[:110]} |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
