"""dev: per CUDA source line of one captured kernel -- executed warp instructions, stall samples, fp64 share (from the SASS opcode) -- joined
through the nvdisasm line info of the in-tree libdge.so.  Usage: ncu_lines.py rep file.cu [top]"""
import collections, csv, io, os, re, subprocess, sys
rep, cu = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]
kname = rows[2][h.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(srows) if "# Samples" in r)
sh = srows[hi]; ci = {c: i for i, c in enumerate(sh)}
cub = "/tmp/_ncu_sum_cub"; os.makedirs(cub, exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "drl_graph_exploration_b200", "libdge.so")], cwd=cub, capture_output=True)
base = os.path.basename(cu).replace(".cu", "")
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(cub, base + ".sm_100a.cubin")], capture_output=True, text=True).stdout
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
agg = collections.defaultdict(lambda: [0, 0, 0]); b0 = None; tot = [0, 0, 0]
for r in srows[hi + 1:]:
    if len(r) < len(sh):
        continue
    try:
        a = int(r[0], 16) if not r[0].isdigit() else int(r[0]); s = int(r[ci["# Samples"]]); ie = int(r[ci["Instructions Executed"]])
    except ValueError:
        continue
    b0 = a if b0 is None else b0
    op = r[ci["Source"]].split()
    op = [t for t in op if not t.startswith("@")][0] if op else ""
    f64 = ie if re.match(r"^(DFMA|DADD|DMUL|DSETP|MUFU\.RCP64H|DMNMX|F2F\.F64|I2F\.F64|F2I.*F64)", op) else 0
    g = agg[addr2line.get(a - b0, ("?", 0))]; g[0] += ie; g[1] += s; g[2] += f64
    tot[0] += ie; tot[1] += s; tot[2] += f64
lines = open(cu).read().split("\n")
print(f"kernel {kname}: {tot[0]} warp instructions, {tot[2]} fp64 ({100.0 * tot[2] / max(tot[0], 1):.1f} %), {tot[1]} samples")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1 if os.environ.get("BY_SAMPLES") else 0])[:top]:
    txt = lines[k[1] - 1].strip()[:90] if k[0] == os.path.basename(cu) and 0 < k[1] <= len(lines) else ""
    print(f"{k[0]}:{k[1]:4d} inst {v[0]:10d} ({100.0 * v[0] / tot[0]:5.1f} %) fp64 {v[2]:9d} samples {v[1]:6d}  {txt}")
