"""dev: static SASS instruction count per CUDA source line of one kernel (nvdisasm line info of the in-tree libdge.so).
Usage: sass_lines.py <file.cu> <kernel-substring> <first-line> <last-line>"""
import collections, os, re, subprocess, sys, tempfile

cu, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "drl_graph_exploration_b200", "libdge.so")], cwd=tmp, capture_output=True)
base = os.path.basename(cu).replace(".cu", "")
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, base + ".sm_100a.cubin")], capture_output=True, text=True).stdout
cur, infn = None, False
cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for l in sass.split("\n"):
    if l.strip().startswith(".section") and ".text." in l:
        infn = kern in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur and infn:
        cnt[cur] += 1; ops[cur][m.group(2).split(".")[0]] += 1
src = open(cu).read().split("\n")
tot = 0
for ln in range(lo, hi + 1):
    k = (os.path.basename(cu), ln)
    if cnt[k]:
        tot += cnt[k]
        print(f"{ln:5d} {cnt[k]:4d} {dict(ops[k].most_common(5))}  {src[ln - 1].strip()[:80]}")
print("total", tot)
