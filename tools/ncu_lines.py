"""Aggregate ncu per-instruction stall samples by CUDA source line.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel substring> <cubin file> [top N] [mangled substring]
(ncu's CSV source page is SASS only; the SASS -> line map comes from `nvdisasm -g` on the cubin that
`cuobjdump -xelf all libfcest_b200.so` extracts -- instruction order is identical.)"""
import csv, re, subprocess, sys, collections

rep, kname, cubin = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dname = sys.argv[5] if len(sys.argv) > 5 else kname
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": [], "hdr": None}; kern.append(cur); continue
    if cur is None: continue
    if r and r[0] == "Address": cur["hdr"] = r; continue
    if cur["hdr"] and len(r) == len(cur["hdr"]): cur["rows"].append(r)
k = [k for k in kern if kname in k["name"]][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate function
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and dname in l][0]
lines, cur_line = [], None
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or l.startswith(".text."): 
        if lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur_line)
h = k["hdr"]; isamp = h.index("# Samples"); isrc = h.index("Source")
n = min(len(lines), len(k["rows"]))
print(f"{k['name'][:70]}: {len(k['rows'])} SASS rows, {len(lines)} disasm instrs")
agg = collections.Counter(); stall = collections.defaultdict(collections.Counter)
scols = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
for r, ln in zip(k["rows"][:n], lines[:n]):
    s = int(r[isamp] or 0); agg[ln] += s
    for i in scols:
        v = int(r[i] or 0)
        if v: stall[ln][h[i][6:]] += v
tot = sum(agg.values())
srccache = {}
def src(ln):
    if ln is None: return ""
    f, n_ = ln
    import glob
    if f not in srccache:
        c = glob.glob(f"/root/repo/fcest_b200/csrc/{f}")
        srccache[f] = open(c[0]).read().splitlines() if c else []
    return srccache[f][n_ - 1].strip()[:90] if n_ - 1 < len(srccache[f]) else ""
for ln, s in agg.most_common(top):
    st = ", ".join(f"{a}:{b}" for a, b in stall[ln].most_common(3))
    print(f"{s:7d} {100*s/tot:5.1f}%  {str(ln):28s} {src(ln):90s} | {st}")
