"""Executed warp instructions per CUDA source line from an ncu report (companion of ncu_lines.py).
usage: python tools/ncu_inst.py <report.ncu-rep> <cubin> <mangled substring> <source file> [top N]"""
import csv, re, subprocess, sys, collections
rep, cubin, dname, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; data = []
for r in rows:
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(r)
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and dname in l][0]
lines, cur_line = [], None
for l in dis[start + 1:]:
    if l.startswith("//---------------------") or l.startswith(".text."):
        if lines: break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur_line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): lines.append(cur_line)
ie = hdr.index("Instructions Executed")
agg = collections.Counter()
for r, ln in zip(data, lines): agg[ln] += int(r[ie] or 0)
tot = sum(agg.values()); print("total warp instructions", tot)
src = open(srcfile).read().splitlines()
base = srcfile.split("/")[-1]
for ln, c in agg.most_common(top):
    txt = src[ln[1] - 1].strip()[:90] if ln and ln[0] == base else ''
    print(f"{c:9d} {100*c/tot:5.1f}% {ln} {txt}")
# cumulative by line ranges of 20
rng = collections.Counter()
for ln, c in agg.items():
    if ln and ln[0] == base: rng[ln[1] // 10 * 10] += c
print("by 10-line block:", sorted(rng.items()))
