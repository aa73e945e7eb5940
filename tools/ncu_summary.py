"""Summaries for profiles/: (a) per-kernel totals of an ncu launch list CSV, (b) key metrics of a --set full report."""
import csv, collections, subprocess, sys

def launches(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if r and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in data:
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v = v * 1e3 if u == "us" else v * 1e6 if u == "ms" else v
        k = d["Kernel Name"].split("(")[0][:90]
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {len(data)} launches, {tot/1e6:.3f} ms total device time (cold-cache, serialised: compare shares)")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]/1e6:.3f} | {100*v[1]/tot:.1f}% |")

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); hdr, units = rows[0], rows[1]
    want = {"gpu__time_duration.sum": "time", "dram__bytes_read.sum": "dram read", "dram__bytes_write.sum": "dram write",
            "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active": "DMMA pipe %",
            "sm__warps_active.avg.pct_of_peak_sustained_active": "warps active %", "launch__registers_per_thread": "regs",
            "launch__grid_size": "grid", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue active %",
            "lts__t_sector_hit_rate.pct": "L2 hit %", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "DRAM %",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem bank conflicts"}
    idx = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(want.values()) + " |\n|---|" + "---|" * len(want))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]][:70]
        cells = []
        for m in want:
            cells.append((r[idx[m]] + " " + units[idx[m]]).strip() if m in idx else "-")
        print(f"| `{name}` | " + " | ".join(cells) + " |")

if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else full)(sys.argv[2])
