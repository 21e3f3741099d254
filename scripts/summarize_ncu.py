"""Summarise ncu outputs brought back from the GPU box into profiles/ (run here, no GPU needed).

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches_summary.md
  python scripts/summarize_ncu.py full gpurun_out/corr_prof.ncu-rep profiles/r01_corr_kernel_full.md [profiles/corr_traffic.json]
"""
import csv
import io
import json
import subprocess
import sys
from collections import defaultdict


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        rows.append((r["Kernel Name"], ns, r.get("Grid Size", ""), r.get("Block Size", "")))
    agg = defaultdict(lambda: [0, 0.0])
    for k, ns, *_ in rows:
        name = k.split("(")[0]
        agg[name][0] += 1
        agg[name][1] += ns
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, "
                f"serialised per-launch times: compare SHARES, not absolutes.\n\n{len(rows)} launches, {total/1e6:.3f} ms total\n\n")
        f.write("| kernel | launches | total ms | share | mean us |\n|---|---:|---:|---:|---:|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name}` | {n} | {ns/1e6:.3f} | {100*ns/total:.1f}% | {ns/n/1e3:.1f} |\n")
        f.write("\n## 15 longest launches\n\n| kernel | us | grid | block |\n|---|---:|---|---|\n")
        for k, ns, g, b in sorted(rows, key=lambda r: -r[1])[:15]:
            f.write(f"| `{k.split('(')[0]}` | {ns/1e3:.1f} | {g} | {b} |\n")
    print(open(dst).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_bytes.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__shared_mem_per_block_dynamic"]


def full(src, dst, traffic_json=None):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units, rows = rd[0], rd[1], rd[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows:
            d = dict(zip(hdr, r))
            f.write(f"## {d.get('Kernel Name','?')}  grid {d.get('Grid Size','')} block {d.get('Block Size','')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            stall = [(k, d[k]) for k in hdr if k.startswith("smsp__average_warps_issue_stalled") or k.startswith("smsp__average_warp_latency_issue_stalled")]
            for k, v in stall:
                f.write(f"| {k} | {v} | |\n")
            f.write("\n")
        if traffic_json and rows:
            d = dict(zip(hdr, rows[-1]))
            try:
                def tobytes(key):
                    v = float(d[key].replace(",", "")); u = units[hdr.index(key)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                t = tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")
                json.dump({"dram_bytes_per_launch": t, "kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "source": src}, open(traffic_json, "w"))
            except Exception as e:
                print("traffic json failed", e)
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
