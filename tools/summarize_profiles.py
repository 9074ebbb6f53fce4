#!/usr/bin/env python
"""Turn raw gpurun artefacts (ncu launch list CSV, ncu --set full report) into the small text/JSON
summaries committed under profiles/.  Usage: tools/summarize_profiles.py <round-tag>"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
go = os.path.join(ROOT, "gpurun_out")


def short(name):
    name = re.sub(r"fsm::", "", name)
    name = re.sub(r"\(int\)", "", name)
    return re.sub(r"\(.*", "", name)[:100]


# ---- launch list ---------------------------------------------------------------------------
path = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        key = short(row["Kernel Name"]) + f" grid={row['Grid Size']} block={row['Block Size']}"
        agg.setdefault(key, []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(out_dir, f"{tag}_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 200 ... python bench.py --steps 2 --warmup 3\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# {'total_us':>10} {'share':>6} {'n':>4} {'avg_us':>9}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"  {sum(v):10.1f} {100 * sum(v) / total:5.1f}% {len(v):4d} {sum(v) / len(v):9.1f}  {k}\n")
    print("wrote launches summary")

# ---- full ncu report ------------------------------------------------------------------------
def summarize_full(stem, header, traffic_workload, chunk):
    rep = os.path.join(go, f"prof_{tag}_{stem}.ncu-rep")
    rawcsv = os.path.join(go, f"prof_{tag}_{stem}_raw.csv")
    if os.path.exists(rawcsv):
        raw = open(rawcsv).read()
    elif os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "lts__t_sector_hit_rate.pct",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
    traffic = {}
    with open(os.path.join(out_dir, f"{tag}_ncu_{stem}_summary.txt"), "w") as f:
        f.write(header)
        for d in data:
            name = short(d[hdr.index("Kernel Name")])
            f.write(f"\n== {name}\n")
            for w in want:
                if w in hdr:
                    f.write(f"   {w:88s} {d[hdr.index(w)]:>18s} {units[hdr.index(w)]}\n")

            def val(m):
                v = float(d[hdr.index(m)].replace(",", ""))
                u = units[hdr.index(m)]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
            base = re.sub(r"<.*", "", name).replace("void ", "")
            traffic[base] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                             "kernel": name}
    json.dump({"source": f"profiles/{tag}_ncu_{stem}_summary.txt", "workload": traffic_workload, "chunk": chunk, "kernels": traffic},
              open(os.path.join(out_dir, f"{tag}_ncu_traffic.json" if stem == "step" else f"{tag}_ncu_{stem}_traffic.json"), "w"), indent=1)
    print("wrote ncu summary", {k: round(v["dram_bytes_per_launch"] / 1e6) for k, v in traffic.items()})



summarize_full("step", "# ncu --set full --clock-control none --import-source on --profile-from-start off -c 3 python tools/profile_c3.py\n"
               "# one ETDRK2 step of C3 (1024^2 x 64, library defaults: 2 lanes x 32 samples per launch): first evaluation's IX, PHYS, FX\n",
               "C3 1024^2 x 64, 2 lanes x 32 samples per launch (library default)", 32)
summarize_full("c5", "# FSM_NCU_WINDOW=1 ncu --set full --clock-control none --import-source on --profile-from-start off -c 20 python tools/bench_configs.py c5\n"
               "# one SETDRK4 step of C5 (512^3, one GPU): IX, MID inverse, PHYS, MID forward, FX of each of the four evaluations\n",
               "C5 512^3 x 1", 1)

for name in (f"bench_{tag}.json", f"bench_ref_{tag}.json", f"configs_{tag}.jsonl"):
    src = os.path.join(go, name)
    if os.path.exists(src):
        open(os.path.join(out_dir, f"{tag}_{name.replace('_' + tag, '')}"), "w").write(open(src).read())
        print("copied", name)
