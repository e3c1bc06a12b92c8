"""Turn ncu outputs (brought back in gpurun_out/) into the small summaries committed here.
usage: python profiles/summarize.py launches <launch-list.csv> | full <report.ncu-rep>"""
import collections
import csv
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Name"] == "gpu__time_duration.sum":
            u = row["Metric Unit"]
            v = v if u == "ns" else v * 1e3 if u == "us" else v * 1e6
        agg[row["Kernel Name"].split("(")[0]][row["Metric Name"]].append(v)
    tot = sum(sum(d["gpu__time_duration.sum"]) / len(d["gpu__time_duration.sum"]) for d in agg.values())
    print("| kernel | launches | avg us | share of substep |")
    print("|---|---|---|---|")
    for k, d in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"]) / len(kv[1]["gpu__time_duration.sum"])):
        t = d["gpu__time_duration.sum"]
        avg = sum(t) / len(t)
        print(f"| {k} | {len(t)} | {avg / 1e3:.1f} | {avg / tot * 100:.1f}% |")
    print(f"\nsum of per-substep kernel time: {tot / 1e3:.1f} us (cold-cache, serialised under ncu)")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    idx = [hdr.index(w) for w in WANT if w in hdr]
    print("kernel," + ",".join(f"{hdr[i]} [{units[i]}]" for i in idx))
    for r in rows[2:]:
        name = r[ik].split("(")[0].replace(", ", ";")     # template arguments: keep the row's column count
        print(name + "," + ",".join(r[i].replace(",", "") for i in idx))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
