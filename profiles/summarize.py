"""Summarise ncu outputs into small tracked text files.
  python profiles/summarize.py launches <launches.csv>            -> per-kernel totals / shares
  python profiles/summarize.py raw <report.ncu-rep> [regex]       -> key metrics of the first launch of each kernel
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for row in r:
        if len(row) <= vi:
            continue
        name = re.sub(r"\(.*", "", row[ki])
        name = re.sub(r"void |gvx::", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(row[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"{t / 1e3:10.1f} us  {c:5d} launches  {t / c / 1e3:8.2f} us avg  {100 * t / tot:5.1f}%  {k[:100]}")


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__issue_active.avg.per_cycle_active"]


def raw(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki, gi = hdr.index("Kernel Name"), hdr.index("launch__grid_size")
    seen = set()
    for r in data:
        name = r[ki]
        if pattern and not re.search(pattern, name):
            continue
        key = name[:80] + "|" + r[gi]
        if key in seen:
            continue
        seen.add(key)
        print("----", name[:110])
        for w, i in idx:
            print(f"   {w:86s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
