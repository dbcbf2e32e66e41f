"""Condenses ncu output into the small text files committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv [skip_first_n] > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [label ...]       > profiles/rNN_ncu_full.txt

`launches`: the `--metrics gpu__time_duration.sum` CSV of one bench.py run, grouped by kernel
(count, total and share of the listed launches).  `full`: selected metrics of every kernel in an
`ncu --set full` report (labels name the profiled launches in order).
"""
import collections
import csv
import io
import subprocess
import sys

FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "memory throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX(smem) throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("smsp__cycles_active.avg", "SMSP active cycles"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "HMMA (mma.sync) pipe %"),
]


def launches(path, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    if "Metric Name" in hdr:   # several metrics per launch: keep the duration rows only
        mn = hdr.index("Metric Name")
        rows = [r for r in rows if r[mn] == "gpu__time_duration.sum"]
    rows = rows[skip:]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[k].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[v].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {len(rows)} launches, {tot / 1e6:.3f} ms total (ncu per-launch times: serialised, cold cache)")
    print(f"{'kernel':60s} {'count':>6s} {'total ms':>10s} {'share':>7s}")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:60]:60s} {c:6d} {t / 1e6:10.3f} {100 * t / tot:6.1f}%")


def full(path, labels):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for n, d in enumerate(data):
        lab = labels[n] if n < len(labels) else ""
        print(f"## launch {n}: {lab}  [{d[idx['Kernel Name']].split('(')[0]}]")
        for m, nice in FULL_METRICS:
            if m in idx:
                print(f"  {nice:28s} {d[idx[m]]:>16s} {units[idx[m]]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        full(sys.argv[2], sys.argv[3:])
