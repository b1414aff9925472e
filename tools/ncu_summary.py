"""Turns gpurun_out/*.ncu-rep and launch-list CSVs into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/r5_launches.csv > profiles/r1_launches.md
  python tools/ncu_summary.py kernel   gpurun_out/r5_conv_tc.ncu-rep > profiles/r1_conv_tc_decode1.md
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r".*::", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3
        n += 1
    tot = sum(v[1] for v in agg.values())
    print(f"# launch list ({path}): {n} launches, {tot / 1e3:.2f} ms of kernel time (ncu: cold cache, serialised -- compare shares)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg ",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum ", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread ", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.avg.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg ", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_active.avg "]


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print(f"# {d.get('Kernel Name', '?')[:120]}\n")
        print("| metric | unit | value |\n|---|---|---:|")
        for h, u, v in zip(hdr, units, vals):
            if any(h == w.strip() or (w.endswith(" ") and h == w.strip()) for w in WANT) or h in [w.strip() for w in WANT]:
                print(f"| {h} | {u} | {v} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
