"""Turn raw ncu exports into the markdown summaries committed under profiles/.

    python tools/summarize_profiles.py launches <launch-list.csv> <steps-in-run> > profiles/launches_rN_summary.md
    python tools/summarize_profiles.py report   <raw-page.csv>                   > profiles/ncu_gemm_rN.md

`launch-list.csv` is the log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`;
`raw-page.csv` is `ncu -i <rep> --page raw --csv` of an `ncu --set full` capture.
"""
import collections
import csv
import re
import sys

REPORT_METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__cluster_dim_x", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum", "sm__cycles_elapsed.max",
    "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
]


def short_name(full):
    name = re.sub(r"\(.*", "", full)
    name = name.replace("void ", "")
    return name.split("::")[-1] if "at::" not in name else name[:70]


def launches(path, steps):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        if r[ui] in ("ns", "nsecond"):
            v /= 1e6
        elif r[ui] in ("us", "usecond"):
            v /= 1e3
        a = agg[short_name(r[ki])]
        a[0] += 1
        a[1] += v
    total = sum(v[1] for v in agg.values())
    print(f"Sum of kernel durations: {total / steps:.1f} ms/step, {sum(v[0] for v in agg.values()) / steps:.0f} "
          f"launches/step (averaged over the {steps} steps of the run, bench set-up kernels included).\n")
    print("| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / total < 2e-5:
            continue
        print(f"| `{k}` | {v[0] / steps:.1f} | {v[1] / steps:.2f} | {100 * v[1] / total:.1f} % |")


def report(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " | unit |")
    print("|---|" + "---|" * (len(data) + 1))
    ki = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join("`" + short_name(r[ki]) + "`" for r in data) + " | |")
    for m in REPORT_METRICS:
        if m not in hdr:
            continue
        i = hdr.index(m)
        print(f"| {m} | " + " | ".join(r[i][:12] for r in data) + f" | {units[i]} |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        report(sys.argv[2])
