"""Summarise Nsight Compute CSV exports into the small markdown tables committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01_launches.csv  > profiles/r01_launch_summary.md
    python tools/summarize_ncu.py hot      gpurun_out/r01_hot_raw.csv   > profiles/r01_hot_kernels.md
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("m3d::", "")
    return re.sub(r"\(.*$", "", name)[:70]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ii = hdr.index("ID")
    seq = [(int(r[ii]), short(r[ki]), float(r[vi].replace(",", ""))) for r in rows[1:] if len(r) > vi and r[mi] == "gpu__time_duration.sum"]
    # one step = everything from the last stem launch (first kernel of the plan) onwards
    starts = [i for i, (_, k, _) in enumerate(seq)
              if "stem_s2d_kernel" in k or k.startswith("conv_gather_kernel<64,") or k.startswith("stem_conv7x7")]
    step = seq[starts[-1]:] if starts else seq
    agg = OrderedDict()
    for _, k, v in step:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    unit = "ns"
    print("# ncu launch list, one warm step (batch 8, 384x1280, bf16, eager replay; `--metrics gpu__time_duration.sum --clock-control none`)\n")
    print("Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes.\n")
    print("%d launches in the step, %.3f ms summed.\n" % (len(step), tot / 1e6))
    print("| kernel | launches | sum (us) | share |\n|---|---:|---:|---:|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, v / 1e3, 100 * v / tot))


def hot(path):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], rows[2:]
    cols = OrderedDict([
        ("gpu__time_duration.sum", "dur us"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd MB"), ("dram__bytes_write.sum", "dram wr MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %")])
    print("# ncu --set full, selected launches of one step (batch 8, 384x1280, bf16)\n")
    print("`traffic` for bench.py's roofline = dram rd + dram wr of the kernel below.\n")
    print("| kernel | " + " | ".join(cols.values()) + " | top stalls |\n|---|" + "---:|" * len(cols) + "---|")
    for r in data:
        name = short(r[hdr.index("Kernel Name")])
        vals = []
        for c in cols:
            v = r[hdr.index(c)] if c in hdr else ""
            try:
                vals.append("%.1f" % float(v))
            except ValueError:
                vals.append(v)
        st = [(float(r[i] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
        tot = sum(v for v, _ in st) or 1
        top = ", ".join("%s %.0f%%" % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:3])
        print("| `%s` | " % name + " | ".join(vals) + " | %s |" % top)


if __name__ == "__main__":
    {"launches": launches, "hot": hot}[sys.argv[1]](sys.argv[2])
