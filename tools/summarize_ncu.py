"""Summarise Nsight Compute CSV exports into the small markdown tables committed under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01_launches.csv  > profiles/r01_launch_summary.md
    python tools/summarize_ncu.py hot      gpurun_out/r01_hot_raw.csv   > profiles/r01_hot_kernels.md
    python tools/summarize_ncu.py traffic  gpurun_out/r02_step_metrics.csv > profiles/r02_traffic.json
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


def kernel_key(name):
    """ncu's demangled kernel name -> the name bench.py / m3d_last_kernel use, e.g.
    'void unnamed>::conv_halo_kernel<64, __nv_bfloat16, 1, 1>(ConvTmaParams)' -> 'conv_halo_kernel<64,bf16,1,1>'."""
    k = re.sub(r"^void\s+", "", name)
    k = re.sub(r"\((int|bool|unsigned int)\)", "", k)
    k = re.sub(r"\(.*$", "", k)                      # parameter list
    head, lt, tail = k.partition("<")
    if head == "" or head.endswith("::"):            # '<unnamed>::kernel<...' : the first '<' belongs to the namespace
        k = k[k.index(">::") + 3:] if ">::" in k else k
        head, lt, tail = k.partition("<")
    k = head.split("::")[-1] + lt + tail
    k = k.replace("__nv_bfloat16", "bf16").replace("float", "f32").replace("true", "1").replace("false", "0")
    return k.replace(" ", "")


def traffic(path):
    """Per kernel FUNCTION, over ALL its launches in one captured step: mean DRAM bytes per launch
    (dram__bytes_read.sum + dram__bytes_write.sum), mean duration, mean tensor-pipe activity.  JSON for bench.py."""
    import json
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi, mi, ii, ui = (hdr.index(c) for c in ("Kernel Name", "Metric Value", "Metric Name", "ID", "Metric Unit"))
    per = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[ui].lower()
        if r[mi].startswith("dram__bytes"):
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3}.get(unit, 1)
        per.setdefault((int(r[ii]), kernel_key(r[ki])), {})[r[mi]] = v
    agg = OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k, dict(launches=0, dram=0.0, rd=0.0, wr=0.0, us=0.0, tensor=0.0))
        a["launches"] += 1
        a["rd"] += m.get("dram__bytes_read.sum", 0.0)
        a["wr"] += m.get("dram__bytes_write.sum", 0.0)
        a["us"] += m.get("gpu__time_duration.sum", 0.0)
        a["tensor"] += m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    out = OrderedDict()
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["launches"]
        out[k] = dict(launches_captured=n, dram_bytes_per_launch=round((a["rd"] + a["wr"]) / n),
                      dram_read_bytes_per_launch=round(a["rd"] / n), dram_write_bytes_per_launch=round(a["wr"] / n),
                      ncu_us_per_launch=round(a["us"] / n, 2), ncu_tensor_pipe_active_pct=round(a["tensor"] / n, 2),
                      ncu_dram_gbs=round((a["rd"] + a["wr"]) / max(a["us"], 1e-9) / 1e3, 1),
                      source="%s: ncu --clock-control none, every launch of one warm step (cold-cache, serialised: shares "
                             "and bytes are comparable, absolute times are not)" % path.split("/")[-1])
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    {"launches": launches, "hot": hot, "traffic": traffic}[sys.argv[1]](sys.argv[2])
