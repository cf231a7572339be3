#!/usr/bin/env python
"""Benchmark of the M3DSSD dense forward path on B200 (BASELINE.json metric: images/sec @384x1280 bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the whole hot path over one batch of synthetic KITTI-shaped images on every
rank: DLA-34 + DLAUp/IDAUp (DCNv2) + shape/centre alignment + heads + softmax + decode/top-3000 +
(N > 1: NCCL all-gather of the detections) + batched NMS.  Workload = BASELINE.json configs[1]
(batch-8 384x1280 bf16 inference, DLA-34+DCNv2+align head) per GPU; weak scaling over GPUs.

`value`  : images/s with the input batches already resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the public API with HOST (pinned) inputs: H2D copy of every batch and
           D2H read of the kept detections inside the timed region.
`roofline`, `kernels`: per-kernel-family device time measured live (CUDA events around every launch of
           an instrumented pass) against MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the reference's PyTorch forward restated for CPU (oracle/ref_model.py,
           pinned to the unmodified reference modules by tests/golden) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CROP = (384, 1280)
LOCAL_BATCH = 8
WORKLOAD = "batch-8 384x1280 bf16 inference, DLA-34+DCNv2+align head per GPU (BASELINE.json configs[1])"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(FALLBACK_PEAKS, source="fallback")


class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm),
                "reasons": sorted(reasons)}


def cpu_reference(steps, warmup, want_detect=True):
    """The reference's forward on the host CPU: oracle/ref_model.py (CPU restatement pinned to the
    unmodified reference modules; DCNv2 = torchvision.ops.deform_conv2d, which the C oracle is pinned
    against, because the reference has no CPU DCNv2 at all).  One 384x1280 fp32 image per step."""
    import torch
    from m3dssd_b200 import synth
    from m3dssd_b200.model.M3d_inference_align import build
    from oracle import ref_model as RM
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    conf = synth.make_conf(attention=None, center_align=True, shape_align=True, crop_size=CROP)
    sd = synth.randomize_weights(build(conf, "test"))
    model = RM.RefModel(sd, conf, dcn="tv")
    x = synth.make_images(1, CROP)
    with torch.no_grad():
        for _ in range(warmup):
            out = model.forward(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            out = model.forward(x)
            if want_detect:
                model.detect(out, 0)
        dt = time.perf_counter() - t0
    return dict(value=steps / dt, unit="images/s", cores=cores, kind="port",
                sample="%d x (1 image 384x1280 fp32 forward + decode + NMS), torch CPU %d threads" % (steps, cores),
                ms_per_step=1e3 * dt / steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps, warmup = max(1, min(args.steps, 6)), max(1, min(args.warmup, 2))
    r = cpu_reference(steps, warmup)
    line = {
        "impl": "reference", "metric": "images_per_sec", "value": r["value"], "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU arm: bounded sample, 1 image per step, fp32"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def ncu_traffic(kind):
    """Measured DRAM bytes per launch of this kernel kind (dram__bytes_read.sum + dram__bytes_write.sum from the
    committed `ncu --set full` capture, profiles/r01c_traffic.json), or None when the kind was not captured."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01c_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(kind, {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def summarize_kernels(prof, peaks, step_ms):
    kinds = {}
    for p in prof:
        k = kinds.setdefault(p["kind"], dict(ms=0.0, flops=0.0, bytes=0.0, launches=0))
        k["ms"] += p["ms"]
        k["flops"] += p["flops"]
        k["bytes"] += p["bytes"]
        k["launches"] += p["launches"]
    total_ms = sum(k["ms"] for k in kinds.values())
    pt, ph = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    out = {}
    for name, k in kinds.items():
        t = k["ms"] * 1e-3
        tf = k["flops"] / t / 1e12 if t > 0 else 0.0
        gb = k["bytes"] / t / 1e9 if t > 0 else 0.0
        out[name] = dict(ms=round(k["ms"], 4), share=round(k["ms"] / total_ms, 4), launches=k["launches"],
                         tflops=round(tf, 2), hbm_gbs=round(gb, 1), tensor_frac=round(tf / pt, 4),
                         hbm_frac=round(gb / ph, 4))
    return out, total_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--attention", default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from m3dssd_b200 import synth
    from m3dssd_b200.model.M3d_inference_align import build
    from m3dssd_b200.parallel import ShardedDetector

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL writes its version banner to stdout when the communicator is created; stdout carries
        # exactly one JSON line, so point fd 1 at stderr until the first collective has run.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    W = max(3, args.warmup)
    K = args.steps
    peaks = load_peaks()

    def note(msg):
        if os.environ.get("M3D_BENCH_VERBOSE"):
            print("[bench rank %d] %s" % (rank, msg), file=sys.stderr, flush=True)

    conf = synth.make_conf(attention=args.attention, center_align=True, shape_align=True, crop_size=CROP,
                           batch_size=LOCAL_BATCH)
    conf.precision = "bf16"
    net = build(conf, "test")
    synth.randomize_weights(net)
    net = net.cuda()
    note("weights ready")
    det = ShardedDetector(net, LOCAL_BATCH, CROP[0], CROP[1], precision="bf16", use_graph=True)
    note("engine built")
    eng = det.engine

    NB = 4  # distinct input batches, rotated; activations per step (>1 GB) exceed the 126 MB L2 by themselves
    host = [synth.make_images(LOCAL_BATCH, CROP, seed=100 * rank + i).pin_memory() for i in range(NB)]
    dev = [h.cuda() for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------- device-resident throughput
    def drain():
        torch.cuda.current_stream().wait_event(eng.tail_done)  # the last tail runs on the side stream

    for i in range(W):
        det.step_pipelined(dev[i % NB])
    drain()
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        det.step_pipelined(dev[i % NB])
    drain()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    note("device-resident loop done: %.3f ms/step" % (ms / K))
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * LOCAL_BATCH * K / (ms_total * 1e-3)

    # ---------------------------------------------------------- end to end from host memory
    kept_host = torch.empty(det.kept.shape if world > 1 else eng.kept.shape, dtype=torch.float32).pin_memory()
    num_host = torch.empty((world * LOCAL_BATCH,), dtype=torch.int32).pin_memory()
    copy_stream = torch.cuda.Stream()
    stage_bufs = [torch.empty_like(dev[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        # two-deep pipeline: the H2D copy of batch i+1 overlaps the compute of batch i
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(copy_stream):
            stage_bufs[0].copy_(host[0], non_blocking=True)
            ready[0].record(copy_stream)
        for i in range(n):
            s = i & 1
            if i + 1 < n:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(consumed[(i + 1) & 1])
                    stage_bufs[(i + 1) & 1].copy_(host[(i + 1) % NB], non_blocking=True)
                    ready[(i + 1) & 1].record(copy_stream)
            cur.wait_event(ready[s])
            kept, num = det.step_pipelined(stage_bufs[s])
            consumed[s].record(cur)  # (recorded after the heads; the input buffer is only read by the stem)
            with torch.cuda.stream(eng.tail_stream):  # D2H of the kept detections right behind their NMS
                kept_host.copy_(kept, non_blocking=True)
                num_host.copy_(num, non_blocking=True)
                eng.tail_done.record(eng.tail_stream)
        drain()
        cur.synchronize()

    e2e_loop(W)
    barrier()
    note("e2e warm-up done")
    t0 = time.perf_counter()
    e0.record()
    e2e_loop(K)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms_e2e = max(e0.elapsed_time(e1), 0.0)
    t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * LOCAL_BATCH * K / (float(t.item()) * 1e-3)
    h2d = LOCAL_BATCH * 3 * CROP[0] * CROP[1] * 4
    d2h = kept_host.numel() * 4 + num_host.numel() * 4

    # ---------------------------------------------------------- per-kernel roofline (rank 0)
    line = None
    if rank == 0:
        prof = eng.profile(iters=3)
        kinds, fwd_ms = summarize_kernels(prof, peaks, ms_total / K)
        dom = max(kinds.items(), key=lambda kv: kv[1]["ms"])
        dname, d = dom
        tensor_bound = d["tensor_frac"] >= d["hbm_frac"]
        roofline = {
            "kernel": dname, "bound": "tensor" if tensor_bound else "hbm",
            "achieved": d["tflops"] if tensor_bound else d["hbm_gbs"],
            "peak": peaks["bf16_tflops_sustained"] if tensor_bound else peaks["hbm_gbs"],
            "unit": "TFLOP/s" if tensor_bound else "GB/s",
            "frac": d["tensor_frac"] if tensor_bound else d["hbm_frac"],
            "traffic": ncu_traffic(dname), "peak_source": peaks["source"] + " (sustained: kernel timed inside a long step)",
            "share_of_step": d["share"], "launches_per_step": d["launches"],
        }
        tot_fl = sum(p["flops"] for p in prof)
        tot_by = sum(p["bytes"] for p in prof)
        step_s = ms_total / K * 1e-3
        step_roof = {
            "gflop_per_image": round(tot_fl / LOCAL_BATCH / 1e9, 2), "gb_per_image": round(tot_by / LOCAL_BATCH / 1e9, 4),
            "tensor_frac": round(tot_fl / step_s / 1e12 / peaks["bf16_tflops_sustained"], 4),
            "hbm_frac": round(tot_by / step_s / 1e9 / peaks["hbm_gbs"], 4),
            "forward_ms_eager_sum": round(fwd_ms, 3),
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_reference(steps=4, warmup=1)
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": "images_per_sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": world * LOCAL_BATCH, "image": "384x1280",
                       "backbone": "dla34", "align": True, "attention": args.attention,
                       "parallelism": "dp%d (images sharded; all-gather of detections before NMS)" % world,
                       "l2": "4 rotating input batches; per-step activation footprint > 1 GB >> 126 MB L2",
                       "cuda_graph": True,
                       "pipelining": "detection tail of batch i (decode, top-K, NMS) runs on a side stream under the "
                                     "trunk of batch i+1; every step's work is inside the timed region"},
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "wall_s": wall, "pipeline": "H2D of batch i+1 and the detection tail + D2H of batch i overlap the trunk of batch i+1"},
            "gpu_launches": det.launches_per_step * K,
            "roofline": roofline, "kernels": kinds, "step_roofline": step_roof,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
    if line is not None:
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()
    if world > 1:
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception as e:  # noqa: BLE001  (the measurement is already printed)
            print("[bench] process-group teardown: %r" % (e,), file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
