#!/usr/bin/env python
"""bench.py — RepSurf-U fwd+bwd(+SGD step) throughput on synthetic clouds, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload seg|cls] [--impl reference]

Prints ONE JSON line (rank 0).  Contract: see the task statement / DESIGN.md "Measurement".
  value         clouds/sec, whole job, inputs already resident in HBM
  e2e           same metric through the public module API with HOST (pinned) inputs: H2D of the step's
                inputs and D2H of the loss inside the timed region
  roofline      the dominant repsurf_b200 kernel, timed with CUDA events inside the timed region
  cpu_baseline  the oracle port (oracle/model_ref.py + oracle C) on the host cores, bounded sample
  --impl reference   times that CPU port alone (the reference has no CPU path for segmentation and
                     /root/reference does not exist on the GPU box)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.nn as nn

# -------------------------------------------------------------------------------------------- workloads
WORKLOADS = {
    # BASELINE.json configs[2]/[3]: RepSurf-U S3DIS seg, B=8 clouds x N=40960 per GPU, training mode
    "seg": dict(name="RepSurf-U S3DIS seg (repsurf_umb_ssg), B=8 x N=40960 per GPU, fwd+bwd+SGD", clouds=8, n=40960),
    # BASELINE.json configs[1]: RepSurf-U ScanObjectNN cls, B=32, N=1024
    "cls": dict(name="RepSurf-U ScanObjectNN cls (repsurf_ssg_umb), B=32 x N=1024 per GPU, fwd+bwd+SGD", clouds=32, n=1024),
}


def make_inputs(workload, clouds, n, seed, pin):
    """Synthetic data of SURVEY.md §8(d): seg coord=rand*[8,8,3] mean-centred per cloud, feat=randn, 13 classes;
    cls points=rand*2-1, 15 classes.  Host tensors (pinned when asked)."""
    g = torch.Generator().manual_seed(seed)
    if workload == "seg":
        coord = torch.rand(clouds * n, 3, generator=g) * torch.tensor([8.0, 8.0, 3.0])
        coord = (coord.view(clouds, n, 3) - coord.view(clouds, n, 3).mean(1, keepdim=True)).reshape(-1, 3).contiguous()
        feat = torch.randn(clouds * n, 3, generator=g)
        target = torch.randint(0, 13, (clouds * n,), generator=g)
        offset = (torch.arange(1, clouds + 1) * n).int()
        ts = [coord, feat, offset, target]
    else:
        pts = torch.rand(clouds, 3, n, generator=g) * 2 - 1
        target = torch.randint(0, 15, (clouds,), generator=g)
        ts = [pts, target]
    if pin:
        ts = [t.pin_memory() for t in ts]
    return ts


# -------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  Uses NVML in-process
    (nvidia_ml_py) from a background thread: an external `nvidia-smi -lms` loop measurably slows the step."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.thread, self.ok = gpu_index, [], False, None, False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index follows CUDA_VISIBLE_DEVICES; map through the PCI bus id
            bus = torch.cuda.get_device_properties(self.idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.idx), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                    if int(pynvml.nvmlDeviceGetPciInfo(hi).bus) == int(bus):
                        h = hi
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.nv, self.h, self.ok = pynvml, h, True
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.ok = False
            return

        def loop():
            nv = self.nv
            while not self.stop_flag:
                try:
                    sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                        else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((float(sm), int(rs)))
                except Exception:
                    pass
                time.sleep(0.4)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def stop(self):
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.thread.join(timeout=1.0)
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted({k for _sm, rs in self.rows for k, bit in names.items() if rs & bit})
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm)}


# -------------------------------------------------------------------------------------------- per-entry timing
class _OpndInfo:
    def __init__(self, o):
        self.K, self.ku, self.kind = int(o.K), int(o.ku), int(o.kind)


def _lite(x):
    if isinstance(x, (int, float)) or x is None:
        return x
    if hasattr(x, "kind") and hasattr(x, "ku"):
        return _OpndInfo(x)
    return None


class EntryTimer:
    """CUDA-event timing of every C-ABI call made through repsurf_b200._native.call (torch's current stream)."""

    def __init__(self, native, only=None):
        self.native, self.orig, self.ev, self.only = native, native.call, [], only

    def __enter__(self):
        def timed(name, *args):
            if self.only is not None and name not in self.only:
                return self.orig(name, *args)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.orig(name, *args)
            b.record()
            # keep only plain numbers: holding tensors / descriptors here would pin every activation of the step
            self.ev.append((name, tuple(_lite(x) for x in args[:8]), a, b))
        self.native.call = timed
        return self

    def __exit__(self, *exc):
        self.native.call = self.orig
        return False

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, args, a, b in self.ev:
            d = out.setdefault(name, {"ms": 0.0, "calls": 0, "each": []})
            t = a.elapsed_time(b)
            d["ms"] += t
            d["calls"] += 1
            d["each"].append((args, t))
        return out


# -------------------------------------------------------------------------------------------- algorithmic work
def fps_algorithmic_bytes(n, m):
    """SURVEY.md §8(d) streaming model: (m-1) * n * 20 B (12 B xyz + 4 B read + 4 B write of the running
    minimum) + 4 m B of indices, per segment."""
    return (m - 1) * n * 20 + 4 * m


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# -------------------------------------------------------------------------------------------- CPU arm
def cpu_port_run(workload, steps, warmup, sample_clouds, sample_n):
    """fwd+bwd of the oracle port on the host cores over a bounded sample; returns (clouds/s, seconds/step, cores)."""
    from oracle import model_ref as MR
    from oracle import oracle as O
    torch.set_num_threads(os.cpu_count())
    cores = max(torch.get_num_threads(), O.num_threads())
    torch.manual_seed(0)
    np.random.seed(0)
    if workload == "seg":
        model = MR.SegNet().train()
        coord, feat, offset, target = make_inputs("seg", sample_clouds, sample_n, 0, False)
        crit = nn.CrossEntropyLoss()
        def step():
            model.zero_grad(set_to_none=True)
            crit(model([coord, feat, offset]), target).backward()
    else:
        from repsurf_b200.models import SmoothClsLoss
        model = MR.ClsNet().train()
        pts, target = make_inputs("cls", sample_clouds, sample_n, 0, False)
        crit = SmoothClsLoss()
        def step():
            model.zero_grad(set_to_none=True)
            crit(model(pts), target).backward()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return sample_clouds / dt, dt, cores


# -------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("RSB_WORKLOAD", "seg"), choices=list(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    wl = WORKLOADS[args.workload]
    metric = "clouds/sec fwd+bwd RepSurf-U"
    sample = dict(seg=(1, 40960), cls=(8, 1024))[args.workload]

    # ---------------- reference arm: the CPU port on the host cores (rank 0 only) ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        steps = max(1, min(args.steps, 3))
        warm = 1 if args.warmup > 0 else 0
        v, dt, cores = cpu_port_run(args.workload, steps, warm, *sample)
        desc = f"{sample[0]} cloud(s) x N={sample[1]} of the same workload, fwd+bwd, {steps} step(s)"
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": v, "unit": "clouds/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": wl["name"], "sample": desc},
            "cpu_baseline": {"value": v, "unit": "clouds/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": v, "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- our arm ---------------------------------------------------------------------
    import torch.distributed as dist
    from repsurf_b200 import _native
    from repsurf_b200.models import RepSurfCls, RepSurfSeg, SmoothClsLoss
    from repsurf_b200.seg import pointops as PS

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = False          # fp32-faithful MLP (north star: 1e-5 rel)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(rank)
    np.random.seed(rank)

    model = (RepSurfSeg() if args.workload == "seg" else RepSurfCls()).to(dev).train()
    crit = nn.CrossEntropyLoss() if args.workload == "seg" else SmoothClsLoss()
    from repsurf_b200.dist import FlatGrads, broadcast_module
    params = [p for p in model.parameters()]
    broadcast_module(model)
    # gradients are packed into one flat buffer after backward: ONE all-reduce per step (3.9 MB seg / 5.9 MB cls)
    fg = FlatGrads(params)
    opt = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=1e-4)

    host = make_inputs(args.workload, wl["clouds"], wl["n"], 100 + rank, pin=True)
    devin = [t.to(dev) for t in host]
    if args.workload == "seg":
        PS.register_offsets(devin[2], host[2].tolist())
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    def fwd_bwd(inp):
        fg.zero()
        if args.workload == "seg":
            loss = crit(model([inp[0], inp[1], inp[2]]), inp[3])
        else:
            loss = crit(model(inp[0]), inp[1])
        loss.backward()
        fg.allreduce_mean()
        opt.step()
        return loss

    def step_resident():
        return fwd_bwd(devin)

    def step_e2e():
        inp = [t.to(dev, non_blocking=True) for t in host]
        if args.workload == "seg":
            PS.register_offsets(inp[2], host[2].tolist())   # the host already holds the offsets it uploads
        loss = fwd_bwd(inp)
        return float(loss)                                   # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("RSB_NO_CLOCKS"):
        clocks.start()
    _native.reset_launch_count()
    timed_entries = None if os.environ.get("RSB_TIME_ALL_ENTRIES") else {
        "rsb_furthestsampling_packed", "rsb_furthestsampling_dense", "rsb_knnquery_packed", "rsb_knnquery_dense",
        "rsb_knnquery_grid", "rsb_gemm_wgrad", "rsb_gemm_rows", "rsb_ballquery"}
    with EntryTimer(_native, timed_entries) as et:
        ms_step = timed(step_resident, args.steps)
    launches = _native.launch_count()
    per_entry = et.summary()
    clk = clocks.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    clouds_total = wl["clouds"] * world
    value = clouds_total / (ms_step * 1e-3)
    e2e_value = clouds_total / (ms_e2e * 1e-3)

    # ---- rooflines: every timed C-ABI entry against its own bound; "roofline" = the entry with the most time ----
    hbm_peak, tf_peak, peak_src = peaks()
    entry_share = {k: round(v["ms"] / (ms_step * args.steps), 4) for k, v in per_entry.items()}
    dom = max(per_entry.items(), key=lambda kv: kv[1]["ms"]) if per_entry else None

    def biggest(v):
        big = max(t for _a, t in v["each"])
        sel = [(a, t) for a, t in v["each"] if t > 0.6 * big]
        return sel[0][0], float(np.mean([t for _a, t in sel]))

    def opnd_bytes(o, rows):
        per_row = {0: o.K, 1: o.K, 2: 2 * o.K, 3: o.K + min(o.K, o.ku), 4: o.K}[o.kind]
        return rows * per_row * 4

    def most_bytes(v, alg_of):
        """GEMM entries run at many shapes per step: report the launch shape that moves the most algorithmic bytes
        (the one the ncu traffic capture in profiles/ holds), averaged over its launches."""
        best = max(v["each"], key=lambda at: alg_of(at[0]))[0]
        key = alg_of(best)
        ts = [t for a_, t in v["each"] if alg_of(a_) == key and a_[0] == best[0]]
        return best, float(np.mean(ts))

    rooflines = {}
    for name, v in per_entry.items():
        a, t_ms = biggest(v)
        if name == "rsb_gemm_wgrad":
            a, t_ms = most_bytes(v, lambda x: opnd_bytes(x[1], x[0]) + opnd_bytes(x[2], x[0]))
        elif name == "rsb_gemm_rows":
            a, t_ms = most_bytes(v, lambda x: opnd_bytes(x[2], x[0]) + x[0] * x[1] * 4)
        if name.startswith("rsb_furthestsampling"):
            if name.endswith("packed"):
                nseg, n_max = a[0], a[1]
                m_seg = (wl["n"] // 4 // 4) if (args.workload == "seg" and nseg == wl["clouds"] * 4) else n_max // 4
                alg = nseg * fps_algorithmic_bytes(n_max if nseg != wl["clouds"] * 4 else wl["n"] / 4, m_seg)
                note = f"{nseg} segments, largest {n_max} points, ~{m_seg} samples each (streaming model, SURVEY 8d)"
            else:
                alg = a[0] * fps_algorithmic_bytes(a[1], a[2])
                note = f"{a[0]} clouds {a[1]} -> {a[2]} (streaming model, SURVEY 8d)"
            ach = alg / (t_ms * 1e-3) / 1e9
            rooflines[name] = {"kernel": "fps_kernel", "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                               "frac": ach / hbm_peak, "launch_ms": t_ms, "launch": note, "algorithmic_bytes_per_launch": alg}
        elif name == "rsb_gemm_wgrad":
            rows_, G_, X_ = a[0], a[1], a[2]
            alg = opnd_bytes(G_, rows_) + opnd_bytes(X_, rows_) + G_.K * X_.K * 4
            flops = 2.0 * rows_ * G_.K * X_.K
            ach = alg / (t_ms * 1e-3) / 1e9
            rooflines[name] = {"kernel": "gemm_wgrad_kernel (tcgen05 3xTF32)", "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                               "unit": "GB/s", "frac": ach / hbm_peak, "launch_ms": t_ms,
                               "launch": f"rows={rows_} dW[{G_.K}x{X_.K}], operand kinds {G_.kind}/{X_.kind}",
                               "algorithmic_bytes_per_launch": alg, "algorithmic_tflops": flops / (t_ms * 1e-3) / 1e12}
        elif name == "rsb_gemm_rows":
            rows_, N_, A_ = a[0], a[1], a[2]
            alg = opnd_bytes(A_, rows_) + rows_ * N_ * 4
            ach = alg / (t_ms * 1e-3) / 1e9
            rooflines[name] = {"kernel": "gemm_rows_kernel (tcgen05 3xTF32)", "bound": "hbm", "achieved": ach, "peak": hbm_peak,
                               "unit": "GB/s", "frac": ach / hbm_peak, "launch_ms": t_ms,
                               "launch": f"rows={rows_} K={A_.K} N={N_}, operand kind {A_.kind}",
                               "algorithmic_bytes_per_launch": alg,
                               "algorithmic_tflops": 2.0 * rows_ * N_ * A_.K / (t_ms * 1e-3) / 1e12}
        elif name.startswith("rsb_knnquery"):
            pairs = None
            if name == "rsb_knnquery_grid":
                n_tot, m_tot, b_ = a[5], a[6], a[2]
                pairs = m_tot * (n_tot / max(b_, 1))
                note = f"m={m_tot} queries x n={n_tot // max(b_, 1)} candidates/cloud, k={v['each'][0][0][7] if len(v['each'][0][0]) > 7 else '?'} (uniform-grid search)"
            elif name == "rsb_knnquery_packed":
                note = "all-pairs kernel (small clouds)"
            if pairs:
                fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12     # B200: 148 SMs x 128 FFMA lanes x 2 x 1.965 GHz
                ach = pairs * 9 / (t_ms * 1e-3) / 1e12
                rooflines[name] = {"kernel": "knn_grid_kernel", "bound": "fp32 (all-pairs model: 9 flop/pair)", "achieved": ach,
                                   "peak": fp32_peak, "unit": "TFLOP/s", "frac": ach / fp32_peak, "launch_ms": t_ms, "launch": note,
                                   "algorithmic_pairs_per_launch": pairs}
    # measured DRAM traffic per launch from the committed ncu --set full capture of the same kernels
    # (profiles/r01_ncu_full_seg_v2_traffic.json, produced by scripts/ncu_summary.py; seg workload shapes)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_full_seg_v2_traffic.json")
    if os.path.exists(tpath) and args.workload == "seg":
        traffic = json.load(open(tpath))
    fam = {"rsb_gemm_wgrad": "gemm_wgrad_kernel", "rsb_gemm_rows": "gemm_rows_kernel", "rsb_knnquery_grid": "knn_grid_kernel",
           "rsb_furthestsampling_packed": "fps_kernel", "rsb_furthestsampling_dense": "fps_kernel"}
    for name, r in rooflines.items():
        t = traffic.get(fam.get(name, ""))
        r["traffic"] = t["dram_bytes"] if t else None
        r["peak_source"] = peak_src
    roof = dict(rooflines[dom[0]]) if dom and dom[0] in rooflines else None

    out = {
        "metric": metric, "value": value, "unit": "clouds/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl["name"], "clouds_per_gpu": wl["clouds"], "points_per_cloud": wl["n"], "parallelism": f"dp{world}",
                   "optimizer_step": "SGD momentum inside the timed region", "tf32": False,
                   "l2": "per-step working set (activations > 126 MB) exceeds L2; no explicit flush"},
        "e2e": {"value": e2e_value, "unit": "clouds/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "roofline": roof, "rooflines": rooflines, "entry_time_share": entry_share,
        "dominant_entry": dom[0] if dom else None, "clocks": clk,
    }
    if not args.no_cpu_baseline:
        v, dt, cores = cpu_port_run(args.workload, 1, 0, *sample)
        out["cpu_baseline"] = {"value": v, "unit": "clouds/s", "cores": cores, "kind": "port",
                               "sample": f"{sample[0]} cloud(s) x N={sample[1]}, fwd+bwd, 1 step ({dt:.1f} s)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
