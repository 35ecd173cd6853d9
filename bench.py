#!/usr/bin/env python
"""Benchmark of the Pix2Pix G+D training iteration (BASELINE.json metric: "Pix2Pix G+D train
images/sec at 256^2").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:       one process per GPU (torchrun for N > 1), batch 64 per GPU, 256x256 synthetic RGB pairs,
            bf16 tensor-core kernels from libgap_b200.so; prints ONE JSON line on rank 0.
reference:  the CPU restatement of the reference's train_gan_one_epoch iteration (oracle/, a port of
            train_gan.py:52-74 + models.py; the reference itself is pure Python and cannot travel to
            the GPU box), timed on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

BATCH_PER_GPU = 64
HW = 256
GFLOP_PER_IMG = 86.789062656     # 3*G + 8*D - 2*D.0 - G.0, convolutions only
METRIC = "pix2pix_gd_train_images_per_sec_256"
UNIT = "images/s"


def _peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        d["_source"] = "measured"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    _CHILD = r'''
import sys, time
import pynvml as nv
nv.nvmlInit()
uuid, idx = sys.argv[1], int(sys.argv[2])
try:
    h = nv.nvmlDeviceGetHandleByUUID(uuid.encode())
except Exception:
    h = nv.nvmlDeviceGetHandleByIndex(idx)
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
while True:
    try:
        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, nv.nvmlDeviceGetPowerUsage(h) / 1000.0, rs,
          flush=True)
    time.sleep(0.004)
'''

    def __init__(self, index: int) -> None:
        """Starts an NVML sampler in a child process right away (it needs ~1 s to come up and, unlike a thread, does not
        compete with the launching thread for the GIL); `with sampler:` marks the window whose samples are kept."""
        self.index = index
        self.rows: list[list[str]] = []
        self._stop = threading.Event()
        self._t = None
        self._t0 = self._t1 = 0.0
        self._child = None
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(index).uuid)
            self._child = subprocess.Popen([sys.executable, "-c", self._CHILD, uuid, str(index)], stdout=subprocess.PIPE,
                                           stderr=subprocess.DEVNULL, text=True)
            import atexit
            atexit.register(self._kill_child)      # never leave the sampler behind, whatever happens to the run
        except Exception:
            self._child = None

    def _kill_child(self) -> None:
        c = self._child
        if c is not None and c.poll() is None:
            try:
                c.kill()
            except Exception:
                pass

    def _run_nvml(self) -> bool:
        """In-process NVML sampling every 10 ms (the same counters nvidia-smi prints; one nvidia-smi invocation takes
        longer than a whole default timed region).  Returns False when NVML is unavailable."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            try:        # NVML ignores CUDA_VISIBLE_DEVICES: find the device torch calls `index` by its UUID
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)).encode())
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        bits = ((0x8, 3), (0x40, 4), (0x20, 5), (0x4, 6))      # hw_slowdown, hw_thermal, sw_thermal, sw_power_cap
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                row = [str(sm), str(mx), str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0), "", "", "", ""]
                for mask, col in bits:
                    row[col] = "Active" if rs & mask else "Not Active"
                self.rows.append(row)
            except Exception:
                pass
            self._stop.wait(0.01)
        return True

    def _run(self) -> None:
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([c.strip() for c in out.stdout.strip().splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t0 = time.time()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._t1 = time.time()
        self._stop.set()
        self._t.join(timeout=6)
        if self._child is not None:
            try:
                self._child.terminate()
                out, _ = self._child.communicate(timeout=5)
            except Exception:
                out = ""
            rows = []
            for line in out.splitlines():
                f = line.split()
                if len(f) != 5:
                    continue
                try:
                    ts, sm, mx, pw, rs = float(f[0]), f[1], f[2], f[3], int(f[4])
                except ValueError:
                    continue
                if self._t0 <= ts <= self._t1:
                    act = lambda m: "Active" if rs & m else "Not Active"
                    rows.append([sm, mx, pw, act(0x8), act(0x40), act(0x20), act(0x4)])
            if len(rows) > len(self.rows):      # the child saw more of the window than the in-process thread
                self.rows = rows

    def summary(self) -> dict:
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = max((float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = sorted(float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                "samples": len(self.rows), "power_w": pw[len(pw) // 2] if pw else None}


def _cpu_baseline(seconds_budget: float = 20.0) -> dict:
    """Oracle port of one train_gan_one_epoch iteration (batch 1, 256^2, fp32) on the host cores."""
    from oracle import pix2pix_oracle as O
    from gan_aug_pfa_b200 import spec as MI
    torch.manual_seed(0)
    sd_g, sd_d = MI.default_state_dicts()
    og = O.AdamState(sd_g, O.param_names(sd_g), 1e-4, (0.5, 0.999))
    od = O.AdamState(sd_d, O.param_names(sd_d), 1e-4, (0.5, 0.999))
    gen = torch.Generator().manual_seed(1234)
    A = torch.rand(1, 3, HW, HW, generator=gen) * 2 - 1
    B = torch.rand(1, 3, HW, HW, generator=gen) * 2 - 1
    O.gan_train_step(sd_g, sd_d, og, od, A, B)      # warm-up
    t0 = time.perf_counter()
    it = 0
    while True:
        O.gan_train_step(sd_g, sd_d, og, od, A, B)
        it += 1
        el = time.perf_counter() - t0
        if el > seconds_budget or it >= 40:
            break
    return {"value": it / el, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{it} iterations of batch 1 at 256x256 fp32 (oracle port of train_gan.py:52-74), "
                      f"{el:.1f} s on {os.cpu_count()} host cpus"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pix2pix_oracle as O
    from gan_aug_pfa_b200 import spec as MI
    torch.manual_seed(0)
    sd_g, sd_d = MI.default_state_dicts()
    og = O.AdamState(sd_g, O.param_names(sd_g), 1e-4, (0.5, 0.999))
    od = O.AdamState(sd_d, O.param_names(sd_d), 1e-4, (0.5, 0.999))
    sample_batch = 2
    gen = torch.Generator().manual_seed(1234)
    A = torch.rand(sample_batch, 3, HW, HW, generator=gen) * 2 - 1
    B = torch.rand(sample_batch, 3, HW, HW, generator=gen) * 2 - 1
    for _ in range(max(1, min(args.warmup, 2))):
        O.gan_train_step(sd_g, sd_d, og, od, A, B)
    steps = args.steps
    t0 = time.perf_counter()
    for _ in range(steps):
        O.gan_train_step(sd_g, sd_d, og, od, A, B)
    el = time.perf_counter() - t0
    val = steps * sample_batch / el
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pix2pix_gan_train_b64_256x256", "sample_batch_per_step": sample_batch,
                   "note": "CPU oracle port of train_gan_one_epoch; each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} steps of batch {sample_batch} at 256x256 fp32"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args) -> None:
    import torch.distributed as dist
    from gan_aug_pfa_b200 import _lib, ops
    from gan_aug_pfa_b200.pix2pix import Pix2PixTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the gap_* kernels have no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    clock_sampler = ClockSampler(local)       # the child process needs a moment to come up: start it before the warm-up
    torch.manual_seed(0)
    tr = Pix2PixTrainer(dev, world=world)     # world > 1: bucketed NCCL all-reduce overlapped with the backward pass
    N = BATCH_PER_GPU
    gen = torch.Generator().manual_seed(1234 + rank)
    n_batches = 2
    if args.inputs == "u8":     # raw uint8 images; ToTensor + JointNormalize run on the device
        host = [(torch.randint(0, 256, (N, HW, HW, 3), generator=gen, dtype=torch.uint8).pin_memory(),
                 torch.randint(0, 256, (N, HW, HW, 3), generator=gen, dtype=torch.uint8).pin_memory())
                for _ in range(n_batches)]
    else:                       # fp32 NCHW in [-1, 1]: what the reference's DataLoader hands to train_gan.py:53-54
        host = [(torch.rand(N, 3, HW, HW, generator=gen).mul_(2).sub_(1).pin_memory(),
                 torch.rand(N, 3, HW, HW, generator=gen).mul_(2).sub_(1).pin_memory()) for _ in range(n_batches)]
    devb = [(a.to(dev), b.to(dev)) for a, b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # Single GPU: the iteration's ~175 launches can be replayed from one CUDA graph (same kernels, no launch gaps) or
    # launched eagerly (three streams; the host enqueues a step faster than the GPU runs it).  Which is faster depends
    # on the host: calibrate both during warm-up (untimed, interleaved so clock drift cancels) and keep the faster.
    step_fn, use_graph = tr.train_step, False
    for i in range(args.warmup):
        step_fn(*devb[i % n_batches])
    if world == 1 and args.step_mode != "eager":
        use_graph = True
        if args.step_mode == "auto":
            def block(fn, n=5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                a.record()
                for i in range(n):
                    fn(*devb[i % n_batches])
                b.record()
                torch.cuda.synchronize()
                return a.elapsed_time(b)
            block(tr.train_step_graphed, 3)                      # capture + first replays
            t_e = t_g = 0.0
            for r in range(4):                                   # ABBA order: clock drift cancels
                if r % 2 == 0:
                    t_e += block(tr.train_step)
                    t_g += block(tr.train_step_graphed)
                else:
                    t_g += block(tr.train_step_graphed)
                    t_e += block(tr.train_step)
            use_graph = t_g < t_e
        if use_graph:
            step_fn = tr.train_step_graphed
            for i in range(3):
                step_fn(*devb[i % n_batches])
    # ---- device-resident timing
    barrier()
    _lib.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clock_sampler as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            losses = step_fn(*devb[i % n_batches])
        e1.record()
        barrier()
    launches = _lib.LAUNCHES
    if use_graph:
        # graph replays do not pass through the Python launchers: count the launches of one eager iteration
        _lib.LAUNCHES = 0
        tr.train_step(*devb[0])
        launches = _lib.LAUNCHES * args.steps
    ms = max_over_ranks(e0.elapsed_time(e1))
    value = world * N * args.steps / (ms * 1e-3)
    loss_host = losses.cpu().tolist()

    # ---- end to end: pinned host inputs, H2D inside the timed region (prefetched on a copy stream into two staging
    # buffers while the previous iteration computes), losses read back to the host every step
    stage = [(torch.empty_like(devb[0][0]), torch.empty_like(devb[0][1])) for _ in range(2)]
    loss_pinned = torch.empty(2, dtype=torch.float64).pin_memory()
    cur = torch.cuda.current_stream()
    copy_s = torch.cuda.Stream(dev)
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_used = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(i):
        k = i % 2
        with torch.cuda.stream(copy_s):
            if i >= 2:
                copy_s.wait_event(ev_used[k])            # iteration i-2 has consumed this staging buffer
            stage[k][0].copy_(host[i % n_batches][0], non_blocking=True)
            stage[k][1].copy_(host[i % n_batches][1], non_blocking=True)
            ev_copied[k].record(copy_s)

    barrier()
    e0.record()
    copy_s.wait_stream(cur)
    prefetch(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            prefetch(i + 1)
        cur.wait_event(ev_copied[i % 2])
        out = step_fn(*stage[i % 2])
        ev_used[i % 2].record(cur)
        loss_pinned.copy_(out, non_blocking=True)
        cur.synchronize()                                # the caller reads the step's losses (train_gan.py:72-74)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * N * args.steps / (ms_e2e * 1e-3)

    # ---- per-kernel roofline: time every GEMM launch of a few steps with CUDA events
    # (side streams serialised for these steps: with the generator forward / the wgrads overlapping other kernels a
    # launch's event-to-event time would include the work it shares the GPU with)
    prof_steps = min(2, args.steps)
    saved = (tr.overlap_g_fwd, tr.G.overlap_wgrad, tr.D.overlap_wgrad)
    tr.overlap_g_fwd = tr.G.overlap_wgrad = tr.D.overlap_wgrad = False
    tr.train_step(*devb[0])
    ops.PROFILE = []
    for i in range(prof_steps):
        tr.train_step(*devb[i % n_batches])
    torch.cuda.synchronize()
    prof = ops.PROFILE
    ops.PROFILE = None
    tr.overlap_g_fwd, tr.G.overlap_wgrad, tr.D.overlap_wgrad = saved
    agg: dict = {}
    for name, flops, ev0, ev1 in prof:
        d = agg.setdefault(name, [0.0, 0.0, 0])
        d[0] += flops
        d[1] += ev0.elapsed_time(ev1) * 1e-3
        d[2] += 1
    peaks = _peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    kern = {k: {"tflops": v[0] / v[1] / 1e12 if v[1] > 0 else 0.0, "seconds_per_step": v[1] / prof_steps,
                "launches_per_step": v[2] // prof_steps} for k, v in agg.items()}
    dom = max(agg.items(), key=lambda kv: kv[1][1])[0] if agg else None
    roofline = None
    if dom is not None:
        ach = kern[dom]["tflops"]
        traffic = None
        tpath = ROOT / "profiles" / "r1_roofline_traffic.json"
        if tpath.exists():
            traffic = json.loads(tpath.read_text()).get(dom, {}).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "traffic": traffic,
                    "traffic_source": "ncu dram__bytes_read+write per launch, averaged over the step's launches "
                                      "(profiles/r1_step_metrics_ncu.csv)" if traffic else None,
                    "algorithmic_flop_per_launch": agg[dom][0] / max(1, agg[dom][2]),
                    "peak_source": f"{peaks['_source']} bf16_tflops_sustained (kernel timed inside a long step)",
                    "per_kernel": kern,
                    "step_frac_of_peak": (value / world) * GFLOP_PER_IMG * 1e9 / (peak_tf * 1e12)}

    if rank == 0:
        cpu = _cpu_baseline() if world == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "pix2pix_gan_train_b64_256x256", "batch_per_gpu": N, "image": f"{HW}x{HW}x3",
                       "parallelism": f"dp{world}", "cuda_graph": use_graph, "inputs": args.inputs, "l2": "per-step working set (several GB of activations) >> 126 MB L2; "
                       "two alternating input batches", "algorithmic_gflop_per_image": GFLOP_PER_IMG},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0]),
                    "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": roofline,
            "final_losses": {"loss_d": loss_host[0], "loss_g": loss_host[1]},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_secondary(args) -> None:
    """Secondary workloads of BASELINE.json's configs 4 and 5 (single GPU or independent replicas): generator
    inference throughput and the Siamese U-Net training step.  Same JSON schema, their own metric names."""
    import torch.distributed as dist
    from gan_aug_pfa_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    clock_sampler = ClockSampler(local)       # started early: the sampling child needs a moment to come up
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(1234 + rank)
    if args.workload == "gen_infer":
        from gan_aug_pfa_b200.pix2pix import GeneratorEngine
        N = 256
        eng = GeneratorEngine(dev)
        eng.training = True
        warm = (torch.rand(8, 3, HW, HW, generator=gen) * 2 - 1).to(dev)
        for _ in range(2):
            eng.forward(warm)                      # non-trivial running statistics, as after training
        eng.training = False
        # device-resident leg: fp32 NCHW inputs (the models.py contract).  End-to-end leg: what generate_synthetic_data.py's
        # loop moves — uint8 HWC images in (normalised on the device) and uint8 HWC images out (written by the last layer's
        # epilogue), pinned host buffers, H2D / D2H on copy streams overlapped with the previous / next batch's compute.
        host = [(torch.rand(N, 3, HW, HW, generator=gen) * 2 - 1).pin_memory() for _ in range(2)]
        devb = [h.to(dev) for h in host]
        host_u8 = [torch.randint(0, 256, (N, HW, HW, 3), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(2)]
        in_u8 = [torch.empty(N, HW, HW, 3, device=dev, dtype=torch.uint8) for _ in range(2)]
        out_u8 = [torch.empty(N, HW, HW, 3, device=dev, dtype=torch.uint8) for _ in range(2)]
        out_host = [torch.empty(N, HW, HW, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
        cur = torch.cuda.current_stream()
        h2d_s, d2h_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev_in = [torch.cuda.Event(), torch.cuda.Event()]
        ev_done = [torch.cuda.Event(), torch.cuda.Event()]
        ev_out = [torch.cuda.Event(), torch.cuda.Event()]
        state = {"i": 0}

        def step(i, e2e=False):
            if not e2e:
                eng.forward(devb[i % 2])
                return
            k = i % 2
            if state["i"] == 0:                         # first e2e call: start the input pipeline
                h2d_s.wait_stream(cur)
                d2h_s.wait_stream(cur)
                with torch.cuda.stream(h2d_s):
                    in_u8[k].copy_(host_u8[k], non_blocking=True)
                    ev_in[k].record(h2d_s)
            state["i"] += 1
            with torch.cuda.stream(h2d_s):              # prefetch the next batch while this one computes
                h2d_s.wait_event(ev_done[1 - k])        # (its buffer was consumed by the previous forward)
                in_u8[1 - k].copy_(host_u8[1 - k], non_blocking=True)
                ev_in[1 - k].record(h2d_s)
            cur.wait_event(ev_in[k])
            cur.wait_event(ev_out[k])                   # the D2H of two batches ago has drained out_u8[k]
            eng.forward(in_u8[k], out_u8=out_u8[k])
            ev_done[k].record(cur)
            with torch.cuda.stream(d2h_s):
                d2h_s.wait_event(ev_done[k])
                out_host[k].copy_(out_u8[k], non_blocking=True)
                ev_out[k].record(d2h_s)
            if i == args.steps - 1:
                d2h_s.synchronize()                     # every image of the timed region has reached the host
                cur.wait_stream(d2h_s)
        units, metric, unit = N, "generator_inference_images_per_sec_256", "images/s"
        gflop = 12.046041088
        h2d, d2h = N * 3 * HW * HW, N * 3 * HW * HW          # uint8 images both ways
        cfg = {"workload": "generator_inference_b256_256x256", "batch_per_gpu": N, "bn": "eval (running statistics)"}
    else:
        from gan_aug_pfa_b200.siamese import SiameseEngine
        N, S = 4, 512
        eng = SiameseEngine(dev)
        from gan_aug_pfa_b200 import models as M
        torch.manual_seed(0)
        eng.load_state_dict({k: v.detach() for k, v in M.SiameseUNet(3, 1).state_dict().items()})
        host = [((torch.rand(N, 3, S, S, generator=gen) * 2 - 1).pin_memory(),
                 (torch.rand(N, 3, S, S, generator=gen) * 2 - 1).pin_memory(),
                 (torch.rand(N, S, S, generator=gen) < 0.05).long().pin_memory()) for _ in range(2)]
        devb = [tuple(t.to(dev) for t in b) for b in host]
        stage = tuple(torch.empty_like(t) for t in devb[0])
        loss_host = torch.empty(1, dtype=torch.float64).pin_memory()

        def step(i, e2e=False):
            b = devb[i % 2]
            if e2e:
                for dst, src in zip(stage, host[i % 2]):
                    dst.copy_(src, non_blocking=True)
                b = stage
            loss = eng.train_step(*b, kind="combined")
            if e2e:
                loss_host.copy_(loss, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        units, metric, unit = N, "siamese_train_pairs_per_sec_512", "pairs/s"
        gflop = 2207.23
        h2d, d2h = 2 * N * 3 * S * S * 4 + N * S * S * 8, 8
        cfg = {"workload": "siamese_unet_train_b4_512x512_combined_loss", "batch_per_gpu": N, "loss": "CombinedLoss(0.5, 1:9)",
               "optimizer": "AdamW(1.0152e-4, wd 1.118e-5)"}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    _lib.LAUNCHES = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with clock_sampler as clocks:
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
    launches = _lib.LAUNCHES
    ms = e0.elapsed_time(e1)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i, e2e=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    value = world * units * args.steps / (ms * 1e-3)
    peaks = _peaks()
    peak_tf = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])
    if rank == 0:
        print(json.dumps({
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": cfg, "clocks": clocks.summary(),
            "e2e": {"value": world * units * args.steps / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": (value / world) * gflop * 1e9 / 1e12, "peak": peak_tf,
                         "unit": "TFLOP/s", "frac": (value / world) * gflop * 1e9 / (peak_tf * 1e12), "traffic": None,
                         "note": "whole-step algorithmic conv FLOP/s (SURVEY.md §8d) over the measured sustained bf16 peak"},
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pix2pix_train", choices=["pix2pix_train", "gen_infer", "siamese_train"],
                    help="pix2pix_train = the BASELINE.json metric (default); gen_infer = generate_synthetic_data.py's "
                         "generator forward at batch 256 (config 4); siamese_train = train.py's step at 512x512, batch 4 "
                         "(config 5, CombinedLoss)")
    ap.add_argument("--step-mode", default="auto", choices=["auto", "graph", "eager"],
                    help="single-GPU pix2pix_train: CUDA-graph replay, eager launches, or (default) whichever measures "
                         "faster in an untimed calibration after the warm-up")
    ap.add_argument("--inputs", default="f32", choices=["f32", "u8"],
                    help="pix2pix_train only: host batches as fp32 NCHW in [-1,1] (the reference DataLoader's output, "
                         "default) or raw uint8 HWC images normalised on the device")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "pix2pix_train":
        run_ours(args)
    else:
        run_secondary(args)


if __name__ == "__main__":
    main()
