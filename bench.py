#!/usr/bin/env python
"""Benchmark of the ModeT registration hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--breakdown]

One "step" = ModeT.forward on one synthetic LPBA-shape pair per GPU (160x192x160 fp32,
BASELINE.json configs[1]).  N > 1 is launched by torchrun, one rank per GPU; pairs are independent,
so ranks share nothing on the data path (weak scaling, no collective); the only communication is
the max-over-ranks of the timed region.  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's CPU implementation of the same path on the host cores.
/root/reference does not exist on the GPU box, so that arm runs the oracle port
(oracle/modet_oracle.py with the torch library calls the reference itself makes).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SHAPE = (160, 192, 160)
HEADS = [8, 4, 2, 1, 1]
METRIC = "volume-pairs/sec at 160x192x160 (ModeT forward, fp32)"
UNIT = "pairs/s"
FUSED_L1_BYTES_PER_VOXEL = 80      # q 24 + k 24 + flow 12 + moving 4 read; flow' 12 + moved 4 written


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.03)

    def reset(self):
        """Forget what was sampled so far (the sampler is started during warm-up: the first NVML queries of a fresh
        process take tens of milliseconds and contend with kernel launches, which must not fall into the timed region)."""
        self.samples, self.reasons = [], set()

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_mhz_min": (min(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def physical_device_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


def cpu_reference_run(steps: int, warmup: int):
    """The reference's CPU path (oracle port, library ops) on all host cores: full pairs."""
    from oracle import modet_oracle as orc
    from smilecode_b200.synth import make_pair
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = orc.synth_state_dict(seed=1234, num_heads=HEADS)
    moving, fixed = make_pair(SHAPE, batch=1, seed=24)
    with torch.no_grad():
        small = make_pair((32, 32, 32), batch=1, seed=24)
        orc.modet_forward(*small, sd, num_heads=HEADS, scale=1.0, library_ops=True)        # thread-pool warm-up
        for _ in range(warmup):
            orc.modet_forward(moving, fixed, sd, num_heads=HEADS, scale=1.0, library_ops=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.modet_forward(moving, fixed, sd, num_heads=HEADS, scale=1.0, library_ops=True)
        dt = time.perf_counter() - t0
    return steps / dt, dt / steps, cores, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    pps, spp, cores, threads = cpu_reference_run(steps, warm)
    sample = f"{steps} full 160x192x160 pair(s) after {warm} warm-up, oracle port with torch library ops, {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(1, args.gpus),
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(batch_per_gpu: int, n_gpus: int):
    return {"workload": "LPBA-shape 160x192x160 fp32 pair, full ModeT forward (encoder + 5-level decoder), "
                        "head_dim 6, heads [8,4,2,1,1], scale 1 (BASELINE.json configs[1])",
            "pairs_per_gpu_per_step": batch_per_gpu, "global_pairs_per_step": batch_per_gpu * n_gpus,
            "parallelism": f"replicas x{n_gpus} (pairs sharded by batch, no data-path collective)",
            "l2": "flushed between timed steps (256 MiB memset outside the timed events); per-step working set >> 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--breakdown", action="store_true", help="print a per-kernel CUDA-event breakdown to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: smilecode_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)

    from smilecode_b200 import _lib, models, ops
    from smilecode_b200.synth import make_pair, randomize_weights
    _lib.lib()

    torch.manual_seed(24)
    model = models.ModeT(SHAPE, head_dim=6, num_heads=HEADS, scale=1)
    randomize_weights(model, seed=1234)
    model = model.to(dev).eval()
    moving_h, fixed_h = make_pair(SHAPE, batch=1, seed=24 + rank)
    moving_h, fixed_h = moving_h.pin_memory(), fixed_h.pin_memory()
    moving, fixed = moving_h.to(dev), fixed_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x: float) -> float:
        if world > 1:
            t = torch.tensor([x], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    clk = ClockSampler(physical_device_index(local))
    clk.__enter__()                      # sampling starts with the warm-up, its samples are dropped below
    with torch.no_grad():
        # bring the GPU out of its idle power state before the W warm-up steps (not a step of the workload: a plain
        # matmul loop for ~0.25 s; a cold B200 otherwise spends the first timed steps ramping up)
        pre = torch.randn(4096, 4096, device=dev)
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 0.25:
            for _ in range(10):
                pre = torch.tanh(pre @ pre) * 0.01
            torch.cuda.synchronize()
        del pre
        for _ in range(W):
            y, flow = model(moving, fixed)
        torch.cuda.synchronize()

        # ---------------- value: inputs resident in HBM, device-timed, L2 flushed between steps
        l0 = _lib.LAUNCHES
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        barrier()
        clk.reset()
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            y, flow = model(moving, fixed)
            b.record(stream)
        barrier()
        clk.__exit__()
        launches = _lib.LAUNCHES - l0
        step_ms = [a.elapsed_time(b) for a, b in evs]
        if os.environ.get("SMILE_BENCH_DEBUG"):
            print("step ms:", " ".join(f"{t:.2f}" for t in step_ms), file=sys.stderr)
        total_ms = reduce_max(sum(step_ms))
        value = world * K / (total_ms * 1e-3)

        # ---------------- e2e: host buffers in, host buffers out, copies inside the timed region.
        # The public host-to-host API is smilecode_b200.pipeline.RegistrationPipeline (upload / compute / download
        # streams, 2 slots): every step uploads the pair from pinned host memory and downloads moved + flow.
        from smilecode_b200.pipeline import RegistrationPipeline
        pipe = RegistrationPipeline(model, SHAPE, depth=2, device=dev)
        h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
        for _ in pipe.run([(moving_h, fixed_h)] * 3):
            pass
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n_out = 0
        for y_h, flow_h in pipe.run([(moving_h, fixed_h)] * K):
            n_out += 1
        pipe.s_out.synchronize()
        e1.record(stream)
        barrier()
        assert n_out == K
        e2e_ms = reduce_max(e0.elapsed_time(e1))
        e2e_value = world * K / (e2e_ms * 1e-3)
        flow_h = flow_h.clone()

        # ---------------- roofline of the headline kernel: fused L1 attention + compose + warp
        N1 = SHAPE[0] * SHAPE[1] * SHAPE[2]
        g = torch.Generator(device=dev).manual_seed(7)
        q1 = torch.randn(1, *SHAPE, 6, device=dev, generator=g)
        k1 = torch.randn(1, *SHAPE, 6, device=dev, generator=g)
        rpb1 = model.mdt1.rpb.detach()
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for _ in range(3):
            ops.modet_fused(q1, k1, rpb1, flow, moving, 1.0, 1.0)
        for a, b in kev:
            flush.zero_()
            a.record(stream)
            ops.modet_fused(q1, k1, rpb1, flow, moving, 1.0, 1.0)
            b.record(stream)
        torch.cuda.synchronize()
        k_ms = statistics.mean(a.elapsed_time(b) for a, b in kev)
        peak, peak_src = measured_peaks()
        achieved = FUSED_L1_BYTES_PER_VOXEL * N1 / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture
            with open(os.path.join(ROOT, "profiles", "fused_l1_traffic.json")) as f:
                tj = json.load(f)
            traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj["source"]
        except Exception:
            pass
        roofline = {"kernel": "modet_fused_fwd (L1: attention heads=1 + flow compose + warp moving)", "bound": "hbm",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                    "traffic_source": traffic_src,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": FUSED_L1_BYTES_PER_VOXEL * N1,
                    "launch_ms": k_ms}

        # ---------------- informational: the same forward with four pairs per launch (the coarse levels fill the GPU better)
        batched = None
        if rank == 0 and world == 1:
            mb, fb = moving.repeat(4, 1, 1, 1, 1), fixed.repeat(4, 1, 1, 1, 1)
            for _ in range(2):
                model(mb, fb)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(stream)
            for _ in range(5):
                model(mb, fb)
            b1.record(stream)
            torch.cuda.synchronize()
            bms = b0.elapsed_time(b1) / 5
            batched = {"pairs_per_launch": 4, "value": 4e3 / bms, "unit": UNIT, "ms_per_launch": bms,
                       "note": "not the headline: BASELINE configs[1] is one pair per forward"}
            del mb, fb

        breakdown = None
        if args.breakdown or rank == 0:
            _lib.profile_start()
            model(moving, fixed)
            prof = _lib.profile_stop()
            breakdown = sorted(((v[1], v[0], k) for k, v in prof.items()), reverse=True)
            if args.breakdown and rank == 0:
                tot = sum(b[0] for b in breakdown)
                for ms, calls, name in breakdown:
                    print(f"  {ms:9.3f} ms  {100 * ms / tot:5.1f}%  x{calls:<3d} {name}", file=sys.stderr)
                print(f"  total of kernels {tot:.3f} ms", file=sys.stderr)

    # ---------------- training step: fp32, one pair per GPU, NCC + Grad3d loss, hand-written
    # backward kernels, one flat-bucket NCCL all-reduce of the gradients when N > 1, fused Adam(amsgrad) update
    train = None
    if not args.no_train:
        from smilecode_b200.train import Trainer
        tmodel = models.ModeT(SHAPE, head_dim=6, num_heads=HEADS, scale=1)
        randomize_weights(tmodel, seed=1234)
        tmodel = tmodel.to(dev)
        trainer = Trainer(tmodel, lr=1e-4, distributed=world > 1)
        for _ in range(2):
            trainer.step(moving, fixed)
        barrier()
        l0 = _lib.LAUNCHES
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        TK = min(K, 5)
        t0.record(stream)
        for _ in range(TK):
            tloss, _, _ = trainer.step(moving, fixed)
        t1.record(stream)
        barrier()
        tms = reduce_max(t0.elapsed_time(t1)) / TK
        train = {"value": world * 1e3 / tms, "unit": "pairs/s", "ms_per_step": tms, "steps": TK,
                 "gpu_launches_per_step": (_lib.LAUNCHES - l0) // TK, "loss": float(tloss),
                 "config": "fp32 training step, 1 pair per GPU, NCC_vxm(9) + Grad3d(l2), Adam(amsgrad); "
                           + ("flat-bucket NCCL gradient all-reduce" if world > 1 else "single GPU, no collective")}
        del trainer, tmodel
        torch.cuda.empty_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        from oracle import modet_oracle as orc     # checker / CPU baseline leg only
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        with torch.no_grad():
            small = make_pair((32, 32, 32), batch=1, seed=24)
            orc.modet_forward(*small, sd, num_heads=HEADS, scale=1.0, library_ops=True)
            t0 = time.perf_counter()
            y_ref, flow_ref = orc.modet_forward(moving_h, fixed_h, sd, num_heads=HEADS, scale=1.0, library_ops=True)
            dt = time.perf_counter() - t0
            # same pair in fp64 (not timed): the exact answer both fp32 paths are measured against
            sd64 = {k: v.double() for k, v in sd.items()}
            _, flow_ref64 = orc.modet_forward(moving_h.double(), fixed_h.double(), sd64, num_heads=HEADS, scale=1.0,
                                              library_ops=True)
        err = float((flow_h - flow_ref).abs().max())
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 full 160x192x160 pair ({dt:.1f} s), oracle port with torch library ops",
               "max_abs_flow_diff_vs_gpu": err,
               "max_abs_flow_err_vs_fp64": {"gpu": float((flow_h.double() - flow_ref64).abs().max()),
                                            "cpu_fp32": float((flow_ref.double() - flow_ref64).abs().max())}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(1, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / K},
                "step_ms_min_max": [min(step_ms), max(step_ms)], "gpu_launches": launches, "gpu_launches_per_step": launches // K, "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu, "train": train, "batched": batched}
        if breakdown:
            line["kernel_ms"] = {name: round(ms, 4) for ms, _, name in breakdown[:12]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
