#!/usr/bin/env python
"""Benchmark of the ModeT registration hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config lpba|mindboggle]
                    [--train-batch B] [--breakdown] [--comparators]

One "step" = ModeT.forward on one synthetic pair per GPU: `--config lpba` (default) is BASELINE.json configs[1]
(160x192x160 fp32, heads [8,4,2,1,1]); `--config mindboggle` is the configs[4] shape (160x192x224 fp32 with the 6-head
level list [6,6,6,1,1]: levels 2 and 1 have no CWM in ModeT, so their head count is 1 -- SURVEY 8).  N > 1 is launched by
torchrun, one rank per GPU; pairs are independent, so ranks share nothing on the data path (weak scaling, no collective);
the only communication is the max-over-ranks of the timed region.  Prints ONE JSON line on rank 0.

`--impl reference` times the reference's own CPU implementation of the same path on the host cores: the UNMODIFIED
`ModeT/models.py` from the staged reference tree (baseline/_ref, written by oracle/stage_reference.py; kind "reference")
when it is there, otherwise the oracle port (oracle/modet_oracle.py with the torch library calls the reference itself
makes; kind "port").
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

CONFIGS = {
    "lpba": {"shape": (160, 192, 160), "heads": [8, 4, 2, 1, 1],
             "workload": "LPBA-shape 160x192x160 fp32 pair, full ModeT forward (encoder + 5-level decoder), "
                         "head_dim 6, heads [8,4,2,1,1], scale 1 (BASELINE.json configs[1])"},
    "mindboggle": {"shape": (160, 192, 224), "heads": [6, 6, 6, 1, 1],
                   "workload": "Mindboggle-shape 160x192x224 fp32 pair, full ModeT forward, head_dim 6, 6-head level list "
                               "[6,6,6,1,1] (levels 2 and 1 have no CWM, so 1 head), scale 1 (BASELINE.json configs[4] shape)"},
}
SHAPE = CONFIGS["lpba"]["shape"]
HEADS = CONFIGS["lpba"]["heads"]
WORKLOAD = CONFIGS["lpba"]["workload"]
METRIC = "volume-pairs/sec at 160x192x160 (ModeT forward, fp32)"
UNIT = "pairs/s"
FUSED_L1_BYTES_PER_VOXEL = 80      # q 24 + k 24 + flow 12 + moving 4 read; flow' 12 + moved 4 written


def select_config(name: str):
    global SHAPE, HEADS, WORKLOAD, METRIC
    c = CONFIGS[name]
    SHAPE, HEADS, WORKLOAD = c["shape"], c["heads"], c["workload"]
    METRIC = f"volume-pairs/sec at {SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]} (ModeT forward, fp32)"


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


def reference_forward_fn(sd):
    """(callable(moving, fixed) -> (moved, flow), kind, description): the staged reference's own ModeT when available."""
    from oracle import modet_oracle as orc
    try:
        from oracle import reference_loader as rl
        ref = rl.reference_models()
    except Exception:
        ref = None
    if ref is not None:
        m = ref.ModeT(SHAPE, head_dim=6, num_heads=HEADS, scale=1)
        missing, unexpected = m.load_state_dict(sd, strict=False)
        assert not unexpected and all(k.endswith("grid") for k in missing), (missing, unexpected)
        m.eval()
        return (lambda a, b: m(a, b)), "reference", "the reference's own ModeT/models.py (unmodified, staged under baseline/_ref)"
    fn = lambda a, b: orc.modet_forward(a, b, sd, num_heads=HEADS, scale=1.0, library_ops=True)
    return fn, "port", "oracle port with torch library ops (reference tree not staged on this box)"


def cpu_reference_run(steps: int, warmup: int):
    """The reference's CPU path on all host cores: full pairs."""
    from oracle import modet_oracle as orc
    from smilecode_b200.synth import make_pair
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = orc.synth_state_dict(seed=1234, num_heads=HEADS)
    fn, kind, desc = reference_forward_fn(sd)
    moving, fixed = make_pair(SHAPE, batch=1, seed=24)
    with torch.no_grad():
        small = make_pair((32, 32, 32), batch=1, seed=24)
        orc.modet_forward(*small, sd, num_heads=HEADS, scale=1.0, library_ops=True)        # thread-pool warm-up
        for _ in range(warmup):
            fn(moving, fixed)
        t0 = time.perf_counter()
        for _ in range(steps):
            fn(moving, fixed)
        dt = time.perf_counter() - t0
    return steps / dt, dt / steps, cores, torch.get_num_threads(), kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    pps, spp, cores, threads, kind, desc = cpu_reference_run(steps, warm)
    sample = f"{steps} full {SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]} pair(s) after {warm} warm-up, {desc}, {threads} threads"
    line = {"impl": "reference", "metric": METRIC, "value": pps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": spp * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(1, args.gpus),
            "cpu_baseline": {"value": pps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": pps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def gpu_comparators(model, moving, fixed, dev):
    """The staged reference on this GPU, same pair and weights (informational; None entries when pieces are missing)."""
    out = {"ref_eager_ms": None, "ref_cu_ms": None}
    try:
        from oracle import reference_loader as rl
    except Exception:
        return out
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

    def timed(m):
        m = m.to(dev).eval()
        with torch.no_grad():
            for _ in range(2):
                m(moving, fixed)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                m(moving, fixed)
            b.record()
            torch.cuda.synchronize()
        return a.elapsed_time(b) / 3

    for key, loader, cls in (("ref_eager_ms", rl.reference_models, "ModeT"), ("ref_cu_ms", rl.reference_models_cu, "ModeT_cu")):
        try:
            mod = loader()
            if mod is None:
                continue
            m = getattr(mod, cls)(SHAPE, head_dim=6, num_heads=HEADS, scale=1)
            m.load_state_dict(sd, strict=False)
            out[key] = timed(m)
            del m
            torch.cuda.empty_cache()
        except Exception as e:           # informational leg: never fail the bench
            out[key + "_error"] = repr(e)[:200]
    out["note"] = "reference ModeT in PyTorch eager (TF32 off) and ModeT-cu with its own modet extension (sm_100 build), ms per forward"
    return out


def workload_config(batch_per_gpu: int, n_gpus: int):
    return {"workload": WORKLOAD,
            "pairs_per_gpu_per_step": batch_per_gpu, "global_pairs_per_step": batch_per_gpu * n_gpus,
            "parallelism": f"replicas x{n_gpus} (pairs sharded by batch, no data-path collective)",
            "l2": "flushed between timed steps (256 MiB memset outside the timed events); per-step working set >> 126 MB L2",
            "launch": "CUDA-graph replay of the forward (value, e2e); kernel-by-kernel number under `eager`"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--breakdown", action="store_true", help="print a per-kernel CUDA-event breakdown to stderr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    ap.add_argument("--config", default="lpba", choices=sorted(CONFIGS))
    ap.add_argument("--train-batch", type=int, default=1, help="pairs per GPU in the training-step leg")
    ap.add_argument("--train-dtype", default="fp32", choices=["fp32", "bf16"],
                    help="bf16: Conv3d forward and data-gradient products on bf16 tensor cores, fp32 accumulation "
                         "(BASELINE.json configs[2..3]); everything else fp32")
    ap.add_argument("--comparators", action="store_true",
                    help="also time the staged reference on this GPU (PyTorch eager, and ModeT-cu with its own extension)")
    args = ap.parse_args()
    select_config(args.config)
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
        # bring the GPU out of its idle power state before the W warm-up steps: ~0.25 s of the workload itself (a cold B200
        # otherwise spends the first timed steps ramping up).  The hot path's own kernels, so that a launch list of this
        # process holds nothing but them.
        t_pre = time.perf_counter()
        while time.perf_counter() - t_pre < 0.25:
            for _ in range(5):
                model(moving, fixed)
            torch.cuda.synchronize()
        for _ in range(W):
            y, flow = model(moving, fixed)
        torch.cuda.synchronize()

        # ---------------- value: inputs resident in HBM, device-timed, L2 flushed between steps.
        # Steady-state inference replays a CUDA graph of the forward (smilecode_b200.graph.GraphedForward: the same ~50
        # kernels, their launches recorded once); the kernel-by-kernel (eager) number is reported next to it.
        from smilecode_b200.graph import GraphedForward

        def timed_steps(step_fn):
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
            barrier()
            for a, b in evs:
                flush.zero_()
                a.record(stream)
                step_fn()
                b.record(stream)
            barrier()
            return [a.elapsed_time(b) for a, b in evs]

        l0 = _lib.LAUNCHES
        eager_ms = timed_steps(lambda: model(moving, fixed))
        launches_eager = _lib.LAUNCHES - l0
        graphed = GraphedForward(model, moving, fixed)
        for _ in range(2):
            graphed.replay()
        clk.reset()
        l0 = _lib.LAUNCHES
        step_ms = timed_steps(graphed.replay)
        clk.__exit__()
        launches = _lib.LAUNCHES - l0
        assert launches == launches_eager, (launches, launches_eager)
        y, flow = graphed.replay()
        y, flow = y.clone(), flow.clone()
        if os.environ.get("SMILE_BENCH_DEBUG"):
            print("step ms:", " ".join(f"{t:.2f}" for t in step_ms), file=sys.stderr)
        total_ms = reduce_max(sum(step_ms))
        value = world * K / (total_ms * 1e-3)
        eager_total = reduce_max(sum(eager_ms))
        eager = {"value": world * K / (eager_total * 1e-3), "unit": UNIT, "ms_per_step": eager_total / K,
                 "note": "same forward launched kernel by kernel from Python (no CUDA graph)"}
        del graphed

        # ---------------- e2e: host buffers in, host buffers out, copies inside the timed region.
        # The public host-to-host API is smilecode_b200.pipeline.RegistrationPipeline (upload / compute / download
        # streams, 3 slots).  Headline mode = what the reference's loop moves per pair (infer.py:79-89): the pair goes up from
        # pinned host memory, the deformation field comes back into pinned host memory.  Two more modes are reported next
        # to it: both outputs of ModeT.forward down (round-1 behaviour), and metrics-only (the Jacobian fold count of
        # infer.py:89-90 evaluated on the device by smilecode_b200.metrics: 8 bytes down).
        from smilecode_b200 import metrics as smetrics
        from smilecode_b200.pipeline import RegistrationPipeline

        def time_pipeline(pipe):
            for _ in pipe.run([(moving_h, fixed_h)] * 3):
                pass
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            n_out, last = 0, None
            for last in pipe.run([(moving_h, fixed_h)] * K):
                n_out += 1
            pipe.s_out.synchronize()
            e1.record(stream)
            barrier()
            assert n_out == K
            ms = reduce_max(e0.elapsed_time(e1))
            return world * K / (ms * 1e-3), ms / K, last

        pipe = RegistrationPipeline(model, SHAPE, depth=3, device=dev, outputs=("flow",))
        h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
        e2e_value, e2e_ms_step, flow_h = time_pipeline(pipe)
        flow_h = flow_h.clone()
        del pipe
        e2e_modes = {}
        pipe = RegistrationPipeline(model, SHAPE, depth=3, device=dev, outputs=("moved", "flow"))
        v, ms, _ = time_pipeline(pipe)
        e2e_modes["moved+flow"] = {"value": v, "ms_per_step": ms, "d2h_bytes_per_step": pipe.d2h_bytes}
        del pipe
        pipe = RegistrationPipeline(model, SHAPE, depth=3, device=dev, outputs=(),
                                    reduce=lambda moved, flow, extra: smetrics.jacobian_determinant_vxm(flow, want_det=False)[1])
        v, ms, _ = time_pipeline(pipe)
        e2e_modes["metrics_only"] = {"value": v, "ms_per_step": ms, "d2h_bytes_per_step": 8,
                                     "what": "non-positive Jacobian count of infer.py:89-90 computed on the device"}
        del pipe
        torch.cuda.empty_cache()

        # ---------------- roofline of the headline kernel: fused L1 attention + compose + warp
        N1 = SHAPE[0] * SHAPE[1] * SHAPE[2]
        g = torch.Generator(device=dev).manual_seed(7)
        pb1 = model.projblock1
        lnp = {"ln_gamma": pb1.norm.weight.detach(), "ln_beta": pb1.norm.bias.detach()}   # as ModeT.forward calls it
        lnf = lambda t: torch.nn.functional.layer_norm(t, (6,), lnp["ln_gamma"], lnp["ln_beta"])
        q1 = lnf(torch.randn(1, *SHAPE, 6, device=dev, generator=g))
        k1 = lnf(torch.randn(1, *SHAPE, 6, device=dev, generator=g))
        rpb1 = model.mdt1.rpb.detach()
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for _ in range(3):
            ops.modet_fused(q1, k1, rpb1, flow, moving, 1.0, 1.0, **lnp)
        for a, b in kev:
            flush.zero_()
            a.record(stream)
            ops.modet_fused(q1, k1, rpb1, flow, moving, 1.0, 1.0, **lnp)
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
            if tuple(tj.get("shape", CONFIGS["lpba"]["shape"])) != tuple(SHAPE):
                traffic, traffic_src = None, None          # the capture is of the LPBA shape
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

        # ---------------- informational: the same forward with the convolutions on bf16 tensor cores
        bf16_fwd = None
        if rank == 0 and world == 1:
            model.conv_precision = "bf16"
            for _ in range(2):
                _, flow_b = model(moving, fixed)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            b0.record(stream)
            for _ in range(5):
                _, flow_b = model(moving, fixed)
            b1.record(stream)
            torch.cuda.synchronize()
            model.conv_precision = "fp32"
            bms = b0.elapsed_time(b1) / 5
            bf16_fwd = {"value": 1e3 / bms, "unit": UNIT, "ms_per_step": bms,
                        "max_abs_flow_diff_vs_fp32_path": float((flow_b - flow).abs().max()),
                        "note": "conv_precision='bf16': Conv3d products on tcgen05 kind::f16; not the headline (fp32 is the reference's dtype)"}
            del flow_b

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

    # ---------------- training step: NCC + Grad3d loss, hand-written backward kernels, one
    # flat-bucket NCCL all-reduce of the gradients when N > 1, fused Adam(amsgrad) update.  Two blocks:
    #   train       fp32, --train-batch pairs per GPU (default 1: the reference's own recipe, train.py:43)
    #   train_bf16  BASELINE.json configs[2] at N = 1 (batch 8 on one GPU) / configs[3] at N > 1 (global batch 32, sharded by
    #               batch): Conv3d forward + data-gradient products on bf16 tensor cores (tcgen05 kind::f16), fp32 accumulate
    train, train_bf16 = None, None
    if not args.no_train:
        from smilecode_b200.train import Trainer

        def train_leg(dtype: str, TB: int, label: str):
            tmodel = models.ModeT(SHAPE, head_dim=6, num_heads=HEADS, scale=1)
            randomize_weights(tmodel, seed=1234)
            tmodel = tmodel.to(dev)
            tmodel.conv_precision = dtype
            trainer = Trainer(tmodel, lr=1e-4, distributed=world > 1)
            mv, fx = moving, fixed
            if TB > 1:
                mvs, fxs = zip(*[make_pair(SHAPE, batch=1, seed=24 + rank + 100 * i) for i in range(TB)])
                mv, fx = torch.cat(mvs).to(dev), torch.cat(fxs).to(dev)
            for _ in range(2):
                trainer.step(mv, fx)
            barrier()
            l0 = _lib.LAUNCHES
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            TK = min(K, 5)
            t0.record(stream)
            for _ in range(TK):
                tloss, _, _ = trainer.step(mv, fx)
            t1.record(stream)
            barrier()
            tms = reduce_max(t0.elapsed_time(t1)) / TK
            out = {"value": world * TB * 1e3 / tms, "unit": "pairs/s", "ms_per_step": tms, "steps": TK,
                   "pairs_per_gpu_per_step": TB, "global_pairs_per_step": TB * world,
                   "dtype": "f32" if dtype == "fp32" else "bf16 conv MMA operands (tcgen05 kind::f16), f32 accumulation / "
                                                         "activations / everything else",
                   "gpu_launches_per_step": (_lib.LAUNCHES - l0) // TK, "loss": float(tloss),
                   "config": f"{label}: {dtype} training step, {TB} pair(s) per GPU, NCC_vxm(9) + Grad3d(l2), Adam(amsgrad); "
                             + ("flat-bucket NCCL gradient all-reduce" if world > 1 else "single GPU, no collective")}
            del trainer, tmodel, mv, fx
            torch.cuda.empty_cache()
            return out

        train = train_leg(args.train_dtype, max(1, args.train_batch), "reference recipe (train.py:43, batch 1)"
                          if args.train_batch <= 1 else "batched")
        if args.config == "lpba":
            tb = 8 if world == 1 else max(1, 32 // world)
            train_bf16 = train_leg("bf16", tb, "BASELINE.json configs[2] (batch-8 bf16 step, 1 GPU)" if world == 1 else
                                   f"BASELINE.json configs[3] (global batch {tb * world} bf16 sharded over {world} GPUs)")

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        from oracle import modet_oracle as orc     # checker / CPU baseline leg only
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        ref_fn, kind, desc = reference_forward_fn(sd)
        with torch.no_grad():
            small = make_pair((32, 32, 32), batch=1, seed=24)
            orc.modet_forward(*small, sd, num_heads=HEADS, scale=1.0, library_ops=True)
            t0 = time.perf_counter()
            y_ref, flow_ref = ref_fn(moving_h, fixed_h)
            dt = time.perf_counter() - t0
            # same pair in fp64 (not timed): the exact answer both fp32 paths are measured against
            sd64 = {k: v.double() for k, v in sd.items()}
            _, flow_ref64 = orc.modet_forward(moving_h.double(), fixed_h.double(), sd64, num_heads=HEADS, scale=1.0,
                                              library_ops=True)
        err = float((flow_h - flow_ref).abs().max())
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
               "sample": f"1 full {SHAPE[0]}x{SHAPE[1]}x{SHAPE[2]} pair ({dt:.1f} s), {desc}",
               "max_abs_flow_diff_vs_gpu": err, "rel_flow_diff_vs_gpu": err / float(flow_ref64.abs().max()),
               "max_abs_flow_err_vs_fp64": {"gpu": float((flow_h.double() - flow_ref64).abs().max()),
                                            "cpu_fp32": float((flow_ref.double() - flow_ref64).abs().max())}}

    # ---------------- informational GPU comparators: the staged reference on this GPU
    comparators = None
    if args.comparators and rank == 0 and world == 1:
        comparators = gpu_comparators(model, moving, fixed, dev)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": workload_config(1, world),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms_step,
                        "what": "RegistrationPipeline(outputs=('flow',)): pair up from pinned host memory, flow down into "
                                "pinned host memory (what infer.py:79-89 moves per pair)", "other_modes": e2e_modes},
                "step_ms_min_max": [min(step_ms), max(step_ms)], "gpu_launches": launches, "gpu_launches_per_step": launches // K, "clocks": clk.summary(), "roofline": roofline, "eager": eager, "cpu_baseline": cpu, "train": train, "train_bf16": train_bf16, "batched": batched,
                "bf16_forward": bf16_fwd,
                "comparators": comparators}
        if breakdown:
            line["kernel_ms"] = {name: round(ms, 4) for ms, _, name in breakdown[:12]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
