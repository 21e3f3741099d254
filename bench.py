#!/usr/bin/env python
"""bench.py — image-pairs/sec of the IRR-PWC inference hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1|2|3|4|5] [--impl ours|reference] [--no-graph]
                    [--math fp32|3xtf32|tf32|3xf16]
    N>1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              bench.py --gpus N --steps K --warmup W

`--config` selects one of BASELINE.json's five configs (default 3 = the one the metric is quoted on):
  1  correlation op alone, 2 x (1,64,64,128)                        a step = one cost-volume launch
  2  PWCNet (pwcnet.py), 256x256, batch 1                           replicas (one pair per GPU)
  3  IRR_PWC, 1024x436, batch 8 PER GPU                             weak scaling (global batch 8N) + a `strong` key
  4  PWCNet_irr_occ_bi, 1024x436, GLOBAL batch 32                   sharded with shard.shard_range (32/N per GPU)
  5  IRR_PWC, KITTI shape 1242x375, GLOBAL batch 16, bf16 features  sharded (16/N per GPU)
One "step" = one full eval-mode forward of the config's model over this rank's batch.  There is no data-path
collective; the only collective is one NCCL all-gather of the per-sample EPE after the timed region.

Printed JSON (rank 0, one line): value = device-timed pairs/s with inputs resident in HBM (CUDA-graph replay of the
level loop unless --no-graph); e2e = the same through the public API (harness.PipelinedInference) with pinned HOST
inputs, H2D and D2H copies inside the timed region; roofline = the dominant correlation(+warp) launch's achieved
algorithmic HBM GB/s (per-launch CUDA events on the launching stream, an eager pass of the same steps) against
MEASURED_PEAKS.json; roofline_conv = the conv stack's achieved FLOP/s; strong = the same global batch as N=1 split over
the N ranks (config 3: B_local = 8/N) with its efficiency against this run's own one-GPU batch-8 time;
cpu_baseline = the reference forward on this box's host cores on a bounded sample (the UNMODIFIED reference classes
from baseline/_ref when staged, kind "reference"; else the bit-exact oracle port, kind "port");
torch_gpu_baseline = the reference's PyTorch-op sequence on the same GPU (cudnn.benchmark=True as main.py:73, fp32
TF32 off; plus torch's stock TF32-on setting as context).  `--impl reference` times the CPU implementation alone.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

REF_DIR = os.path.join(ROOT, "baseline", "_ref")

# BASELINE.json configs.  mode: "weak" = `batch` pairs per GPU; "sharded" = `batch` pairs globally, split over the ranks.
CONFIGS = {
    1: dict(model=None, H=64, W=128, batch=1, mode="weak", seed=0, max_flow=0.0, ckpt=None, feat="fp32",
            metric="cost volumes/sec, correlation op alone 2x(1,64,64,128)", unit="volumes/s",
            workload="correlation op alone, 2x(1,64,64,128) random tensors, BASELINE configs[0]"),
    2: dict(model="PWCNet", H=256, W=256, batch=1, mode="weak", seed=2, max_flow=12.0, ckpt="PWCNet", feat="fp32",
            metric="image-pairs/sec PWCNet 256x256 b1", unit="pairs/s",
            workload="pwcnet.py PWCNet forward (5 levels, uni-directional), 256x256 pair, batch 1, BASELINE configs[1]"),
    3: dict(model="IRR_PWC", H=436, W=1024, batch=8, mode="weak", seed=3, max_flow=20.0, ckpt="IRR-PWC_sintel",
            feat="fp32", metric="image-pairs/sec IRR-PWC 1024x436 b8", unit="pairs/s",
            workload="IRR_PWC full forward (7 pyramid levels, bi-directional flow + occlusion), 1024x436, "
                     "BASELINE configs[2]"),
    4: dict(model="PWCNet_irr_occ_bi", H=436, W=1024, batch=32, mode="sharded", seed=4, max_flow=20.0, ckpt=None,
            feat="fp32", metric="image-pairs/sec PWCNet_irr_occ_bi 1024x436 b32", unit="pairs/s",
            workload="pwcnet_irr_occ_bi forward (bi-directional + occlusion), 1024x436, global batch 32 sharded over the "
                     "GPUs, BASELINE configs[3]"),
    5: dict(model="IRR_PWC", H=375, W=1242, batch=16, mode="sharded", seed=5, max_flow=20.0, ckpt="IRR-PWC_kitti",
            feat="bf16", metric="image-pairs/sec IRR-PWC KITTI 1242x375 b16 bf16-features", unit="pairs/s",
            workload="IRR_PWC full forward, KITTI shape 1242x375, global batch 16 sharded over the GPUs, feature pyramid "
                     "in bf16, BASELINE configs[4]"),
}


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d.get("bf16_tflops", 1590.0)),
                    "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", 1400.0)), "source": "measured"}
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "200", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def cpu_threads():
    """Host threads for the CPU reference arm.  The forward is ~16 k small ATen ops: beyond ~16 threads the
    per-op fork/join cost grows faster than the work shrinks (measured on the 128-core box: 45 s/pair with 128 threads vs
    ~2.5 s/pair with 8 in the build container), so the arm uses min(cores, 16) and reports that count beside
    os.cpu_count()."""
    return max(1, min(os.cpu_count() or 1, env_int("IRR_CPU_THREADS", 16)))


def corr_bytes(B, C, H, W, s_in=4):  # B*H*W*(2*C*s_in + 81*4); fused warp adds the flow read B*H*W*8
    return B * H * W * (2 * C * s_in + 324)


def conv_flops(meta):
    B, Cin, H, W, Cout, ks, stride, dil, Ho, Wo, math = meta
    return 2.0 * B * Ho * Wo * Cout * Cin * ks * ks


# ------------------------------------------------------------------------------------------------- weights / inputs
def ref_staged():
    return os.path.isdir(os.path.join(REF_DIR, "models"))


def load_params(cfg):
    """(flat {name: tensor} parameters, description).  Trained weights (the checkpoint BASELINE.md §3 names for this
    config) when baseline/_ref is staged on the box; else the deterministic MSRA-like random init."""
    from irr_b200 import synthetic as S
    from irr_b200.checkpoint import strip_prefix
    if cfg["ckpt"]:
        path = os.path.join(REF_DIR, "saved_check_point", "pwcnet", cfg["ckpt"], "checkpoint_best.ckpt")
        if os.path.isfile(path):
            sd = torch.load(path, map_location="cpu", weights_only=True)["state_dict"]
            return strip_prefix(sd), f"trained: {cfg['ckpt']}/checkpoint_best.ckpt (reference checkpoint, staged in baseline/_ref)"
    return S.synthetic_params(cfg["model"], seed=1234, gain=0.7), "deterministic MSRA-like random init (no checkpoint for this config on the box)"


@contextlib.contextmanager
def cuda_shim():
    """The reference hard-codes .cuda() (pwc_modules.py:111,129, IRR_PWC.py:68-71): a no-op shim runs it on the host."""
    old = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = old


def reference_forward_fn(cfg, params, device):
    """A callable (img1, img2) -> output dict running the reference forward on `device`, and its kind:
    "reference" = the UNMODIFIED classes of baseline/_ref/models, "port" = oracle/irr_oracle.py (bit-exact restatement)."""
    name = cfg["model"]
    if ref_staged():
        import importlib
        sys.path.insert(0, REF_DIR)
        try:
            models = importlib.import_module("models")
        finally:
            sys.path.remove(REF_DIR)
        m = getattr(models, name)(None)
        m.load_state_dict(params)
        m = m.to(device).eval()

        def fn(a, b):
            with torch.no_grad():
                return m({"input1": a, "input2": b})
        return fn, "reference"
    from oracle import irr_oracle as O
    p = {k: v.to(device) for k, v in params.items()}

    def fn(a, b):
        with torch.no_grad():
            return O.FORWARDS[name](p, a, b)
    return fn, "port"


# ------------------------------------------------------------------------------------------------- reference arm
def cpu_forward_sample(cfg, steps, warmup, batch, threads=None, budget_s=150.0):
    """Times the reference forward on the host cores: `batch` pairs per step, stops early when `budget_s` is used up."""
    from irr_b200 import synthetic as S
    torch.set_num_threads(threads or cpu_threads())
    params, _ = load_params(cfg)
    fn, kind = reference_forward_fn(cfg, params, torch.device("cpu"))
    i1, i2, _ = S.synthetic_pair(batch, cfg["H"], cfg["W"], seed=cfg["seed"], max_flow=cfg["max_flow"])
    ts = []
    t_start = time.perf_counter()
    with cuda_shim():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fn(i1, i2)
            dt = time.perf_counter() - t0
            if i >= warmup:
                ts.append(dt)
            if ts and time.perf_counter() - t_start + dt > budget_s:
                break
    return ts, kind


def cpu_corr_sample(steps, warmup):
    """Config 1 on the host: the reference's compute_cost_volume (or its restatement) on 2 x (1,64,64,128)."""
    torch.set_num_threads(cpu_threads())
    f1 = torch.randn(1, 64, 64, 128, generator=torch.Generator().manual_seed(0))
    f2 = torch.randn(1, 64, 64, 128, generator=torch.Generator().manual_seed(1))
    kind = "port"
    if ref_staged():
        import importlib
        sys.path.insert(0, REF_DIR)
        try:
            fn = importlib.import_module("models.pwc_modules").compute_cost_volume
            kind = "reference"
        finally:
            sys.path.remove(REF_DIR)
        call = lambda: fn(f1, f2, {"max_disp": 4})
    else:
        from oracle import irr_oracle as O
        call = lambda: O.cost_volume(f1, f2)
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            call()
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    return ts, kind


def run_reference(args, cfg):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    world = env_int("WORLD_SIZE", 1)
    if args.config == 1:
        ts, kind = cpu_corr_sample(max(5, min(args.steps, 50)), 3)
        per_step, b_step = 1, 1
    else:
        # one step = this config's per-GPU batch (the same step our arm times on one GPU), capped to a time budget
        b_step = cfg["batch"] if cfg["mode"] == "weak" else max(1, cfg["batch"] // world)
        b_step = min(b_step, env_int("IRR_CPU_BATCH", 8))
        ts, kind = cpu_forward_sample(cfg, max(1, min(args.steps, 5)), 1, b_step)
        per_step = b_step
    total = sum(ts)
    value = len(ts) * per_step / total
    gb = cfg["batch"] * world if cfg["mode"] == "weak" else cfg["batch"]
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": args.gpus,
        "steps": len(ts), "warmup": 1, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
        "scaling": "weak" if cfg["mode"] == "weak" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "per_gpu_batch": cfg["batch"] if cfg["mode"] == "weak" else cfg["batch"] // world,
                   "global_batch": gb},
        "cpu_baseline": {"value": value, "unit": cfg["unit"], "cores": torch.get_num_threads(),
                         "host_cores": os.cpu_count(), "kind": kind,
                         "sample": f"{len(ts)} step(s) x {per_step} pair(s)/step of this workload on the host: " +
                                   ("the UNMODIFIED reference classes (baseline/_ref/models, .cuda() shimmed to a no-op)"
                                    if kind == "reference" else
                                    "oracle/irr_oracle.py (torch CPU ops, bit-exact restatement; reference not staged)")},
        "e2e": {"value": value, "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------- our arm
_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries print there too (NCCL's version banner under torchrun goes
    to fd 1 regardless of NCCL_DEBUG_FILE): keep a private duplicate of fd 1 for the JSON line and point fd 1 at stderr
    for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def capture(model, inp):
    """Warm up (packs weights, fills caches), count launches, capture the whole forward into one CUDA graph."""
    from irr_b200 import ops
    for _ in range(2):
        out = model(inp)
    torch.cuda.synchronize()
    ops.LAUNCHES = 0
    out = model(inp)
    launches = ops.LAUNCHES
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model(inp)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = model(inp)
    return graph, out, launches


def timed_ms(step, steps, warmup, barrier):
    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def torch_gpu_sample(cfg, params, i1, i2, dev, steps=2):
    """"also time the reference on GPU (PyTorch ops, fp32, TF32 off) as the honest kernel to beat".
    The unmodified reference classes (baseline/_ref) when staged, else the oracle's device-agnostic restatement; run
    with the reference's own cudnn.benchmark=True (main.py:73).  Two figures: fp32 with TF32 off (the parity
    configuration) and torch's stock setting (cudnn.allow_tf32=True: what a user of the reference gets by default)."""
    saved = (torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    res = {}
    try:
        a, b = i1.to(dev), i2.to(dev)
        for tag, tf32 in (("fp32", False), ("stock_tf32", True)):
            torch.backends.cudnn.benchmark = True
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            fn, kind = reference_forward_fn(cfg, params, dev)
            out = fn(a, b)  # warm-up: cuDNN autotuning of every conv shape
            out = fn(a, b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                out = fn(a, b)
            e1.record()
            torch.cuda.synchronize()
            res[tag] = (steps * a.shape[0] / (e0.elapsed_time(e1) * 1e-3), {k: v for k, v in out.items() if torch.is_tensor(v)}, kind)
            del fn
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return res


def run_config1(args, dev, rank, world, dist):
    """BASELINE configs[0]: the correlation op alone on 2 x (1,64,64,128); a step = one launch."""
    from irr_b200 import ops
    cfg = CONFIGS[1]
    f1 = torch.randn(1, 64, 64, 128, generator=torch.Generator().manual_seed(0))
    f2 = torch.randn(1, 64, 64, 128, generator=torch.Generator().manual_seed(1))
    h1, h2 = f1.pin_memory(), f2.pin_memory()
    d1, d2 = h1.to(dev), h2.to(dev)
    out = torch.empty(1, 81, 64, 128, device=dev)
    oh = torch.empty(1, 81, 64, 128).pin_memory()
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2: evicts the 2 x 2 MB inputs between launches

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    steps = max(args.steps, 20)
    for _ in range(args.warmup):
        ops.correlation(d1, d2, out=out)
    barrier()
    evs = []
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); ops.correlation(d1, d2, out=out); e.record()
        evs.append((s, e))
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(s.elapsed_time(e) for s, e in evs)

    def e2e():
        d1.copy_(h1, non_blocking=True); d2.copy_(h2, non_blocking=True)
        ops.correlation(d1, d2, out=out)
        oh.copy_(out, non_blocking=True)
    ms_e2e = timed_ms(e2e, steps, 3, barrier)
    from irr_b200.shard import max_over_ranks
    ms, ms_e2e = max_over_ranks(ms, dev), max_over_ranks(ms_e2e, dev)
    if rank != 0:
        return 0
    pk = peaks()
    by = corr_bytes(1, 64, 64, 128)
    ts, kind = cpu_corr_sample(20, 3)
    from oracle import ops_np as N
    err = float((out.cpu() - torch.from_numpy(N.cost_volume_np(f1.numpy(), f2.numpy()))).abs().max())
    emit({"metric": cfg["metric"], "value": steps * world / (ms * 1e-3), "unit": cfg["unit"], "n_gpus": world, "steps": steps,
          "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f32", "data": "synthetic",
          "config": {"workload": cfg["workload"], "per_gpu_batch": 1, "global_batch": world,
                     "l2": "160 MB flush (zero_) between timed launches; each launch timed alone with CUDA events"},
          "e2e": {"value": steps * world / (ms_e2e * 1e-3), "unit": cfg["unit"], "h2d_bytes_per_step": 2 * f1.numel() * 4,
                  "d2h_bytes_per_step": out.numel() * 4, "ms_per_step": ms_e2e / steps,
                  "api": "irr_b200.ops.correlation (== pwc_modules.compute_cost_volume / Correlation.forward), pinned host in/out"},
          "gpu_launches": steps, "launches_per_step": 1, "clocks": clocks,
          "roofline": {"bound": "hbm", "kernel": "corr_tma_kernel<fused=False> [1, 64, 64, 128]", "achieved": by / (ms / steps * 1e-3) / 1e9,
                       "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": by / (ms / steps * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                       "peak_source": pk["source"], "algo_bytes_per_launch": by,
                       "note": "6.8 MB / 32 tiles: a latency-bound launch, far from any bandwidth roof"},
          "cpu_baseline": {"value": len(ts) / sum(ts), "unit": cfg["unit"], "cores": torch.get_num_threads(),
                           "host_cores": os.cpu_count(), "kind": kind, "sample": f"{len(ts)} x compute_cost_volume on the host"},
          "parity": {"max_abs_vs_oracle": err}})
    return 0


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=env_int("IRR_BENCH_CONFIG", 3), choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--math", default=os.environ.get("IRR_MATH", "auto"), choices=["auto", "fp32", "3xtf32", "tf32", "3xf16"])
    ap.add_argument("--batch", type=int, default=0, help="override the config's batch (per GPU for weak configs, global for sharded)")
    ap.add_argument("--weights", default="auto", choices=["auto", "synthetic"], help="auto = trained checkpoint when staged")
    ap.add_argument("--serial-e2e", action="store_true", help="e2e loop without copy/compute overlap")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the PyTorch-ops-on-GPU baseline")
    ap.add_argument("--no-pruned", action="store_true", help="skip the secondary eval_prune_dead measurement")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling / batch-sweep measurement")
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch > 0:
        cfg["batch"] = args.batch
    if args.weights == "synthetic":
        cfg["ckpt"] = None
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args, cfg)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local = env_int("LOCAL_RANK", 0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if args.config == 1:
        rc = run_config1(args, dev, rank, world, dist)
        if dist is not None:
            dist.destroy_process_group()
        return rc

    import irr_b200
    from irr_b200 import ops, pwc_modules
    from irr_b200 import synthetic as S  # deterministic parameters / inputs (no oracle code on the timed path)
    from irr_b200.shard import gather_metric, max_over_ranks, shard_range

    math = {"fp32": ops.MATH_FP32_SIMT, "3xtf32": ops.MATH_TC_3XTF32, "tf32": ops.MATH_TC_TF32,
            "3xf16": ops.MATH_TC_3XF16}.get(args.math)
    if math is None:  # auto: the fp32-grade tensor-core path (3-term f16 split, TMA-staged activations)
        math = ops.MATH_TC_3XF16
    pwc_modules.set_conv_math(math)
    math_name = {0: "fp32 CUDA-core FFMA", 1: "tcgen05 3xTF32 (fp32-grade)", 2: "tcgen05 TF32",
                 3: "tcgen05 3xF16 split (fp32-grade), TMA-staged activations"}[math]

    H_IM, W_IM = cfg["H"], cfg["W"]
    if cfg["mode"] == "weak":
        B, G = cfg["batch"], cfg["batch"] * world
        lo = rank * B
    else:
        G = cfg["batch"]
        lo, hi = shard_range(G, world, rank)
        B = hi - lo
        if B == 0:
            raise SystemExit(f"bench.py: config {args.config} has global batch {G} < {world} ranks")
    params, weights_note = load_params(cfg)
    model = irr_b200.MODELS[cfg["model"]](None)
    irr_b200.load_state_dict_strict(model, params)
    model = model.to(dev).eval()
    if cfg["feat"] != "fp32":
        model.set_feature_dtype(cfg["feat"])
    has_occ = cfg["model"] != "PWCNet"
    if cfg["mode"] == "weak":   # every rank makes its own batch (seed + rank)
        i1c, i2c, gtc = S.synthetic_pair(B, H_IM, W_IM, seed=cfg["seed"] + rank, max_flow=cfg["max_flow"])
    else:   # sharded: the GLOBAL batch is the same whatever N is; every rank generates it and keeps its slice
        gi1, gi2, ggt = S.synthetic_pair(G, H_IM, W_IM, seed=cfg["seed"], max_flow=cfg["max_flow"])
        i1c, i2c, gtc = gi1[lo:lo + B].contiguous(), gi2[lo:lo + B].contiguous(), ggt[lo:lo + B].contiguous()
        del gi1, gi2, ggt
    h1, h2 = i1c.pin_memory(), i2c.pin_memory()
    d1, d2 = h1.to(dev), h2.to(dev)
    inp = {"input1": d1, "input2": d2}

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    graph = None
    if not args.no_graph:
        graph, out, launches_per_step = capture(model, inp)
    else:
        for _ in range(2):
            model(inp)
        ops.LAUNCHES = 0
        out = model(inp)
        launches_per_step = ops.LAUNCHES
    step = (lambda: graph.replay()) if graph is not None else (lambda: model(inp))

    # the working set of one step (dense buffers ~2 GB at level 4) is >> the 126 MB L2, so nothing stays cached
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed_ms(step, args.steps, 0, barrier)
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: public API (irr_b200.harness.PipelinedInference.submit), pinned host inputs -> H2D -> forward -> D2H of
    # flow (+occ), EVERY step, all inside the timed region.  The runner overlaps batch i+1's H2D and batch i-1's D2H with
    # batch i's forward (double-buffered staging, two copy streams); --serial-e2e times the strictly sequential loop.
    oh_f = [torch.empty((B, 2, H_IM, W_IM), dtype=torch.float32).pin_memory() for _ in range(2)]
    oh_o = [torch.empty((B, 1, H_IM, W_IM), dtype=torch.float32).pin_memory() for _ in range(2)] if has_occ else [None, None]
    from irr_b200.harness import PipelinedInference
    if args.serial_e2e:
        def e2e_step(i):
            d1.copy_(h1, non_blocking=True)
            d2.copy_(h2, non_blocking=True)
            if graph is not None:
                graph.replay()
                o = out
            else:
                o = model(inp)
            oh_f[0].copy_(o["flow"], non_blocking=True)
            if has_occ:
                oh_o[0].copy_(o["occ"], non_blocking=True)
        e2e_join = lambda: None
    else:
        runner = PipelinedInference(model, B, H_IM, W_IM, dev, use_graph=not args.no_graph)
        e2e_step = lambda i: runner.submit(h1, h2, oh_f[i & 1], oh_o[i & 1])
        e2e_join = runner.join

    for i in range(2):
        e2e_step(i)
    e2e_join()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        e2e_step(i)
    e2e_join()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    e2e_check = float((oh_f[(args.steps - 1) & 1] - out["flow"].cpu()).abs().max())
    if not args.serial_e2e:
        del runner
        torch.cuda.empty_cache()

    # ---- strong scaling (config 3): the N=1 global batch (8 pairs) split over the N ranks, B_local = 8/N, in the same
    # run.  Efficiency is against THIS rank's own batch-8 step (the weak run above is exactly the one-GPU workload).
    # At N=1 a batch sweep (4, 2, 1 pairs) predicts what each rank will run at N = 2, 4, 8.
    strong = None
    if args.config == 3 and not args.no_strong and graph is not None:
        def time_batch(b):
            x = {"input1": d1[:b].contiguous(), "input2": d2[:b].contiguous()}
            g_b, _, _ = capture(model, x)
            t = timed_ms(g_b.replay, args.steps, args.warmup, barrier)
            del g_b
            return t / args.steps
        if world == 1:
            sweep = {b: time_batch(b) for b in (4, 2, 1) if b < B}
            strong = {"global_batch": B, "per_gpu_batch": B, "value": B * args.steps / (ms * 1e-3), "unit": "pairs/s",
                      "efficiency_vs_n1": 1.0,
                      "batch_sweep_ms": {str(B): ms / args.steps, **{str(b): t for b, t in sweep.items()}},
                      "predicted_efficiency": {str(B // b): (ms / args.steps) / ((B // b) * t) for b, t in sweep.items()},
                      "limit": "levels 0-2 are launch/latency-bound (fixed cost per step independent of the batch)"}
        elif B % world == 0:
            bl = B // world
            t_loc = time_batch(bl)
            t_max = max_over_ranks(t_loc, dev)
            t_n1 = max_over_ranks(ms / args.steps, dev)
            strong = {"global_batch": B, "per_gpu_batch": bl, "value": B / (t_max * 1e-3), "unit": "pairs/s",
                      "ms_per_step": t_max, "n1_ms_per_step": t_n1, "efficiency_vs_n1": t_n1 / (world * t_max),
                      "limit": "levels 0-2 are launch/latency-bound (fixed cost per step independent of the batch)"}

    # ---- secondary figure (never the headline): the same forward with the eval-dead backward-occlusion chain pruned
    # (IRR_PWC.eval_prune_dead): fewer layers, same outputs.  `value` above is the FULL forward.
    pruned = None
    if not args.no_pruned and graph is not None and hasattr(model, "eval_prune_dead"):
        model.eval_prune_dead = True
        try:
            g2, out2, _ = capture(model, inp)
            ms_p = timed_ms(g2.replay, args.steps, args.warmup, barrier)
            pruned = {"ms": ms_p, "max_abs_flow_vs_full": float((out2["flow"] - out["flow"]).abs().max()),
                      "max_abs_occ_vs_full": float((out2["occ"] - out["occ"]).abs().max())}
            del g2, out2
        finally:
            model.eval_prune_dead = False

    # ---- per-kernel timing pass (eager, CUDA events on the launching stream around every launch)
    # (single stream: with the flow / occlusion branches overlapped on two streams per-launch times are not additive)
    _m = sys.modules.get("irr_b200.IRR_PWC")
    _m.set_side_stream(False)
    ops.TIMING = []
    nprof = min(args.steps, 3)
    for _ in range(nprof):
        model(inp)
    torch.cuda.synchronize()
    timing, ops.TIMING = ops.TIMING, None
    _m.set_side_stream(True)
    agg = {}
    for what, meta, s, e in timing:
        agg.setdefault((what, meta), []).append(s.elapsed_time(e))
    total_kernel_ms = sum(sum(v) for v in agg.values()) / nprof
    if rank == 0 and os.environ.get("IRR_DUMP_TIMES"):
        rows = [{"kernel": k[0], "meta": list(k[1]) if k[1] is not None else None, "launches_per_step": len(v) / nprof,
                 "ms_per_step": sum(v) / nprof, "mean_us": 1e3 * sum(v) / len(v)} for k, v in agg.items()]
        rows.sort(key=lambda r: -r["ms_per_step"])
        json.dump(rows, open(os.environ["IRR_DUMP_TIMES"], "w"), indent=0)

    # ---- metric reduction (the only collective): per-sample EPE vs the synthetic ground truth, all-gathered
    epe_local = torch.norm(out["flow"] - gtc.to(dev), p=2, dim=1).mean(dim=(1, 2))
    epe_all = gather_metric(epe_local, G) if cfg["mode"] == "sharded" else gather_metric(epe_local, B * world)
    ms, ms_e2e = max_over_ranks(ms, dev), max_over_ranks(ms_e2e, dev)
    if pruned is not None:
        pruned["ms"] = max_over_ranks(pruned["ms"], dev)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    pk = peaks()
    pairs = G * args.steps
    value = pairs / (ms * 1e-3)
    # bytes per f1 / f2 element the cost volumes read: 2 when the model feeds them from packed bf16 storage (config 5)
    s_in = 2 if (cfg["feat"] == "bf16" and getattr(model, "bf16_storage", False)) else 4
    # dominant correlation launch = the finest-level call (largest algorithmic bytes)
    corr = [(k, v) for k, v in agg.items() if k[0] in ("correlation", "warp_correlation")]
    levels = []
    for (what, meta), v in sorted(corr, key=lambda kv: -corr_bytes(*kv[0][1])):
        by = corr_bytes(*meta, s_in=s_in) + (meta[0] * meta[2] * meta[3] * 8 if what == "warp_correlation" else 0)
        t = statistics.mean(v)
        levels.append({"kernel": what, "B_C_H_W": list(meta), "ms": t, "algo_bytes": by, "GBps": by / (t * 1e-3) / 1e9,
                       "gflops": 2.0 * 81 * meta[0] * meta[1] * meta[2] * meta[3] / (t * 1e-3) / 1e9})
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "corr_traffic.json")
    if os.path.exists(tpath) and args.config == 3:
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roof = None
    if levels:
        top = levels[0]
        roof = {"bound": "hbm", "kernel": f"corr_tma_kernel<fused={top['kernel'] == 'warp_correlation'}> {top['B_C_H_W']}",
                "achieved": top["GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": top["GBps"] / pk["hbm_gbs"],
                "traffic": traffic, "peak_source": pk["source"], "ms_per_launch": top["ms"],
                "algo_bytes_per_launch": top["algo_bytes"],
                "share_of_step": sum(statistics.mean(v) for _, v in corr) / total_kernel_ms}
    convs = [(k, v) for k, v in agg.items() if k[0] == "conv2d"]
    conv_ms = sum(sum(v) for _, v in convs) / nprof
    conv_fl = sum(conv_flops(k[1]) * len(v) for k, v in convs) / nprof
    # fp32-grade math: peak for the tensor path is the measured bf16 GEMM rate / 2 (tf32) / 3 (three passes)
    if math == ops.MATH_FP32_SIMT:
        cpeak, cnote = 2 * 128 * 148 * 1.965e9 / 1e12, "nominal fp32 FFMA peak 148 SM x 128 lanes x 2 x 1.965 GHz"
    else:
        div = {ops.MATH_TC_3XTF32: 6.0, ops.MATH_TC_TF32: 2.0, ops.MATH_TC_3XF16: 3.0}[math]
        cpeak, cnote = pk["bf16_tflops_sustained"] / div, f"measured bf16 GEMM (sustained) / {div:g}"
    roof_conv = {"bound": "tensor" if math != ops.MATH_FP32_SIMT else "fp32-simt", "achieved": conv_fl / (conv_ms * 1e-3) / 1e12,
                 "peak": cpeak, "unit": "TFLOP/s", "frac": conv_fl / (conv_ms * 1e-3) / 1e12 / cpeak, "peak_note": cnote,
                 "share_of_step": conv_ms / total_kernel_ms, "flop_per_step": conv_fl, "math": math_name}

    # ---- baselines, after the timed regions: reference op sequence on this GPU, then on the host cores
    tgpu = None
    if not args.no_torch_gpu:
        try:
            nb = min(B, 8)
            r = torch_gpu_sample(cfg, params, i1c[:nb], i2c[:nb], dev)
            ops.set_grid_mode(ops.GRID_RECIP_MUL)  # torch-CUDA scalar-division arithmetic, so the hard masks can match
            try:
                mine = model({"input1": d1[:nb].contiguous(), "input2": d2[:nb].contiguous()})
            finally:
                ops.set_grid_mode(ops.GRID_TRUE_DIV)
            v32, o32, kind = r["fp32"]
            vtf, otf, _ = r["stock_tf32"]
            tgpu = {"value": v32, "unit": "pairs/s", "kind": kind,
                    "what": "the reference's PyTorch-op sequence on this GPU (ATen / cuDNN, eager, cudnn.benchmark=True as "
                            "main.py:73, fp32 with TF32 off): " + ("UNMODIFIED reference classes from baseline/_ref"
                                                                    if kind == "reference" else "oracle restatement"),
                    "sample": f"2 x batch {nb}",
                    "stock_tf32_value": vtf,
                    "epe_ours_vs_torch_gpu": float(S.epe(mine["flow"], o32["flow"])),
                    "max_abs_flow_ours_vs_torch_gpu": float((mine["flow"] - o32["flow"]).abs().max()),
                    "epe_stock_tf32_vs_fp32_reference": float(S.epe(otf["flow"], o32["flow"]))}
            del r, o32, otf, mine
        except Exception as ex:  # a reported baseline must never take the bench line down
            tgpu = {"unavailable": repr(ex)[:300]}
        torch.cuda.empty_cache()
    cpu = None
    if args.cpu_baseline_steps > 0:
        try:
            ts, kind = cpu_forward_sample(cfg, args.cpu_baseline_steps, 1, 1, budget_s=45.0)
            cpu = {"value": len(ts) / sum(ts), "unit": "pairs/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(),
                   "kind": kind, "sample": f"{len(ts)} x one {W_IM}x{H_IM} pair through " +
                   ("the UNMODIFIED reference classes (baseline/_ref/models) " if kind == "reference" else "oracle/irr_oracle.py ") +
                   f"on the host, {sum(ts):.1f} s"}
        except Exception as ex:
            cpu = {"unavailable": repr(ex)[:300]}

    line = {
        "metric": cfg["metric"], "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak" if cfg["mode"] == "weak" else "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "per_gpu_batch": B, "global_batch": G,
                   "parallelism": f"batch-sharded replicas x{world} (no data-path collective)",
                   "conv_math": math_name, "features": cfg["feat"] + (" (values in the fp32 layout for convs / warps; packed bf16 storage for the cost volumes' f1 / f2)"
                                             if s_in == 2 else ""), "cuda_graph": graph is not None,
                   "l2": "per-step working set (GBs of level-4 activations) exceeds the 126 MB L2; no explicit flush",
                   "weights": weights_note},
        "e2e": {"value": pairs / (ms_e2e * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": 2 * h1.numel() * 4,
                "d2h_bytes_per_step": (oh_f[0].numel() + (oh_o[0].numel() if has_occ else 0)) * 4,
                "ms_per_step": ms_e2e / args.steps,
                "api": "strictly sequential H2D -> forward -> D2H per step" if args.serial_e2e else
                       "irr_b200.harness.PipelinedInference.submit (H2D of step i+1 / D2H of step i-1 overlap step i)",
                "host_output_max_abs_vs_device": e2e_check},
        "gpu_launches": launches_per_step * args.steps,
        "launches_per_step": launches_per_step,
        "clocks": clocks,
        "roofline": roof,
        "roofline_corr_levels": levels[:6],
        "roofline_conv": roof_conv,
        "strong": strong,
        "cpu_baseline": cpu,
        "torch_gpu_baseline": tgpu,
        "eval_pruned": None if pruned is None else {
            "value": pairs / (pruned["ms"] * 1e-3), "unit": "pairs/s", "ms_per_step": pruned["ms"] / args.steps,
            "max_abs_flow_vs_full": pruned["max_abs_flow_vs_full"], "max_abs_occ_vs_full": pruned["max_abs_occ_vs_full"],
            "note": "SECONDARY, not the headline: the eval-dead backward occlusion chain pruned "
                    "(IRR_PWC.eval_prune_dead); `value` is the full reference-equivalent forward"},
        "metric_reduction": {"epe_vs_synthetic_gt_mean": float(epe_all.mean()), "samples": int(epe_all.numel()),
                             "collective": "nccl all_gather" if dist is not None else "none (1 GPU)"},
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
