"""CPU tests of the N>1 host logic with a real 2-process gloo group."""
import os

import pytest
import torch
import torch.multiprocessing as mp


def test_shard_range_partitions():
    from irr_b200.shard import shard_range
    for gb in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [shard_range(gb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from irr_b200.shard import gather_metric, max_over_ranks, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gb = 7  # ragged on purpose
    a, b = shard_range(gb, world, rank)
    local = torch.arange(a, b, dtype=torch.float32) * 10.0
    full = gather_metric(local, gb)
    t = max_over_ranks(1.0 + rank, torch.device("cpu"))
    q.put((rank, full.tolist(), t))
    dist.destroy_process_group()


def test_gather_metric_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, t in res:
        assert full == [float(i) * 10.0 for i in range(7)]
        assert t == 2.0


def _run_bench(*extra, env=None):
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), *extra], capture_output=True, text=True, timeout=600,
                       env=e)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1   # the driver parses stdout: one JSON line, nothing else (library chatter goes to stderr)
    return json.loads(lines[0])


def test_bench_reference_arm_prints_exactly_one_json_line():
    """`bench.py --impl reference`: the reference's CPU implementation of the default workload (config 3), one JSON line.
    (IRR_CPU_BATCH=1 bounds the sample to one pair per step so the CPU suite stays short.)"""
    d = _run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", env={"IRR_CPU_BATCH": "1"})
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["host_cores"] >= d["cpu_baseline"]["cores"] >= 1
    assert "BASELINE configs[2]" in d["config"]["workload"] and d["config"]["per_gpu_batch"] == 8


def test_bench_reference_arm_other_configs():
    """--config 1 (correlation op alone) and --config 2 (PWCNet 256x256 b1) name their workload and metric."""
    d1 = _run_bench("--impl", "reference", "--config", "1", "--steps", "5", "--warmup", "1")
    assert "BASELINE configs[0]" in d1["config"]["workload"] and d1["unit"] == "volumes/s" and d1["value"] > 0
    d2 = _run_bench("--impl", "reference", "--config", "2", "--steps", "1", "--warmup", "1")
    assert "BASELINE configs[1]" in d2["config"]["workload"] and d2["metric"].startswith("image-pairs/sec PWCNet")


def test_bench_sharded_configs_split_the_global_batch():
    """Configs 4 and 5 shard a GLOBAL batch (32 over 8 GPUs, 16 over 4) with shard.shard_range: 4 pairs per rank."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from irr_b200.shard import shard_range
    for cfg, world in ((4, 8), (5, 4)):
        c = bench.CONFIGS[cfg]
        assert c["mode"] == "sharded"
        sizes = [shard_range(c["batch"], world, r) for r in range(world)]
        assert [b - a for a, b in sizes] == [4] * world and sizes[0][0] == 0 and sizes[-1][1] == c["batch"]
    assert bench.CONFIGS[3]["mode"] == "weak" and bench.CONFIGS[3]["batch"] == 8
