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


def test_bench_reference_arm_prints_exactly_one_json_line():
    """The driver parses bench.py's stdout: one JSON line, nothing else (library chatter is routed to stderr)."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
