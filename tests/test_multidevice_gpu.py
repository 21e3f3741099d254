"""One process driving two GPUs (the per-device caches of cudaFuncAttributeMaxDynamicSharedMemorySize, SM counts
and side streams): the kernels with more than 48 KB of dynamic shared memory — TMA correlation, tcgen05 convs — must work on
cuda:1 after cuda:0 has run in the same process, and a model placed on cuda:1 must work while cuda:0 is the current device.
Skipped on a single-GPU box."""
import numpy as np
import pytest
import torch

from oracle import irr_oracle as O

pytestmark = pytest.mark.gpu


def _ops_on(dev):
    from irr_b200 import ops
    g = torch.Generator().manual_seed(5)
    f = torch.randn(2, 32, 28, 64, generator=g)
    flow = torch.randn(2, 2, 28, 64, generator=g) * 0.3
    x = torch.randn(2, 115, 28, 64, generator=g)
    w = torch.randn(128, 115, 3, 3, generator=g) * 0.03
    b = torch.randn(128, generator=g) * 0.1
    xr = torch.randn(2, 32, 20, 132, generator=g)
    wr = torch.randn(32, 32, 3, 3, generator=g) * 0.06
    br = torch.randn(32, generator=g) * 0.1
    with torch.cuda.device(dev):
        fd = f.to(dev)
        a = ops.correlation(fd, fd, shift=1, slope=0.1)
        c = ops.warp_correlation(fd, fd, flow.to(dev), 448, 1024, 0.05, shift=1, slope=0.1)
        y = ops.conv2d(x.to(dev), ops.pack_weights(w.to(dev), ops.MATH_TC_3XF16), b.to(dev), 128, 3, slope=0.1,
                       math=ops.MATH_TC_3XF16)
        r = ops.conv2d(xr.to(dev), ops.pack_weights(wr.to(dev), ops.MATH_TC_3XF16), br.to(dev), 32, 3, slope=0.1,
                       math=ops.MATH_TC_3XF16)
        torch.cuda.synchronize(dev)
    return [t.cpu() for t in (a, c, y, r)]


def test_two_devices_in_one_process(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    d0, d1 = torch.device("cuda:0"), torch.device("cuda:1")
    r0 = _ops_on(d0)
    r1 = _ops_on(d1)       # first use of every kernel on device 1, after device 0 raised its attributes
    r0b = _ops_on(d0)
    for a, b, c in zip(r0, r1, r0b):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_model_on_second_device_while_first_is_current(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import irr_b200
    from irr_b200 import ops
    p = O.synthetic_params("IRR_PWC", seed=1234, gain=0.7)
    i1, i2, _ = O.synthetic_pair(1, 64, 96, seed=21, max_flow=4.0)
    outs = []
    torch.cuda.set_device(0)
    for dev in ("cuda:0", "cuda:1"):
        m = irr_b200.IRR_PWC(None)
        irr_b200.load_state_dict_strict(m, p)
        m = m.to(dev).eval()
        outs.append({k: v.cpu() for k, v in m({"input1": i1.to(dev), "input2": i2.to(dev)}).items()})   # current device stays 0
    assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0]["flow"], outs[1]["flow"]) and torch.equal(outs[0]["occ"], outs[1]["occ"])
    # an op called directly on a tensor of the non-current device says so instead of faulting
    x = torch.randn(1, 8, 9, 16, device="cuda:1")
    with pytest.raises(RuntimeError, match="current device"):
        ops.correlation(x, x)
