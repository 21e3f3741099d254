"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol include/irr_b200.h
declares, the ctypes prototype table matches the header's argument counts, and the host mirror keeps the reference's
names / constructor signatures / parameter names."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "irr_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    fns = {}
    for m in re.finditer(r"\b(?:int|size_t|const char\*)\s+(irr_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        fns[m.group(1)] = n
    return fns


def test_library_loads_and_exports_every_declared_symbol():
    from irr_b200 import _lib
    lib = _lib.load()
    fns = header_functions()
    assert len(fns) >= 18
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert lib.irr_abi_version() == 2


def test_ctypes_table_matches_header():
    from irr_b200 import _lib
    fns = header_functions()
    assert set(fns) == set(_lib.PROTOTYPES), set(fns) ^ set(_lib.PROTOTYPES)
    for name, n in fns.items():
        assert len(_lib.PROTOTYPES[name]) == n, name


def test_argument_errors_do_not_need_a_gpu():
    """Argument validation happens before any CUDA call, returns IRR_E_ARG and sets irr_last_error()."""
    from irr_b200 import _lib
    lib = _lib.load()
    rc = lib.irr_correlation_fwd(None, 0, None, 0, None, 0, 1, 1, 1, 1, 4, 0, 1.0, None)
    assert rc == -1 and b"null pointer" in lib.irr_last_error()
    rc = lib.irr_correlation_fwd(16, 0, 16, 0, 16, 0, 1, 8, 8, 8, 3, 0, 1.0, None)
    assert rc == -1 and b"max_disp" in lib.irr_last_error()
    assert lib.irr_conv2d_packed_bytes(128, 115, 3, 0) == ((115 * 9 + 15) // 16 * 16) * 128 * 4
    assert lib.irr_conv2d_packed_bytes(128, 115, 5, 0) == 0
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.irr_correlation_generic_out_shape(64, 128, 4, 1, 4, 1, 1, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (81, 64, 128)   # correlation_cuda.cc:23-32
    assert lib.irr_correlation_generic_out_shape(64, 128, 20, 1, 20, 1, 2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)) == 0
    assert (oc.value, oh.value, ow.value) == (441, 64, 128)  # FlowNetC parameters


def test_host_raises_without_cuda_tensors():
    """No CPU fallback: CPU tensors are rejected loudly."""
    from irr_b200 import ops
    with pytest.raises(RuntimeError):
        ops.correlation(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8))


def test_missing_library_fails_loudly(monkeypatch):
    from irr_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libirr_b200.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _lib.load()


def test_model_classes_keep_reference_contract():
    import irr_b200
    from oracle import irr_oracle as O
    for name, cls in irr_b200.MODELS.items():
        sig = inspect.signature(cls.__init__)
        assert list(sig.parameters)[1:] == ["args", "div_flow"] and sig.parameters["div_flow"].default == 0.05
        m = cls(None)
        own = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert own == {k: tuple(v) for k, v in O.param_shapes(name).items()}
        # '_model.'-prefixed checkpoints (configuration.py:23,291) load strictly
        sd = {"_model." + k: v for k, v in O.synthetic_params(name).items()}
        irr_b200.load_state_dict_strict(m, sd)
        with pytest.raises(RuntimeError):
            bad = dict(sd); bad.pop(next(iter(bad)))
            irr_b200.load_state_dict_strict(m, bad)
    from irr_b200 import pwc_modules as P, irr_modules as I
    assert list(inspect.signature(P.WarpingLayer.forward).parameters) == ["self", "x", "flow", "height_im", "width_im", "div_flow"]
    assert list(inspect.signature(P.compute_cost_volume).parameters) == ["feat1", "feat2", "param_dict"]
    assert list(inspect.signature(I.RefineFlow.forward).parameters) == ["self", "flow", "diff_img", "feature"]
    assert list(inspect.signature(I.RefineOcc.forward).parameters) == ["self", "occ", "feat1", "feat2"]
    assert list(inspect.signature(I.OccUpsampleNetwork.forward).parameters) == ["self", "occ", "x"]
    c = irr_b200.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    assert (c.pad_size, c.kernel_size, c.max_displacement, c.stride1, c.stride2, c.corr_multiply) == (4, 1, 4, 1, 1, 1)


@pytest.mark.skipif(not os.path.isdir("/root/reference/saved_check_point"), reason="reference checkpoints not mounted")
def test_reference_checkpoints_load():
    import irr_b200
    base = "/root/reference/saved_check_point/pwcnet"
    stats = irr_b200.load_reference_checkpoint(irr_b200.IRR_PWC(None), f"{base}/IRR-PWC_sintel/checkpoint_best.ckpt")
    assert abs(stats["epe"] - 2.7811) < 1e-3
    irr_b200.load_reference_checkpoint(irr_b200.PWCNet(None), f"{base}/PWCNet/checkpoint_best.ckpt")
    irr_b200.load_reference_checkpoint(irr_b200.IRR_PWC(None), f"{base}/IRR-PWC_kitti/checkpoint_best.ckpt")
