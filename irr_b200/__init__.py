"""irr_b200 — B200-native (sm_100a) implementation of the IRR-PWC dense-inference hot path.

Public surface mirrors the reference's ``models`` package for that path (models/__init__.py:19-35):
``irr_b200.IRR_PWC`` / ``PWCNet`` / ``PWCNet_irr_occ_bi`` and the six other ``PWCNet_*`` ablation classes with the reference constructor
``(args, div_flow=0.05)`` and ``forward(input_dict)`` contract, the building blocks of ``pwc_modules`` /
``irr_modules``, the ``Correlation`` module, and ``install()`` to swap the kernels in under the reference's own
model classes.  All arithmetic runs in ``libirr_b200.so`` (C ABI: include/irr_b200.h); importing this package
without that library built raises on first use — there is no fallback path.
"""
from . import ops  # noqa: F401
from . import pwc_modules, irr_modules  # noqa: F401
from .correlation import Correlation  # noqa: F401
from .IRR_PWC import PWCNet as IRR_PWC  # noqa: F401
from .pwcnet import PWCNet as PWCNet  # noqa: F401
from .pwcnet_irr_occ_bi import PWCNet as PWCNet_irr_occ_bi  # noqa: F401
from .pwcnet_bi import PWCNet as PWCNet_bi  # noqa: F401
from .pwcnet_occ import PWCNet as PWCNet_occ  # noqa: F401
from .pwcnet_occ_bi import PWCNet as PWCNet_occ_bi  # noqa: F401
from .pwcnet_irr import PWCNet as PWCNet_irr  # noqa: F401
from .pwcnet_irr_bi import PWCNet as PWCNet_irr_bi  # noqa: F401
from .pwcnet_irr_occ import PWCNet as PWCNet_irr_occ  # noqa: F401
from .checkpoint import load_reference_checkpoint, load_state_dict_strict  # noqa: F401
from .install import install, patch_instances, uninstall  # noqa: F401

MODELS = {"IRR_PWC": IRR_PWC, "PWCNet": PWCNet, "PWCNet_irr_occ_bi": PWCNet_irr_occ_bi, "PWCNet_bi": PWCNet_bi,
          "PWCNet_occ": PWCNet_occ, "PWCNet_occ_bi": PWCNet_occ_bi, "PWCNet_irr": PWCNet_irr,
          "PWCNet_irr_bi": PWCNet_irr_bi, "PWCNet_irr_occ": PWCNet_irr_occ}
