"""Loading the reference's checkpoints into the irr_b200 model classes.

The reference saves ``{'epe':…, 'F1':…, 'epoch':…, 'state_dict': ModelAndLoss.state_dict()}`` (configuration.py:290-300);
every key carries the ``_model.`` prefix because the saved module is ``ModelAndLoss`` (configuration.py:23,291).
The files hold only tensors, floats and ints, so they are read with ``weights_only=True`` (no arbitrary pickle code
is executed — checkpoints are usually downloaded); ``trust=True`` is the explicit opt-in for a file that needs the
full unpickler."""
from __future__ import annotations

import torch

PREFIX = "_model."


def strip_prefix(state_dict):
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in state_dict.items()}


def load_state_dict_strict(model, state_dict):
    """Strict load (names AND shapes must match, like configuration.py:211-233 with the default include-all filter)."""
    sd = strip_prefix(state_dict)
    own = model.state_dict()
    missing = sorted(set(own) - set(sd))
    extra = sorted(set(sd) - set(own))
    if missing or extra:
        raise RuntimeError(f"checkpoint mismatch: missing {missing[:5]}… extra {extra[:5]}…")
    for k, v in sd.items():
        if tuple(own[k].shape) != tuple(v.shape):
            raise RuntimeError(f"checkpoint shape mismatch for {k}: {tuple(v.shape)} vs {tuple(own[k].shape)}")
    model.load_state_dict(sd)
    return model


def load_reference_checkpoint(model, path, map_location="cpu", trust: bool = False):
    """Returns the stats dict stored beside the weights (epe / F1 / outlier / epoch).  ``trust=True`` unpickles with
    ``weights_only=False`` (runs arbitrary code from the file: only for files you made yourself)."""
    ck = torch.load(path, map_location=map_location, weights_only=not trust)
    load_state_dict_strict(model, ck["state_dict"])
    return {k: v for k, v in ck.items() if k != "state_dict"}
