"""Run the UNMODIFIED reference (/root/reference) on CPU.  TEST INFRASTRUCTURE ONLY; works only in the build container
(nothing in `-m gpu` tests or smoke() imports this; `bench.py --impl reference` / `cpu_baseline` use it on the GPU box
through the copy that oracle/make_ref.py vendors into the git-ignored oracle/_ref/).

Recipe: put import stubs for the two absent third-party packages (`warp`,
`nvalchemiops`) ahead of the reference on sys.path, build the production graph from the reference's own YAML
(aimnet/models/aimnet2_dftd3_wb97m.yaml) through its own builder (aimnet/config.py:154 build_module,
aimnet/models/utils.py:379 strip_lr_modules_from_yaml), load OUR seeded state_dict into it, and wrap it in the
reference's AIMNet2Calculator(device="cpu", deterministic=True) so DSF / DFT-D3 go through the reference's in-tree
pure-torch paths (calculator.py:971-975, 1002-1008).  The only non-reference arithmetic is the brute-force
neighbor list (oracle/nblist_oracle.py) standing in for nvalchemiops.neighbor_list.
"""
from __future__ import annotations

import copy
import os
import sys
import warnings

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def _find_reference_root() -> str:
    """The reference tree itself (build container), else the byte-for-byte copy vendored by oracle/make_ref.py into
    the git-ignored oracle/_ref/ (what travels to the GPU box)."""
    env = os.environ.get("AIMNET_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/aimnet"):
        return "/root/reference"
    return os.path.join(_HERE, "_ref")


REF_ROOT = _find_reference_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "aimnet"))


def _bootstrap():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    for p in (REF_ROOT, os.path.join(_HERE, "refstubs"), _REPO):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, os.path.join(_HERE, "refstubs"))
    sys.path.insert(0, _REPO)


def build_reference_calculator(state_dict, spec=None, deterministic: bool = True, double: bool = False):
    """Reference AIMNet2Calculator on CPU holding `state_dict` (keys as in aimnetcentral_b200.model_spec)."""
    _bootstrap()
    import torch
    import yaml

    from aimnetcentral_b200.model_spec import ModelSpec

    spec = spec or ModelSpec()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from aimnet.calculators import AIMNet2Calculator
        from aimnet.config import build_module
        from aimnet.models.utils import convert_atomic_shifts_to_float64, strip_lr_modules_from_yaml

    with open(os.path.join(REF_ROOT, "aimnet", "models", "aimnet2_dftd3_wb97m.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg = copy.deepcopy(cfg)
    cfg["kwargs"]["num_charge_channels"] = spec.num_charge_channels
    core = strip_lr_modules_from_yaml(cfg, {})[0]
    model = build_module(copy.deepcopy(core))
    convert_atomic_shifts_to_float64(model)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("outputs.srcoulomb") for k in missing), missing
    if double:
        model = model.double()
    model.__dict__["_metadata"] = spec.metadata()
    calc = AIMNet2Calculator(model, device="cpu", deterministic=deterministic)
    return calc


def run_reference(calc, data: dict, forces=True, stress=False):
    import torch

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = calc(dict(data), forces=forces, stress=stress)
    return {k: v.detach().cpu().numpy() for k, v in out.items() if isinstance(v, torch.Tensor)}
