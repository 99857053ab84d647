import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Parity tolerances (BASELINE.json north_star: 1e-4 eV, 1e-4 eV/A; charges as the reference's CHARGE_ATOL
# tests/conftest.py:162-165)
ENERGY_ATOL = 1e-4
FORCE_ATOL = 1e-4
CHARGE_ATOL = 1e-4


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    inputs = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    ref = {k[4:]: z[k] for k in z.files if k.startswith("ref_")}
    meta = {k: z[k].item() for k in ("weights_seed", "weights_scale", "num_charge_channels", "weights_checksum")}
    return inputs, ref, meta


_SD_CACHE = {}


def golden_state_dict(meta):
    from aimnetcentral_b200.model_spec import ModelSpec, random_state_dict

    key = (meta["weights_seed"], meta["num_charge_channels"], meta["weights_scale"])
    if key not in _SD_CACHE:
        spec = ModelSpec(num_charge_channels=int(meta["num_charge_channels"]))
        sd = random_state_dict(int(meta["weights_seed"]), spec, scale=float(meta["weights_scale"]))
        tot = 0.0
        import torch

        for k in sorted(sd):
            tot += float(torch.nan_to_num(sd[k].double(), nan=0.0).abs().sum())
        assert abs(tot - meta["weights_checksum"]) < 1e-6 * abs(tot), "random_state_dict drifted from the golden weights"
        _SD_CACHE[key] = (sd, spec)
    return _SD_CACHE[key]


@pytest.fixture(scope="session")
def has_cuda():
    import torch

    return torch.cuda.is_available()
