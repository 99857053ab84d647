"""Architecture constants of the AIMNet2 graph and a seeded random `state_dict` generator.

The shipped reference YAMLs (aimnet/models/aimnet2.yaml, aimnet2_dftd3_wb97m.yaml) fully define the graph; real
checkpoints are downloaded from GCS and are not available offline (SURVEY.md §8c "Weights"), so parity and benchmarks
run on seeded random weights with exactly the reference's `state_dict` key names and shapes
(aimnet/models/aimnet2.py:24-84, aimnet/modules/core.py:11-46, aimnet/modules/aev.py:66-81).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import torch

NFEATURE = 16  # A  (aimnet2.yaml: nfeature)
NSHIFTS = 16  # G  (aimnet2.yaml: aev.nshifts_s)
NCOMB_V = 12  # H  (aimnet2.yaml: ncomb_v)
AIM_SIZE = 256
RC_S = 5.0
RMIN = 0.8
HIDDEN = ([512, 380], [512, 380], [512, 380, 380])
HEAD_HIDDEN = [128, 128]
NUM_EMBED = 64

Hartree = 27.211386024367243  # aimnet/constants.py:6
Bohr = 0.5291772105638411  # aimnet/constants.py:8

# the species the public aimnet2 wB97M-D3 models implement (docs/models); used for random-weight embeddings
DEFAULT_SPECIES = (1, 5, 6, 7, 8, 9, 14, 15, 16, 17, 33, 34, 35, 53)


@dataclass
class ModelSpec:
    num_charge_channels: int = 1
    cutoff: float = RC_S
    coulomb_sr_rc: float = 4.6
    coulomb_sr_envelope: str = "exp"
    d3_params: dict = field(default_factory=lambda: {"s8": 0.3908, "a1": 0.5660, "a2": 3.1280, "s6": 1.0})
    implemented_species: tuple = DEFAULT_SPECIES

    @property
    def C(self) -> int:
        return self.num_charge_channels

    @property
    def conv_a_size(self) -> int:
        return NFEATURE * NSHIFTS + NFEATURE * NCOMB_V  # 448

    @property
    def conv_q_size(self) -> int:
        return self.C * (NSHIFTS + NCOMB_V)  # 28 C

    def mlp_sizes(self) -> list[list[int]]:
        """[n_in, hidden..., n_out] per pass (aimnet/models/aimnet2.py:57-84)."""
        nf = NFEATURE * NSHIFTS
        n0 = self.conv_a_size + nf
        n1 = n0 + self.conv_q_size + self.C
        out = nf + 2 * self.C
        return [[n0, *HIDDEN[0], out], [n1, *HIDDEN[1], out], [n1, *HIDDEN[2], AIM_SIZE]]

    def mlp_final_act(self) -> list[bool]:
        """Whether the last Linear of each pass MLP is followed by GELU (last_linear flag, aimnet2.py:56,65,75)."""
        return [False, True, True]

    def metadata(self) -> dict:
        return {
            "format_version": 2,
            "cutoff": self.cutoff,
            "needs_coulomb": True,
            "needs_dispersion": True,
            "coulomb_mode": "sr_embedded",
            "coulomb_sr_rc": self.coulomb_sr_rc,
            "coulomb_sr_envelope": self.coulomb_sr_envelope,
            "d3_params": dict(self.d3_params),
            "has_embedded_lr": True,
            "implemented_species": list(self.implemented_species),
        }


def aev_constants() -> dict[str, torch.Tensor]:
    """AEV buffers exactly as AEVSV._init_basis builds them (aimnet/modules/aev.py:66-81)."""
    eta = (1 / ((RC_S - RMIN) / NSHIFTS)) ** 2
    shifts = torch.linspace(RMIN, RC_S, NSHIFTS + 1)[:NSHIFTS]
    out = {}
    for mod in ("_s", "_v"):
        out[f"aev.rc{mod}"] = torch.tensor(RC_S, dtype=torch.float)
        out[f"aev.eta{mod}"] = torch.tensor(eta, dtype=torch.float)
        out[f"aev.shifts{mod}"] = shifts.clone()
    return out


def random_state_dict(seed: int = 0, spec: ModelSpec | None = None, scale: float = 0.5) -> dict[str, torch.Tensor]:
    """Seeded random weights with the reference's key names/shapes/dtypes.

    Biases and the fp64 atomic shifts are randomised too so every term of the path is exercised; `aev.*` keeps the
    reference's deterministic values (SURVEY.md Appendix A pitfall).  numpy's PCG64 stream makes it reproducible
    across machines.
    """
    spec = spec or ModelSpec()
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: dict[str, torch.Tensor] = dict(aev_constants())

    def normal(shape, std):
        return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32))

    afv = torch.full((NUM_EMBED, NFEATURE * NSHIFTS), float("nan"), dtype=torch.float)
    afv[0] = 0.0
    for z in spec.implemented_species:
        afv[z] = normal((NFEATURE * NSHIFTS,), 0.5)
    sd["afv.weight"] = afv
    sd["conv_a.agh"] = normal((NFEATURE, NSHIFTS, NCOMB_V), 0.5)
    sd["conv_q.agh"] = normal((spec.C, NSHIFTS, NCOMB_V), 0.5)
    for p, sizes in enumerate(spec.mlp_sizes()):
        for li in range(len(sizes) - 1):
            n_in, n_out = sizes[li], sizes[li + 1]
            std = scale * math.sqrt(2.0 / (n_in + n_out)) * 1.5
            sd[f"mlps.{p}.{2 * li}.weight"] = normal((n_out, n_in), std)
            sd[f"mlps.{p}.{2 * li}.bias"] = normal((n_out,), 0.1)
    hs = [AIM_SIZE, *HEAD_HIDDEN, 1]
    for li in range(3):
        std = scale * math.sqrt(2.0 / (hs[li] + hs[li + 1])) * 1.5
        sd[f"outputs.energy_mlp.mlp.{2 * li}.weight"] = normal((hs[li + 1], hs[li]), std)
        sd[f"outputs.energy_mlp.mlp.{2 * li}.bias"] = normal((hs[li + 1],), 0.1)
    sae = np.zeros((NUM_EMBED, 1), np.float64)
    for z in spec.implemented_species:
        sae[z, 0] = -13.6 * z * (1.0 + 0.05 * rng.standard_normal())
    sd["outputs.atomic_shift.shifts.weight"] = torch.from_numpy(sae)
    sd["outputs.srcoulomb.rc"] = torch.tensor(spec.coulomb_sr_rc, dtype=torch.float)
    return sd
