"""B200-native AIMNet2 inference engine behind the AIMNet2Calculator API."""
from .model_spec import ModelSpec, random_state_dict  # noqa: F401

__all__ = ["AIMNet2Calculator", "AIMNet2ASE", "AIMNet2TorchSim", "AIMNet2Pysis", "Engine", "ModelSpec", "random_state_dict"]


def __getattr__(name):
    # compute-path objects are imported lazily so that `import aimnetcentral_b200` works where only the host-side
    # utilities are needed (the CUDA library is loaded — or the import fails loudly — on first use)
    if name == "AIMNet2Calculator":
        from .calculator import AIMNet2Calculator

        return AIMNet2Calculator
    if name == "AIMNet2ASE":
        from .aimnet2ase import AIMNet2ASE

        return AIMNet2ASE
    if name == "AIMNet2TorchSim":
        from .aimnet2torchsim import AIMNet2TorchSim

        return AIMNet2TorchSim
    if name == "AIMNet2Pysis":
        from .aimnet2pysis import AIMNet2Pysis

        return AIMNet2Pysis
    if name == "Engine":
        from .engine import Engine

        return Engine
    raise AttributeError(name)
