"""B200-native AIMNet2 inference engine behind the AIMNet2Calculator API."""
from .model_spec import ModelSpec, random_state_dict  # noqa: F401

__all__ = ["AIMNet2Calculator", "Engine", "ModelSpec", "random_state_dict"]


def __getattr__(name):
    # compute-path objects are imported lazily so that `import aimnetcentral_b200` works where only the host-side
    # utilities are needed (the CUDA library is loaded — or the import fails loudly — on first use)
    if name == "AIMNet2Calculator":
        from .calculator import AIMNet2Calculator

        return AIMNet2Calculator
    if name == "Engine":
        from .engine import Engine

        return Engine
    raise AttributeError(name)
