"""Import stub for `warp-lang` (absent offline). TEST INFRASTRUCTURE ONLY.

The reference imports `warp` at module import (aimnet/kernels/__init__.py:24,
conv_sv_2d_sp_wp.py:29,75) but never executes a Warp kernel on CPU
(aimnet/modules/aev.py:163). This stub satisfies the import; nothing here
performs arithmetic.
"""


class _Cfg:
    version = "0.0-stub"
    quiet = True


config = _Cfg()


def init():
    return None


def get_cuda_device_count():
    return 0


def kernel(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]

    def deco(fn):
        return fn

    return deco


class _T:
    def __init__(self, *a, **k):
        pass

    def __class_getitem__(cls, item):
        return cls


class _Arr:
    def __call__(self, *a, **k):
        return _T

    def __getitem__(self, item):
        return _T


array = array1d = array2d = array3d = array4d = _Arr()
float32 = int32 = vec4f = vec3f = float64 = int64 = _T


def stream_from_torch(*a, **k):
    raise RuntimeError("warp stub: no CUDA")


def launch(*a, **k):
    raise RuntimeError("warp stub: kernels cannot run")


def from_torch(*a, **k):
    raise RuntimeError("warp stub: kernels cannot run")


def atomic_add(*a, **k):
    raise RuntimeError("warp stub")


def tid():
    raise RuntimeError("warp stub")


def dot(*a, **k):
    raise RuntimeError("warp stub")
