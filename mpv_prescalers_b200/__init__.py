"""B200-native implementation of the mpv-prescalers per-pixel upscaling hot path."""
from .hookfile import HookError, HookFile, eval_rpn, find_hook  # noqa: F401

__all__ = ["HookError", "HookFile", "eval_rpn", "find_hook", "prescale", "resample"]


def __getattr__(name):  # lazy: importing the package must not require torch / the CUDA library
    if name in ("prescale", "resample", "plan"):
        from . import api

        return getattr(api, name)
    raise AttributeError(name)
