"""stopro_b200 -- B200-native (sm_100a) PIGP hot path behind stopro's GP call signatures.

Layout mirrors the reference for the path only: ``GP/`` (model classes and kernel factory),
``sub_modules/loss_modules.py`` (logposterior), ``solver/optimizers.py`` (the Adam driver that calls the path),
``csrc/`` (CUDA kernels + the C ABI of include/pigp.h, built into ``libpigp.so``).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
