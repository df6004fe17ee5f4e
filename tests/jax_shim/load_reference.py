"""Import the reference (/root/reference) UNMODIFIED as package `stopro` on top of the torch-backed jax shim.

    from load_reference import load_reference, REFERENCE_ROOT
    stopro = load_reference()          # None when /root/reference does not exist (e.g. on the GPU box)
    from stopro.GP.gp_poiseuille_independent import GPPoiseuilleIndependent

Test infrastructure only (tests/jax_shim/README.md).
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("STOPRO_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "GP"))


def load_reference():
    if not reference_available():
        return None
    if "stopro" in sys.modules and getattr(sys.modules["stopro"], "_jax_shim", False):
        return sys.modules["stopro"]
    try:
        import jax  # noqa: F401
        real = not os.path.abspath(jax.__file__).startswith(HERE)
    except ImportError:
        real = False
    if real:
        raise RuntimeError("a real jax is installed: use it directly instead of the shim")
    for p in (os.path.join(HERE, "stubs"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    import jax  # noqa: F401,F811  (the shim)
    # package `stopro` whose __path__ is the reference tree: its files are executed in place, byte for byte
    pkg = types.ModuleType("stopro")
    pkg.__path__ = [REFERENCE_ROOT]
    pkg.__spec__ = importlib.machinery.ModuleSpec("stopro", None, is_package=True)
    pkg.__spec__.submodule_search_locations = [REFERENCE_ROOT]
    pkg._jax_shim = True
    sys.modules["stopro"] = pkg
    sys.dont_write_bytecode = True  # never write __pycache__ into the read-only reference tree
    return pkg
