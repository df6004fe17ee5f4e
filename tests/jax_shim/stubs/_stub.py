"""A module whose every attribute is a callable / attribute-bearing dummy: stands in for plotting and I/O packages the
reference imports at module load but never calls on the numerical path (tests/jax_shim/README.md)."""


class Dummy:
    def __init__(self, name="dummy"):
        self._name = name

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return Dummy(f"{self._name}.{item}")

    def __call__(self, *a, **k):
        return Dummy(f"{self._name}()")

    def __iter__(self):
        return iter(())

    def __getitem__(self, item):
        return Dummy(f"{self._name}[]")

    def __repr__(self):
        return f"<stub {self._name}>"


def install(module_globals, name):
    def __getattr__(item):
        if item.startswith("__"):
            raise AttributeError(item)
        return Dummy(f"{name}.{item}")

    module_globals["__getattr__"] = __getattr__
