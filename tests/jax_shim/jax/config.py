"""`from jax.config import config` (test/test_1_sinusoidal_direct_main.py:9): the shim is always float64."""


class _Config:
    def __init__(self):
        self.values = {"jax_enable_x64": True}

    def update(self, name, value):
        self.values[name] = value


config = _Config()
