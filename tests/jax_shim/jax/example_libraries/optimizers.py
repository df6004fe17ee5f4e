"""Imported by solver/optimizers.py at module load; the Adam loop of the reference uses optax, not this module."""
