"""Model classes and kernel factory of the B200 path (mirror of the reference GP/ package for the hot path)."""
