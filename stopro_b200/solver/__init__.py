"""The optimiser driver that calls the path (optimize_by_adam with the reference signature)."""
