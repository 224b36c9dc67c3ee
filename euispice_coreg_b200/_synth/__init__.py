"""Seeded synthetic FITS generators for the BASELINE.json configurations (no network, no real data)."""
from .scene import make_pair, make_config1, PairSpec  # noqa: F401
