"""coma_b200 — B200-native (sm_100a) implementation of snuvclab/coma's data-parallel hot paths."""
__version__ = "0.1.0"
