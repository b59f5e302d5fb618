"""drba_b200 -- Blackwell (sm_100a) implementation of DRBA's per-triplet interpolation hot path."""
__version__ = "0.1.0"
