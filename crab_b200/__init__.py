"""crab_b200 — B200-native (sm_100a) kernels + host mirror for Crab's AV-prompt -> prefill -> decode hot path."""
__version__ = "0.1.0"
