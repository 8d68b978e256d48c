"""Host-side mirror of the reference's `dataset/` front-end for the hot path, backed by GPU kernels."""
