"""Host-side mirror of the reference's `models/` module boundary (unified_arch / unified_llama / unified_qwen)."""
