"""Inference-side hot path (peak finding, PAF grouping)."""
