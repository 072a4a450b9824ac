"""Baselines: the unmodified reference run offline through a shim, and a minimal NCCL translation.  Not product code."""
