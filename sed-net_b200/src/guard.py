"""src/guard.py:7-14 (tiny elementwise guards used around the hot path; torch ops on the caller's device)."""
import torch


def guard_exp(x, max_value=75, min_value=-75):
    return torch.exp(torch.clamp(x, max=max_value, min=min_value))


def guard_sqrt(x, minimum=1e-5):
    return torch.sqrt(torch.clamp(x, min=minimum))
