"""Elementwise guards of the reference's src/guard.py:7-14, kept for API parity (`from src.guard import guard_exp,
guard_sqrt`).  Inside the fused kernels the same clamps are applied in registers -- the [-75, 75] exponent clamp in the
mean-shift weight (csrc/meanshift.cu: ms_kernel_weight) and the 1e-5 floor under the square root of the residuals
(csrc/fit.cu: guard_sqrtf) -- so these tensor versions only serve callers that use them directly, on whatever device the
argument lives."""


def guard_exp(x, max_value=75, min_value=-75):
    """exp of x limited to [min_value, max_value] first (overflow / underflow guard)."""
    return x.clamp(min_value, max_value).exp()


def guard_sqrt(x, minimum=1e-5):
    """sqrt of x with a floor at `minimum` (keeps the gradient finite at zero in the reference's training code)."""
    return x.clamp_min(minimum).sqrt()
