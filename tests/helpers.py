"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# Tolerances (north_star): 1e-5 relative in fp32, 1e-2 in bf16.  "Relative" is
# max-abs(diff) / max-abs(reference) -- element-wise relative error is
# meaningless where a correlation of zero-mean features cancels to ~0
# ("Parity metric caveat").
TOL_F32 = 1e-5
TOL_BF16 = 1e-2


def rel_err(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    return float(np.abs(a - b).max() / denom) if a.size else 0.0


def load_cases(name):
    """tests/golden/<name>.npz -> {case: {field: array}}."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    cases = {}
    for k in z.files:
        case, field = k.split(".", 1)
        cases.setdefault(case, {})[field] = z[k]
    return cases
