"""Small host helpers with the reference's names (xsdba.utils)."""
from __future__ import annotations

import numpy as np


def equally_spaced_nodes(n: int, eps: float | None = None) -> np.ndarray:
    """Nodes at the middle of ``n`` equal bins of [0, 1]; optional end points (utils.py:251-281)."""
    dq = 1 / n / 2
    q = np.linspace(dq, 1 - dq, n)
    if eps is None:
        return q
    return np.insert(np.append(q, 1 - eps), 0, eps)
