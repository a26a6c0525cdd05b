"""Container mirror of /root/reference/fMPO.py:21-51 — the type the classifier
arrives in.  Only the layout contract is mirrored (list of (d, s, i, j)
arrays); the compressors stay on the reference's host code (BASELINE.json)."""
from __future__ import annotations

import numpy as np


class fMPO:
    """finite MPO: list of site arrays (d, s, i, j)."""

    def __init__(self, data=None, d=None, s=None, D=None):
        if data is not None:
            data = list(data)
            self.L = len(data)
            self.d = d if d is not None else data[-1].shape[0]
            self.s = s if s is not None else max(x.shape[1] for x in data)
            self.D = max(max(x.shape[2:]) for x in data)
            self.data = data

    def __call__(self, k):
        return self.data[k + 1]

    def __getitem__(self, k):
        return self.data[k]

    def __setitem__(self, key, value):
        self.data[key] = value

    def __str__(self):
        return "fMPO: L={}, d={}, s={}, D={}".format(self.L, self.d, self.s, self.D)
