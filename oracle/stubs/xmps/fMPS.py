"""Restated subset of xmps.fMPS (finite MPS, sites (d, Dl, Dr)).

``left_from_state``  : exact TT-SVD, left-canonical, final 1x1 S.V dropped.
``left_canonicalise``: right-to-left SVD sweep truncating each bond to
                       min(D, rank) ("truncate during right sweep"), then a
                       left-to-right sweep back.  RESTATED FROM MEMORY of the
                       published source; cannot be verified offline.
"""
import numpy as np

from .tensor import rank as _rank


class fMPS:
    def __init__(self, data=None, d=None, D=None):
        self.data = None
        if data is not None:
            self.data = list(data)
            self.L = len(self.data)
            self.d = self.data[0].shape[0] if d is None else d
            self.D = max(max(x.shape[1:]) for x in self.data)

    def __getitem__(self, k):
        return self.data[k]

    def __setitem__(self, k, v):
        self.data[k] = v

    def from_product_state(self, vectors):
        self.data = [np.expand_dims(np.expand_dims(np.asarray(v), -1), -1) for v in vectors]
        self.L = len(self.data)
        self.d = self.data[0].shape[0]
        self.D = 1
        return self

    def left_from_state(self, state):
        self.L = state.ndim
        self.d = state.shape[0]
        d, L = self.d, self.L
        psi = np.asarray(state).reshape(1, -1)
        data = []
        for n in range(L):
            b = psi.shape[0]
            q = psi.shape[1] // d
            M = psi.reshape(b, d, q).transpose(1, 0, 2).reshape(d * b, q)
            U, S, V = np.linalg.svd(M, full_matrices=False)
            r = max(1, _rank(S))
            data.append(U[:, :r].reshape(d, b, r))
            psi = S[:r, None] * V[:r]
        self.data = data
        self.D = max(max(x.shape[1:]) for x in data)
        return self

    def left_canonicalise(self, D=None, testing=False, sweep_back=True, minD=True):
        L, d = self.L, self.d
        if D is not None:
            self.D = min(D, self.D)
        for m in range(L - 1, 0, -1):
            _, Dl, Dr = self.data[m].shape
            M = self.data[m].transpose(1, 0, 2).reshape(Dl, d * Dr)
            U, S, V = np.linalg.svd(M, full_matrices=False)
            r = S.size
            if minD:
                r = max(1, _rank(S))
            if D is not None:
                r = min(r, D)
            self.data[m] = V[:r].reshape(r, d, Dr).transpose(1, 0, 2)
            self.data[m - 1] = np.einsum("dab,bc->dac", self.data[m - 1], U[:, :r] * S[None, :r])
        carry = np.ones((1, 1))
        for m in range(L):
            A = np.einsum("ab,dbc->dac", carry, self.data[m])
            _, Dl, Dr = A.shape
            U, S, V = np.linalg.svd(A.reshape(d * Dl, Dr), full_matrices=False)
            r = S.size
            if minD:
                r = max(1, _rank(S))
            self.data[m] = U[:, :r].reshape(d, Dl, r)
            carry = S[:r, None] * V[:r]
        return self
