"""Named-index tensors with full contraction — the subset of quimb the
reference's encode/classify/batch-adding code uses."""
import itertools
import uuid

import numpy as np


def rand_uuid():
    return "_" + uuid.uuid4().hex[:12]


class Tensor:
    def __init__(self, data=1.0, inds=(), tags=None):
        self.data = np.asarray(data)
        self.inds = tuple(inds)
        self.tags = tags
        assert self.data.ndim == len(self.inds), (self.data.shape, self.inds)

    @property
    def shape(self):
        return self.data.shape

    @property
    def H(self):
        return Tensor(np.conj(self.data), self.inds, self.tags)

    def squeeze(self):
        keep = [i for i, d in enumerate(self.data.shape) if d != 1]
        return Tensor(self.data.reshape([self.data.shape[i] for i in keep]),
                      [self.inds[i] for i in keep], self.tags)

    def __truediv__(self, x):
        return Tensor(self.data / x, self.inds, self.tags)

    def __mul__(self, x):
        return Tensor(self.data * x, self.inds, self.tags)

    def __matmul__(self, other):
        return _contract(_combine(_as_list(self), _as_list(other)))

    def __abs__(self):
        return abs(self.data)


def _as_list(x):
    return list(x.tensors) if isinstance(x, TensorNetwork) else [x]


def _inner(tensors):
    cnt = {}
    for t in tensors:
        for i in t.inds:
            cnt[i] = cnt.get(i, 0) + 1
    return {i for i, c in cnt.items() if c >= 2}


def _combine(left, right):
    """quimb renames the inner indices of the network being added when they
    collide with existing ones, so ``tn.H @ tn`` is <tn|tn>."""
    used = {i for t in left for i in t.inds}
    clash = {i: rand_uuid() for i in _inner(right) if i in used}
    out = list(left)
    for t in right:
        out.append(Tensor(t.data, [clash.get(i, i) for i in t.inds], t.tags))
    return out


def _contract_pairwise(tensors):
    """Greedy pairwise contraction for networks with more than 52 distinct indices (einsum's
    alphabet): always contract the pair sharing an index whose result is smallest."""
    cnt = {}
    for t in tensors:
        for i in t.inds:
            cnt[i] = cnt.get(i, 0) + 1
    out_inds = [i for t in tensors for i in t.inds if cnt[i] == 1]
    ts = [(np.asarray(t.data), list(t.inds)) for t in tensors]
    while len(ts) > 1:
        best = None
        for a in range(len(ts)):
            for b in range(a + 1, len(ts)):
                shared = [i for i in ts[a][1] if i in ts[b][1]]
                if not shared:
                    continue
                size = 1
                for d, i in zip(ts[a][0].shape, ts[a][1]):
                    if i not in shared:
                        size *= d
                for d, i in zip(ts[b][0].shape, ts[b][1]):
                    if i not in shared:
                        size *= d
                if best is None or size < best[0]:
                    best = (size, a, b, shared)
        if best is None:  # disconnected pieces: outer product of the first two
            a, b, shared = 0, 1, []
        else:
            _, a, b, shared = best
        (da, ia), (db, ib) = ts[a], ts[b]
        data = np.tensordot(da, db, axes=([ia.index(i) for i in shared], [ib.index(i) for i in shared]))
        inds = [i for i in ia if i not in shared] + [i for i in ib if i not in shared]
        ts = [t for k, t in enumerate(ts) if k not in (a, b)] + [(data, inds)]
    data, inds = ts[0]
    if not out_inds:
        return data.item() if data.ndim == 0 else data
    return Tensor(np.transpose(data, [inds.index(i) for i in out_inds]), out_inds)


def _contract(tensors):
    if len({i for t in tensors for i in t.inds}) > 52:
        return _contract_pairwise(tensors)
    names = {}
    letters = itertools.chain("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ")
    for t in tensors:
        for i in t.inds:
            if i not in names:
                names[i] = next(letters)
    cnt = {}
    for t in tensors:
        for i in t.inds:
            cnt[i] = cnt.get(i, 0) + 1
    out_inds = [i for t in tensors for i in t.inds if cnt[i] == 1]
    expr = ",".join("".join(names[i] for i in t.inds) for t in tensors)
    expr += "->" + "".join(names[i] for i in out_inds)
    data = np.einsum(expr, *[t.data for t in tensors], optimize="greedy")
    if not out_inds:
        return data.item() if data.ndim == 0 else data
    return Tensor(data, out_inds)


class TensorNetwork:
    def __init__(self, tensors=()):
        ts = []
        for t in tensors:
            if isinstance(t, TensorNetwork):
                ts = _combine(ts, list(t.tensors))
            else:
                ts.append(t)
        self.tensors = tuple(ts)

    @property
    def num_tensors(self):
        return len(self.tensors)

    @property
    def H(self):
        return TensorNetwork([t.H for t in self.tensors])

    def squeeze(self):
        return TensorNetwork([t.squeeze() for t in self.tensors])

    def __or__(self, other):
        return TensorNetwork(_combine(list(self.tensors), _as_list(other)))

    __and__ = __or__

    def __matmul__(self, other):
        return _contract(_combine(list(self.tensors), _as_list(other)))

    def __rmatmul__(self, other):
        return _contract(_combine(_as_list(other), list(self.tensors)))

    def __xor__(self, what):
        return _contract(list(self.tensors))

    def __truediv__(self, x):
        ts = list(self.tensors)
        ts[0] = ts[0] / x
        return TensorNetwork(ts)

    def __iter__(self):
        return iter(self.tensors)
