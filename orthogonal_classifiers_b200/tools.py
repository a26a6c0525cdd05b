"""Mirror of the hot-path entry points of /root/reference/tools.py with the
same names, argument meaning and array layouts, backed by the CUDA library.

quimb / xmps are not dependencies here: the returned objects are light
stand-ins exposing what the reference's callers touch (``.data``, ``.L``,
``.d``, ``.D``, ``.tensors``, ``.num_tensors``, ``.H``, ``.squeeze()`` and the
``@`` contraction against a classifier)."""
from __future__ import annotations

import os

import numpy as np

from . import batched


class fMPS:
    """What ``image_to_mps`` returns (tools.py:217-221; xmps.fMPS there):
    ``.data`` = list of L site arrays (d, Dl, Dr)."""

    def __init__(self, data, D=None):
        self.data = list(data)
        self.L = len(self.data)
        self.d = self.data[0].shape[0]
        self.D = D if D is not None else max(max(x.shape[1:]) for x in self.data)

    def __getitem__(self, k):
        return self.data[k]


class SiteTensor:
    """Stand-in for qtn.Tensor: data + index names."""

    def __init__(self, data, inds, tags=None):
        self.data = data
        self.inds = tuple(inds)
        self.tags = tags

    @property
    def shape(self):
        return self.data.shape


class OverlapResult:
    """What ``mps.H.squeeze() @ classifier.squeeze()`` evaluates to: a tensor
    whose ``.data`` is the label vector.  The reference's is signed and every
    caller takes ``np.abs`` of it (experiments.py:332-337); the sign of an MPS
    is a gauge choice, so here ``.data`` is already the non-negative vector."""

    def __init__(self, data):
        self.data = data

    def __abs__(self):
        return np.abs(self.data)


class MPSNetwork:
    """Stand-in for the quimb TensorNetwork ``fMPS_to_QTN`` builds
    (tools.py:224-234): ``.tensors[j].data`` is site j, inds (k{j}, prev, next).
    It may be a lazy view of one image inside a device-resident batch."""

    def __init__(self, sites=None, batch=None, index=None):
        self._sites = sites
        self.batch = batch
        self.index = index

    def _host_sites(self):
        if self._sites is None:
            self._sites = self.batch.sites_numpy(self.index)
        return self._sites

    @property
    def tensors(self):
        sites = self._host_sites()
        return tuple(SiteTensor(s, (f"k{j}", f"b{j}", f"b{j + 1}"), [f"{j}"]) for j, s in enumerate(sites))

    @property
    def num_tensors(self):
        return len(self._host_sites()) if self._sites is not None else self.batch.L

    @property
    def H(self):  # real data: conjugation is the identity
        return self

    def squeeze(self):
        return self

    def _single_batch(self, dtype):
        if self.batch is not None and self.index is not None:
            b = self.batch
            return batched.MPSBatch(b.data[self.index:self.index + 1], b.bonds, b.offsets, b.L, b.D, b.dtype)
        import torch
        sites = self._host_sites()
        L = len(sites)
        bonds = [1] + [s.shape[2] for s in sites]
        offs = [0]
        for s in sites:
            offs.append(offs[-1] + s.size)
        flat = np.concatenate([np.ascontiguousarray(s).reshape(-1) for s in sites])
        dt = "f64" if flat.dtype == np.float64 else "f32"
        t = torch.from_numpy(flat.astype(np.float64 if dt == "f64" else np.float32))[None].to(batched._device())
        return batched.MPSBatch(t, bonds, offs, L, None, dt)

    def __matmul__(self, classifier):
        """``mps.H.squeeze() @ classifier.squeeze()`` — one image; the batched
        form is ``label_overlaps``.  Returns the non-negative label vector
        (the sign of an MPS is a gauge choice; the reference takes np.abs)."""
        b = self._single_batch(None)
        if isinstance(classifier, MPONetwork):  # reuse the packed device copy across calls
            classifier = classifier.device_classifier(b.dtype)
        ov, _ = batched.overlaps_from_mps(b, classifier, return_argmax=False)
        return OverlapResult(ov[0].double().cpu().numpy())


class MPONetwork:
    """Stand-in for the TensorNetwork ``data_to_QTN`` builds (tools.py:244-254):
    sites (d, s, i, j), inds (k{j}, s{j}, prev, next)."""

    def __init__(self, data):
        self.data = [np.asarray(w) for w in data]
        self._device = {}

    @property
    def tensors(self):
        return tuple(SiteTensor(w, (f"k{j}", f"s{j}", f"c{j}", f"c{j + 1}"), [f"{j}"])
                     for j, w in enumerate(self.data))

    @property
    def num_tensors(self):
        return len(self.data)

    @property
    def H(self):
        return MPONetwork([np.conj(w) for w in self.data])

    def squeeze(self):
        return self

    def device_classifier(self, dtype="f32"):
        if dtype not in self._device:
            self._device[dtype] = batched.DeviceClassifier(self.data, dtype=dtype)
        return self._device[dtype]


"""
MPS encoding tools
"""


def image_to_mps(f_image, D, dtype="f32", trim=False, trunc=None):
    """tools.py:217-221.  ``trim=True`` removes zero trailing bond columns the
    way xmps' minD does (variable bonds per image); default keeps the fixed
    natural profile, zero padded.  ``trunc``: which sweep truncates to D —
    "sweep" (default: ``left_from_state(..).left_canonicalise(D)`` as the
    reference calls it) or "inline" (batched.DEFAULT_TRUNC)."""
    f_image = np.asarray(f_image)
    if f_image.ndim != 1:
        raise ValueError("image_to_mps takes one flat image")
    batch = batched.encode_images(f_image[None], D, dtype=dtype, trunc=trunc)
    L = batch.L
    Dmax = max(batch.bonds) if D is None else min(D, 2 ** (L // 2))
    return fMPS(batch.sites_numpy(0, trim=trim), D=Dmax if D is None else D)


def fMPS_to_QTN(fmps):
    """tools.py:224-234."""
    return MPSNetwork(sites=list(fmps.data))


"""
MPO encoding tools
"""


def data_to_QTN(data):
    """tools.py:244-254."""
    return MPONetwork(data)


"""
Classifier tools
"""


def save_qtn_classifier(QTN, dir, root="Classifiers"):
    """tools.py:262-266 — one ``site_{i}.npy`` per site."""
    path = os.path.join(root, dir)
    os.makedirs(path, exist_ok=True)
    for i, site in enumerate(QTN.tensors):
        np.save(os.path.join(path, f"site_{i}"), site.data)


def load_qtn_classifier(dir, root="Classifiers"):
    """tools.py:269-291 — re-expands squeezed dimensions on load."""
    path = os.path.join(root, dir)
    files = [f for f in os.listdir(path) if f.startswith("site_") and f.endswith(".npy")]
    num_files = len(files)
    data = []
    for i in range(num_files):
        site = np.load(os.path.join(path, f"site_{i}.npy"), allow_pickle=False)  # plain ndarrays: no pickle needed
        if i == 0:
            if len(site.shape) == 2:
                site = np.expand_dims(np.expand_dims(site, 1), 1)
            if len(site.shape) == 3:
                site = np.expand_dims(site, 2)
        elif i == num_files - 1:
            if len(site.shape) == 2:  # (d, Bl): a squeezed non-hairy last site (centre-hairy
                site = np.expand_dims(site, 1)  # classifiers; the reference's loader stops at 3-d)
            if len(site.shape) == 3:
                site = np.expand_dims(site, -1)
        else:
            if len(site.shape) != 4:
                site = np.expand_dims(site, 1)
        data.append(site)
    return data_to_QTN(data)
