"""Batched host API over the C ABI: whole data sets per call instead of the
reference's one-image-per-Python-iteration loops
(/root/reference/variational_mpo_classifiers.py:102-104, experiments.py:332-337).

torch is used only for device memory and streams; all arithmetic happens in the
CUDA library (``_lib``).  Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import OC_F32, OC_F64, OC_U8, TRUNC, check, i32_array, lib

# Which sweep applies the D_encode truncation (include/oc_b200.h).  "sweep" follows
# the reference's call order at tools.py:221 — left_from_state(...) exact, then
# left_canonicalise(D): right-to-left truncating sweep, sweep back — and is the default of every
# reference-named entry point; "inline" truncates inside the left-to-right chain (BASELINE.json's
# wording).  The two coincide whenever D truncates nothing (D >= 32 for 28 x 28 images).
DEFAULT_TRUNC = "sweep"


def _trunc(trunc) -> int:
    if trunc is None:
        trunc = DEFAULT_TRUNC
    try:
        return TRUNC[trunc]
    except KeyError:
        raise ValueError(f"trunc must be 'sweep' or 'inline', got {trunc!r}") from None

_TORCH_DT = {"f32": torch.float32, "f64": torch.float64}
_OC_DT = {"f32": OC_F32, "f64": OC_F64}
_NP_DT = {"f32": np.float32, "f64": np.float64}


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("orthogonal_classifiers_b200 needs a CUDA device; there is no CPU path")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _stream_ptr() -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def set_force_generic(flag: bool) -> None:
    """Route calls through the runtime-shaped CUDA kernels (test aid)."""
    check(lib.oc_set_option(b"force_generic", int(bool(flag))))


def set_option(key: str, value: int) -> None:
    """``oc_set_option`` (include/oc_b200.h): implementation switches used by the
    tests and the timing scripts — ``enc_gram`` (Gram front end of the encoder),
    ``enc_pack`` (packed step 4), ``ov_static`` (compile-time overlap shapes), ..."""
    check(lib.oc_set_option(key.encode(), int(value)))


def n_sites_for(n_pix: int) -> int:
    """tools.py:218 — L = ceil(log2(n_pix))."""
    return max(1, int(np.ceil(np.log2(n_pix))))


def mps_layout(L: int, D: int | None):
    bonds = (C.c_int32 * (L + 1))()
    offs = (C.c_int64 * (L + 1))()
    check(lib.oc_mps_layout(L, 0 if D is None else int(D), bonds, offs))
    return list(bonds), list(offs)


def _in_dtype(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return OC_F32
    if t.dtype == torch.float64:
        return OC_F64
    if t.dtype == torch.uint8:
        return OC_U8
    raise TypeError(f"unsupported image dtype {t.dtype}")


def to_device_images(images, device=None) -> torch.Tensor:
    """(N, n_pix) float32/float64/uint8, numpy or torch, -> device tensor with
    unit inner stride.  float64 host arrays are converted to float32 on the
    host only when the compute dtype is f32 (done by the caller)."""
    dev = _device(device)
    if isinstance(images, torch.Tensor):
        t = images
    else:
        arr = np.asarray(images)
        if arr.ndim == 1:
            arr = arr[None]
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.ndim != 2:
        raise ValueError("images must be (N, n_pix)")
    if t.device != dev:
        t = t.to(dev, non_blocking=True)
    if t.stride(1) != 1 or (t.shape[0] > 1 and t.stride(0) < t.shape[1]):
        t = t.contiguous()
    return t


def _ld(t: torch.Tensor) -> int:
    """Row stride in elements (a single row may report any stride)."""
    return t.stride(0) if t.shape[0] > 1 else t.shape[1]


@dataclass
class MPSBatch:
    """N image MPSs in the fixed-shape layout of include/oc_b200.h.

    ``data`` is a device tensor (N, elems_per_image); site n of image i is
    ``data[i, offsets[n]:offsets[n+1]].view(2, bonds[n], bonds[n+1])`` — the
    reference's site layout (d, Dl, Dr) (tools.py:224-234)."""
    data: torch.Tensor
    bonds: list
    offsets: list
    L: int
    D: int | None
    dtype: str
    svals: torch.Tensor | None = None   # (N, L, width)
    sweeps: torch.Tensor | None = None  # (N, L) int32

    def __len__(self) -> int:
        return self.data.shape[0]

    def site(self, i: int, n: int) -> torch.Tensor:
        return self.data[i, self.offsets[n]:self.offsets[n + 1]].view(2, self.bonds[n], self.bonds[n + 1])

    def sites_numpy(self, i: int, trim: bool = False) -> list:
        """Site arrays of image i on the host.  ``trim`` removes the zero
        trailing bond columns/rows (what xmps' minD does)."""
        row = self.data[i].cpu().numpy()
        out = [row[self.offsets[n]:self.offsets[n + 1]].reshape(2, self.bonds[n], self.bonds[n + 1])
               for n in range(self.L)]
        if trim:
            keep = [1]
            for n in range(self.L - 1):
                nz = np.nonzero(np.any(out[n] != 0, axis=(0, 1)))[0]
                keep.append(int(nz.max()) + 1 if nz.size else 1)
            keep.append(1)
            out = [a[:, :keep[n], :keep[n + 1]] for n, a in enumerate(out)]
        return out


def encode_images(images, D: int | None, dtype: str = "f32", return_svals: bool = False,
                  return_sweeps: bool = False, device=None, trunc: str | None = None) -> MPSBatch:
    """Batched ``mps_encoding`` (variational_mpo_classifiers.py:102-104):
    every image -> unit-norm left-canonical MPS truncated to bond D.
    ``trunc``: "sweep" (default, the reference's order) or "inline"."""
    t = to_device_images(images, device)
    N, n_pix = t.shape
    L = n_sites_for(n_pix)
    bonds, offs = mps_layout(L, D)
    dev = t.device
    out = torch.empty((N, offs[L]), dtype=_TORCH_DT[dtype], device=dev)
    svals = sweeps = None
    if return_svals:
        w = lib.oc_svals_width(L, 0 if D is None else int(D))
        svals = torch.zeros((N, L, w), dtype=_TORCH_DT[dtype], device=dev)
    if return_sweeps:
        sweeps = torch.zeros((N, L), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.oc_encode_batch(
            t.data_ptr(), _in_dtype(t), N, n_pix, _ld(t), L, 0 if D is None else int(D),
            _OC_DT[dtype], _trunc(trunc), out.data_ptr(),
            svals.data_ptr() if svals is not None else None,
            sweeps.data_ptr() if sweeps is not None else None, _stream_ptr()))
    return MPSBatch(out, bonds, offs, L, D, dtype, svals, sweeps)


class DeviceClassifier:
    """A one-hairy-site classifier MPO packed on the device.

    ``sites``: the reference's container — ``fMPO(...).data`` /
    ``[t.data for t in qtn.tensors]``: L arrays (2, s_n, B_n, B_{n+1})
    (fMPO.py:21-51, tools.py:244-254), float64 or complex128 with zero
    imaginary part (fMPO.add, fMPO.py:668-679).  Squeezed arrays as
    ``load_qtn_classifier`` re-expands them (tools.py:269-291) are accepted
    in 4-d form only."""

    def __init__(self, sites, dtype: str = "f32", device=None):
        sites = [np.asarray(getattr(w, "data", w)) for w in sites]
        L = len(sites)
        for n, w in enumerate(sites):
            if w.ndim != 4:
                raise ValueError(f"classifier site {n} must be 4-d (d, s, Bl, Br), got {w.shape}")
        is_complex = any(np.iscomplexobj(w) for w in sites)
        host_dt = np.complex128 if is_complex else np.float64
        sites = [np.ascontiguousarray(w, dtype=host_dt) for w in sites]
        ptrs = (C.c_void_p * L)(*[w.ctypes.data for w in sites])
        shapes = i32_array([d for w in sites for d in w.shape])
        n_written = C.c_int64(0)
        bonds = (C.c_int32 * (L + 1))()
        s = (C.c_int32 * L)()
        check(lib.oc_pack_classifier(ptrs, shapes, L, int(is_complex), _OC_DT[dtype], None, 0,
                                     C.byref(n_written), bonds, s))
        packed = np.empty(n_written.value, dtype=_NP_DT[dtype])
        check(lib.oc_pack_classifier(ptrs, shapes, L, int(is_complex), _OC_DT[dtype],
                                     packed.ctypes.data, packed.size, C.byref(n_written), bonds, s))
        self.L = L
        self.dtype = dtype
        self.bonds = list(bonds)
        self.s = list(s)
        self.S = max(self.s)
        self.centre = int(np.argmax(self.s)) if self.S > 1 else L - 1
        self.packed_host = packed
        self._bonds_c = bonds
        self._s_c = s
        self.device = _device(device)
        self.packed = torch.from_numpy(packed).to(self.device)
        self._cached_on = set()

    def to(self, device) -> "DeviceClassifier":
        self._uncache()
        self.device = _device(device)
        self.packed = self.packed.to(self.device)
        return self

    def cache_on_stream(self, stream_ptr: int) -> None:
        """Tell the library the packed buffer is immutable, so that the overlap kernel's
        re-layout of it runs once per stream instead of once per call (oc_classifier_cache)."""
        if stream_ptr in self._cached_on:
            return
        with torch.cuda.device(self.device):
            check(lib.oc_classifier_cache(self.packed.data_ptr(), 1, stream_ptr))
        self._cached_on.add(stream_ptr)

    def _uncache(self) -> None:
        for sp in list(getattr(self, "_cached_on", ())):
            try:
                with torch.cuda.device(self.device):
                    lib.oc_classifier_cache(self.packed.data_ptr(), 0, sp)
            except Exception:
                pass
        self._cached_on = set()

    def __del__(self):
        self._uncache()


def _as_classifier(classifier, dtype: str, device=None) -> DeviceClassifier:
    if isinstance(classifier, DeviceClassifier):
        if classifier.dtype != dtype:
            raise ValueError(f"classifier packed as {classifier.dtype}, requested {dtype}")
        return classifier
    sites = getattr(classifier, "tensors", None)
    if sites is None:
        sites = getattr(classifier, "data", classifier)
    return DeviceClassifier(sites, dtype=dtype, device=device)


def overlaps_from_mps(mps: MPSBatch, classifier, return_argmax: bool = True, signed: bool = False, out=None):
    """Device tensors (N, S) |overlap| and (N,) int32 argmax.  ``signed``: the signed label
    vector (defined up to one global sign per image) instead of its absolute values.
    ``out``: optional (N, S) device tensor to write into."""
    clf = _as_classifier(classifier, mps.dtype, mps.data.device)
    if clf.L != mps.L:
        raise ValueError(f"classifier has {clf.L} sites, images have {mps.L}")
    N = len(mps)
    dev = mps.data.device
    ov = torch.empty((N, clf.S), dtype=_TORCH_DT[mps.dtype], device=dev) if out is None else out
    am = torch.empty((N,), dtype=torch.int32, device=dev) if return_argmax else None
    fn = lib.oc_overlap_batch_signed if signed else lib.oc_overlap_batch
    with torch.cuda.device(dev):
        clf.cache_on_stream(_stream_ptr())
        check(fn(
            mps.data.data_ptr(), i32_array(mps.bonds), clf.packed.data_ptr(), clf._bonds_c, clf._s_c,
            mps.L, N, _OC_DT[mps.dtype], ov.data_ptr(), am.data_ptr() if am is not None else None,
            _stream_ptr()))
    return ov, am


def label_mix(signed_ov: torch.Tensor, V, n_out: int, n_sum: int = 1) -> torch.Tensor:
    """``out[i, l] = sum_p | signed_ov[i] . V[:, p * n_out + l] |`` on the device
    (``oc_label_mix``): the label legs contracted with n_out * n_sum product bitstrings
    at once.  V: (S, n_out * n_sum) host array."""
    N, S = signed_ov.shape
    dt = "f64" if signed_ov.dtype == torch.float64 else "f32"
    Vd = torch.as_tensor(np.ascontiguousarray(V, dtype=_NP_DT[dt])).to(signed_ov.device)
    if Vd.shape != (S, n_out * n_sum):
        raise ValueError(f"V must be ({S}, {n_out * n_sum}), got {tuple(Vd.shape)}")
    out = torch.empty((N, n_out), dtype=signed_ov.dtype, device=signed_ov.device)
    with torch.cuda.device(signed_ov.device):
        check(lib.oc_label_mix(signed_ov.data_ptr(), N, S, Vd.data_ptr(), n_out, n_sum, _OC_DT[dt],
                               out.data_ptr(), _stream_ptr()))
    return out


def ensemble_vote(pred: torch.Tensor, n_labels: int, want_norm: bool = True):
    """``oc_ensemble_vote`` on a (K, N, S) device tensor of member |overlaps|: returns
    (normalised (K, N, n_labels) or None, soft sum (N, n_labels), soft argmax (N,), hard
    argmax (N,)) as device tensors."""
    K, N, S = pred.shape
    dt = "f64" if pred.dtype == torch.float64 else "f32"
    dev = pred.device
    norm = torch.empty((K, N, n_labels), dtype=pred.dtype, device=dev) if want_norm else None
    soft = torch.empty((N, n_labels), dtype=pred.dtype, device=dev)
    sa = torch.empty((N,), dtype=torch.int32, device=dev)
    ha = torch.empty((N,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.oc_ensemble_vote(pred.data_ptr(), K, N, S, n_labels, _OC_DT[dt],
                                   norm.data_ptr() if norm is not None else None, soft.data_ptr(),
                                   sa.data_ptr(), ha.data_ptr(), _stream_ptr()))
    return norm, soft, sa, ha


def classify_device(images_dev: torch.Tensor, clf: DeviceClassifier, D: int | None,
                    workspace: torch.Tensor | None = None, out=None, trunc: str | None = None):
    """encode + classify with device-resident inputs; returns device tensors
    (overlaps (N, S), argmax (N,)).  ``workspace``: optional (N, elems) MPS
    scratch to avoid the library's internal allocation."""
    N, n_pix = images_dev.shape
    L = n_sites_for(n_pix)
    dev = images_dev.device
    if out is None:
        ov = torch.empty((N, clf.S), dtype=_TORCH_DT[clf.dtype], device=dev)
        am = torch.empty((N,), dtype=torch.int32, device=dev)
    else:
        ov, am = out
    with torch.cuda.device(dev):
        clf.cache_on_stream(_stream_ptr())
        check(lib.oc_encode_classify(
            images_dev.data_ptr(), _in_dtype(images_dev), N, n_pix, _ld(images_dev), L,
            0 if D is None else int(D), _OC_DT[clf.dtype], _trunc(trunc),
            clf.packed.data_ptr(), clf._bonds_c, clf._s_c,
            workspace.data_ptr() if workspace is not None else None,
            ov.data_ptr(), am.data_ptr(), _stream_ptr()))
    return ov, am


class HostPipeline:
    """End-to-end encode + classify for HOST-resident images: the batch is cut
    into chunks; the pinned-host -> device copy of chunk k+1 runs on a copy
    stream while the kernels of chunk k run on the compute stream, and the
    results return to pinned host buffers asynchronously.  Device staging is
    allocated once and reused across calls.  (The same schedule inside the
    library, behind the C ABI: ``classify_host`` / ``oc_classify_host``.)"""

    def __init__(self, clf: DeviceClassifier, n_pix: int, D: int | None, chunk: int = 17_500,
                 in_dtype=torch.float32, trunc: str | None = None):
        self.clf, self.D, self.chunk, self.n_pix, self.trunc = clf, D, int(chunk), n_pix, trunc
        dev = clf.device
        self.dev = dev
        L = n_sites_for(n_pix)
        _, offs = mps_layout(L, D)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.stage = [torch.empty((self.chunk, n_pix), dtype=in_dtype, device=dev) for _ in range(2)]
        self.ws = torch.empty((self.chunk, offs[-1]), dtype=_TORCH_DT[clf.dtype], device=dev)
        self.ov = [torch.empty((self.chunk, clf.S), dtype=_TORCH_DT[clf.dtype], device=dev) for _ in range(2)]
        self.am = [torch.empty((self.chunk,), dtype=torch.int32, device=dev) for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)]
        self.freed = [torch.cuda.Event() for _ in range(2)]

    def run(self, images_host: torch.Tensor, ov_host: torch.Tensor, am_host: torch.Tensor) -> None:
        """images_host (N, n_pix) pinned; ov_host (N, S), am_host (N,) pinned.
        Returns after the results are on the host."""
        N = images_host.shape[0]
        comp = torch.cuda.current_stream(self.dev)
        for k, lo in enumerate(range(0, N, self.chunk)):
            hi = min(N, lo + self.chunk)
            b = k & 1
            with torch.cuda.stream(self.copy_stream):
                if k >= 2:
                    self.copy_stream.wait_event(self.freed[b])  # kernels of chunk k-2 done with this buffer
                self.stage[b][: hi - lo].copy_(images_host[lo:hi], non_blocking=True)
                self.copied[b].record(self.copy_stream)
            comp.wait_event(self.copied[b])
            classify_device(self.stage[b][: hi - lo], self.clf, self.D, workspace=self.ws,
                            out=(self.ov[b][: hi - lo], self.am[b][: hi - lo]), trunc=self.trunc)
            self.freed[b].record(comp)
            ov_host[lo:hi].copy_(self.ov[b][: hi - lo], non_blocking=True)
            am_host[lo:hi].copy_(self.am[b][: hi - lo], non_blocking=True)
        comp.synchronize()


def _host_array(x):
    """numpy view of a host tensor / array with unit inner stride (no copy when possible)."""
    if isinstance(x, torch.Tensor):
        if x.device.type != "cpu":
            raise ValueError("classify_host takes HOST images")
        x = x.numpy()
    x = np.asarray(x)
    if x.ndim == 1:
        x = x[None]
    if x.ndim != 2:
        raise ValueError("images must be (N, n_pix)")
    if x.dtype not in (np.float32, np.float64, np.uint8):
        x = x.astype(np.float64)
    if x.strides[1] != x.itemsize or (x.shape[0] > 1 and x.strides[0] % x.itemsize):
        x = np.ascontiguousarray(x)
    return x


_NP_IN = {np.dtype(np.float32): OC_F32, np.dtype(np.float64): OC_F64, np.dtype(np.uint8): OC_U8}


def classify_host(images_host, classifier, D_encode: int | None = 32, dtype: str = "f32", out=None,
                  trunc: str | None = None, device=None):
    """HOST buffers in, HOST buffers out through ONE C-ABI call (``oc_classify_host``):
    the library pipelines host->device copies, kernels and the result copies itself.

    images_host: (N, n_pix) numpy array or CPU tensor, float32 / float64 / uint8 (MNIST's
    own type; /255 happens in the encoder's load).  Pinned memory overlaps the copies.
    ``out``: optional (overlaps (N, S), argmax (N,) int32) host arrays/tensors to fill.
    Returns (overlaps, argmax) as numpy arrays of the compute dtype / int32."""
    dev = _device(device)
    clf = _as_classifier(classifier, dtype, dev)
    x = _host_array(images_host)
    N, n_pix = x.shape
    L = n_sites_for(n_pix)
    if out is None:
        ov = np.empty((N, clf.S), dtype=_NP_DT[dtype])
        am = np.empty((N,), dtype=np.int32)
    else:
        ov, am = (o.numpy() if isinstance(o, torch.Tensor) else o for o in out)
        if ov.shape != (N, clf.S) or ov.dtype != _NP_DT[dtype] or not ov.flags.c_contiguous:
            raise ValueError("out[0] must be a C-contiguous (N, S) array of the compute dtype")
        if am.shape != (N,) or am.dtype != np.int32 or not am.flags.c_contiguous:
            raise ValueError("out[1] must be a C-contiguous (N,) int32 array")
    ld = x.strides[0] // x.itemsize if N > 1 else n_pix
    with torch.cuda.device(dev):
        check(lib.oc_classify_host(
            x.ctypes.data, _NP_IN[x.dtype], N, n_pix, ld, L, 0 if D_encode is None else int(D_encode),
            _OC_DT[dtype], _trunc(trunc), clf.packed_host.ctypes.data, clf.packed_host.size,
            clf._bonds_c, clf._s_c, ov.ctypes.data, am.ctypes.data, _stream_ptr()))
    return ov, am


def label_overlaps(images_or_mps, classifier, D_encode: int | None = 32, dtype: str = "f32",
                   device=None, trunc: str | None = None) -> np.ndarray:
    """The reference's inline prediction expression, batched:
    ``[np.abs((mps.H.squeeze() @ classifier.squeeze()).data) for mps in ...]``
    (experiments.py:332-337) -> (N, 16) float64 host array."""
    if isinstance(images_or_mps, MPSBatch):
        ov, _ = overlaps_from_mps(images_or_mps, classifier, return_argmax=False)
    elif hasattr(images_or_mps, "batch") and isinstance(images_or_mps.batch, MPSBatch):
        ov, _ = overlaps_from_mps(images_or_mps.batch, classifier, return_argmax=False)
    else:
        t = to_device_images(images_or_mps, device)
        clf = _as_classifier(classifier, dtype, t.device)
        ov, _ = classify_device(t, clf, D_encode, trunc=trunc)
    return ov.double().cpu().numpy()


def classify(images, classifier, D_encode: int | None = 32, dtype: str = "f32", device=None,
             trunc: str | None = None):
    """Host in, host out: (overlaps (N, S) float64, argmax (N,) int32)."""
    t = to_device_images(images, device)
    clf = _as_classifier(classifier, dtype, t.device)
    ov, am = classify_device(t, clf, D_encode, trunc=trunc)
    return ov.double().cpu().numpy(), am.cpu().numpy()


def count_correct(argmax_dev: torch.Tensor, labels_dev: torch.Tensor) -> torch.Tensor:
    """Device int64[2] = (correct, total); integer part of
    evaluate_classifier_top_k_accuracy (k = 1)."""
    if labels_dev.dtype != torch.uint8:
        labels_dev = labels_dev.to(torch.uint8)
    out = torch.zeros(2, dtype=torch.int64, device=argmax_dev.device)
    with torch.cuda.device(argmax_dev.device):
        check(lib.oc_count_correct(argmax_dev.data_ptr(), labels_dev.data_ptr(), argmax_dev.numel(),
                                   out.data_ptr(), _stream_ptr()))
    return out


def evaluate(images, labels, classifier, D_encode: int | None = 32, dtype: str = "f32",
             device=None, all_reduce: bool = True, trunc: str | None = None) -> float:
    """Top-1 accuracy of this rank's shard, summed over ranks with one
    all-reduce of two int64 when torch.distributed is initialised."""
    t = to_device_images(images, device)
    clf = _as_classifier(classifier, dtype, t.device)
    _, am = classify_device(t, clf, D_encode, trunc=trunc)
    lab = torch.as_tensor(np.asarray(labels), dtype=torch.uint8).to(t.device)
    counts = count_correct(am, lab)
    if all_reduce:
        from .distributed import all_reduce_counts
        all_reduce_counts(counts)
    c = counts.cpu()
    return float(c[0].item()) / max(1, int(c[1].item()))


# --------------------------------------------------------------------------
# classifier construction ("batch adding") on the device, fp64
# --------------------------------------------------------------------------

class MPOBatch:
    """n one-hairy-site MPOs of one fixed zero-padded shape on the device (fp64): what the
    reference holds as a list of ``fMPO`` (fMPO.py:21-51).  ``batch[i]`` is an ``fMPO`` whose
    ``.data`` are host arrays (d, s, i, j)."""

    def __init__(self, data: torch.Tensor, bonds, L: int, c: int, S: int):
        self.data, self.bonds, self.L, self.c, self.S = data, [int(b) for b in bonds], int(L), int(c), int(S)
        self.s = [S if n == c else 1 for n in range(L)]
        offs = [0]
        for n in range(L):
            offs.append(offs[-1] + 2 * self.s[n] * self.bonds[n] * self.bonds[n + 1])
        self.offsets = offs
        if data.shape[1] != offs[-1]:
            raise ValueError(f"MPO block of {data.shape[1]} elements, bonds say {offs[-1]}")

    def __len__(self) -> int:
        return self.data.shape[0]

    def sites_numpy(self, i: int) -> list:
        row = self.data[i].cpu().numpy()
        return [row[self.offsets[n]:self.offsets[n + 1]].reshape(2, self.s[n], self.bonds[n], self.bonds[n + 1]).copy()
                for n in range(self.L)]

    def __getitem__(self, i):
        from .fMPO import fMPO
        if isinstance(i, slice):
            return MPOBatch(self.data[i], self.bonds, self.L, self.c, self.S)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return fMPO(self.sites_numpy(i))

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def class_encode(mps: MPSBatch, labels, centre: int | None = None, S: int = 16) -> MPOBatch:
    """deterministic_mpo_classifier.py:21-42 ``mpo_encoding`` for one-hot label bitstrings:
    image MPS x label -> one-hairy-site MPO (``oc_class_encode``)."""
    c = mps.L // 2 if centre is None else int(centre)
    N = len(mps)
    lab = torch.as_tensor(np.asarray(labels).astype(np.uint8)).to(mps.data.device)
    if lab.numel() != N:
        raise ValueError("one label per image")
    if N and int(lab.max()) >= S:
        raise ValueError("label does not fit the label leg")
    csz = 2 * mps.bonds[c] * mps.bonds[c + 1]
    out = torch.empty((N, mps.offsets[-1] + csz * (S - 1)), dtype=torch.float64, device=mps.data.device)
    with torch.cuda.device(mps.data.device):
        check(lib.oc_class_encode(mps.data.data_ptr(), _OC_DT[mps.dtype], N, mps.L, i32_array(mps.bonds),
                                  lab.data_ptr(), c, S, out.data_ptr(), _stream_ptr()))
    return MPOBatch(out, mps.bonds, mps.L, c, S)


def add_compress_centre(mpos: MPOBatch, D: int | None, batch_num: int, orthogonalise: bool = False) -> MPOBatch:
    """``adding_centre_batches`` (experiments.py:497-523): every ``batch_num`` consecutive MPOs are
    summed (fMPO.add) and compressed around the centre site to bond D
    (fMPO.compress_centre_one_site, polar factor when ``orthogonalise``) by ONE kernel launch
    (``oc_add_compress_centre``), one CTA per group."""
    n = len(mpos)
    bonds_out = (C.c_int32 * (mpos.L + 1))()
    Dv = 0 if D is None else int(D)
    check(lib.oc_add_compress_centre(None, n, mpos.L, mpos.c, mpos.S, i32_array(mpos.bonds), int(batch_num), Dv,
                                     int(bool(orthogonalise)), None, bonds_out, None))
    bo = list(bonds_out)
    s = mpos.s
    elems = sum(2 * s[k] * bo[k] * bo[k + 1] for k in range(mpos.L))
    n_groups = (n + batch_num - 1) // batch_num
    out = torch.empty((n_groups, elems), dtype=torch.float64, device=mpos.data.device)
    data = mpos.data if mpos.data.is_contiguous() else mpos.data.contiguous()
    with torch.cuda.device(mpos.data.device):
        check(lib.oc_add_compress_centre(data.data_ptr(), n, mpos.L, mpos.c, mpos.S, i32_array(mpos.bonds),
                                         int(batch_num), Dv, int(bool(orthogonalise)), out.data_ptr(), bonds_out,
                                         _stream_ptr()))
    return MPOBatch(out, bo, mpos.L, mpos.c, mpos.S)
