"""Synthetic MNIST / Fashion-MNIST shaped inputs.

The reference downloads MNIST through Keras (``/root/reference/tools.py:25-74``)
and scales pixels to ``k/255``.  There is no network and no dataset here, so the
tests and ``bench.py`` use images with the same shape and similar statistics:
28x28, ~80 % exact zeros, smooth strokes, values quantised to ``k/255`` so that
the singular spectra of the unfoldings decay the way real digits' do (white
noise would make every Jacobi sweep count unrealistic).

Also holds a random classifier generator with the bond profile the reference's
``compress_centre_one_site`` (``/root/reference/fMPO.py:338-453``) or
``compress_one_site`` (``fMPO.py:240-336``) produces, for throughput runs where
only the shapes matter.
"""
from __future__ import annotations

import numpy as np

N_PIXELS = 28 * 28
N_LABEL_LEG = 16  # 2**(int(log2(10))+1), variational_mpo_classifiers.py:39-42


def synthetic_images(n: int, seed: int = 0, dtype=np.float64, chunk: int = 8192,
                     labels: np.ndarray | None = None) -> np.ndarray:
    """``(n, 784)`` array in [0, 1], values k/255, MNIST-like sparsity.

    With ``labels`` the strokes are drawn around per-class prototypes (fixed by
    the class id, jittered per image) so that a batch-added classifier can beat
    chance, as the reference's tests expect of real digits
    (testing/test_deterministic_mpo_classifier.py:299-321)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:28, 0:28].astype(np.float32)
    out = np.empty((n, N_PIXELS), dtype=dtype)
    proto = None
    if labels is not None:
        labels = np.asarray(labels).astype(np.int64)
        prng = np.random.default_rng(424242)
        # per class: 4 strokes x (cx, cy, angle, half-length, width), 3-4 active
        proto = np.stack([
            prng.uniform(8, 20, size=(16, 4)), prng.uniform(7, 21, size=(16, 4)),
            prng.uniform(0, np.pi, size=(16, 4)), prng.uniform(3, 8, size=(16, 4)),
            prng.uniform(0.9, 1.6, size=(16, 4))], axis=-1).astype(np.float32)
    for lo in range(0, n, chunk):
        m = min(chunk, n - lo)
        img = np.zeros((m, 28, 28), dtype=np.float32)
        n_strokes = rng.integers(2, 5, size=m) if proto is None else np.full(m, 4)
        for k in range(4):
            active = (n_strokes > k).astype(np.float32)[:, None, None]
            # a stroke = gaussian tube around a short segment
            cx = rng.uniform(8, 20, size=(m, 1, 1)).astype(np.float32)
            cy = rng.uniform(7, 21, size=(m, 1, 1)).astype(np.float32)
            ang = rng.uniform(0, np.pi, size=(m, 1, 1)).astype(np.float32)
            half = rng.uniform(2, 8, size=(m, 1, 1)).astype(np.float32)
            width = rng.uniform(0.8, 1.8, size=(m, 1, 1)).astype(np.float32)
            if proto is not None:
                pk = proto[labels[lo:lo + m] % 16, k]  # (m, 5)
                cx = pk[:, 0, None, None] + 0.12 * (cx - 14)
                cy = pk[:, 1, None, None] + 0.12 * (cy - 14)
                ang = pk[:, 2, None, None] + 0.08 * (ang - np.pi / 2)
                half = pk[:, 3, None, None] + 0.1 * (half - 5)
                width = pk[:, 4, None, None] + 0.1 * (width - 1.3)
            dx, dy = np.cos(ang), np.sin(ang)
            rx, ry = xx[None] - cx, yy[None] - cy
            t = np.clip(rx * dx + ry * dy, -half, half)
            d2 = (rx - t * dx) ** 2 + (ry - t * dy) ** 2
            img += active * np.exp(-d2 / (2 * width ** 2))
        img = np.clip(img, 0.0, 1.0)
        img[img < 0.1] = 0.0
        img = np.round(img * 255.0) / 255.0
        out[lo:lo + m] = img.reshape(m, N_PIXELS)
    return out


def synthetic_labels(n: int, seed: int = 0, n_classes: int = 10) -> np.ndarray:
    rng = np.random.default_rng(seed + 10_000)
    return rng.integers(0, n_classes, size=n).astype(np.uint8)


def adversarial_images(dtype=np.float64) -> np.ndarray:
    """Edge cases named in SURVEY.md §8(d): zero image, single pixel, constant,
    a duplicate pair, a rank-1 outer-product image."""
    imgs = np.zeros((6, N_PIXELS), dtype=dtype)
    imgs[1, 300] = 1.0
    imgs[2, :] = 0.5
    imgs[3] = synthetic_images(1, seed=123, dtype=dtype)[0]
    imgs[4] = imgs[3]
    r = np.linspace(0.1, 1.0, 28)
    imgs[5] = np.outer(r, r[::-1]).reshape(-1)
    return imgs


def natural_bonds(L: int, D: int | None) -> list[int]:
    """Bond profile ``b_n = min(2^n, 2^(L-n), D)`` for n = 0..L."""
    return [min(2 ** n, 2 ** (L - n), D if D is not None else 2 ** L) for n in range(L + 1)]


def classifier_bonds(L: int, D: int, centre: int) -> list[int]:
    """Bond profile of a one-hairy-site classifier MPO.

    centre-hairy (``compress_centre_one_site``): both sweeps truncate, so
    ``B_n = min(2^n, 2^(L-n) * 1, D)`` — the label leg sits on the centre site and
    does not enlarge the natural bound of any *bond* because both sweeps stop
    at it.  last-site-hairy (``compress_one_site``, left sweep only):
    ``B_n = min(2^n, D)`` for n < L, ``B_L = 1``.
    """
    if centre == L - 1:
        return [min(2 ** n, D) for n in range(L)] + [1]
    return natural_bonds(L, D)


def random_classifier(L: int = 10, D: int = 32, centre: int | None = None,
                      s: int = N_LABEL_LEG, seed: int = 0, dtype=np.float64) -> list[np.ndarray]:
    """Random Gaussian one-hairy-site MPO, sites ``(d=2, s_n, B_n, B_{n+1})``,
    normalised to unit Frobenius norm of the dense operator (approximately, by
    per-site scaling) — shape-only stand-in for a trained classifier."""
    rng = np.random.default_rng(seed + 77)
    if centre is None:
        centre = L // 2
    bonds = classifier_bonds(L, D, centre)
    sites = []
    for n in range(L):
        sn = s if n == centre else 1
        w = rng.standard_normal((2, sn, bonds[n], bonds[n + 1]))
        w /= np.sqrt(2 * bonds[n])  # keeps the environments O(1)
        sites.append(w.astype(dtype))
    return sites
