"""CPU oracle for the encode + classify path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  Nothing under
``orthogonal_classifiers_b200/`` imports it; the product path has no CPU
fallback and fails loudly when the CUDA library is missing.

Parity status:

* ``D_encode >= 32`` — pinned by mathematics: the encoding is exact, so
  ``overlaps == |x_hat^T W|`` with ``W`` the dense contraction of the classifier
  (``dense_classifier``), independent of SVD gauge and of how xmps orders its
  sweeps.
* classifier construction (``fMPO.add`` / ``compress_centre_one_site`` /
  ``compress_one_site``), label bitstrings, class encoding, top-k accuracy —
  pinned against the reference's OWN code, executed here from
  ``/root/reference`` with stand-ins for the un-vendored third-party modules
  (``oracle/stubs``); fixtures under ``tests/golden/`` made by
  ``tests/golden/generate_golden.py``.
* ``D_encode < 32`` — PARITY UNPINNED: the truncation order lives in the
  un-vendored ``xmps.fMPS.left_canonicalise`` (requirements.txt:69,
  ``xmps @ 0fe1c7ec``), no reference test holds numeric values, and the golden
  prediction arrays are listed in ``.MISSING_LARGE_BLOBS``.  Two variants are
  restated (``variant="inline"`` — BASELINE.json's wording — and
  ``variant="sweep"`` — exact chain, right-to-left truncating sweep, sweep back).
"""
from .mps_oracle import (  # noqa: F401
    pad_image,
    encode_image,
    mps_to_state,
    bond_singular_values,
    sweep_singular_values,
    overlaps_chain,
    dense_classifier,
    overlaps_dense,
    top_k_accuracy,
    classify_reference_loop,
)
from .mpo_oracle import (  # noqa: F401
    hairy_bitstrings,
    class_encode,
    mpo_add,
    compress_centre_one_site,
    compress_one_site,
    adding_centre_batches,
    adding_batches_one_site,
    prepare_centred_sum_states,
    build_centred_classifier,
    mpo_norm,
    real_sites,
)
