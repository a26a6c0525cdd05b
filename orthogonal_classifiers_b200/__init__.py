"""B200-native encode + classify path of Lewis1304/Orthogonal_Classifiers.

Importing this package loads the hand-written sm_100a CUDA library through its
C ABI (include/oc_b200.h).  There is no CPU fallback: without the built
library the import fails."""
from . import synthetic  # noqa: F401
from ._lib import OcError  # noqa: F401
from .batched import (  # noqa: F401
    DeviceClassifier,
    HostPipeline,
    MPOBatch,
    MPSBatch,
    classify,
    classify_host,
    count_correct,
    encode_images,
    evaluate,
    label_overlaps,
    set_force_generic,
)
from .deterministic_mpo_classifier import (  # noqa: F401
    adding_centre_batches,
    mpo_encoding,
    prepare_centred_batched_classifier,
)
from .distributed import all_reduce_counts, shard_bounds  # noqa: F401
from .fMPO import fMPO  # noqa: F401
from .tools import (  # noqa: F401
    data_to_QTN,
    fMPS_to_QTN,
    image_to_mps,
    load_qtn_classifier,
    save_qtn_classifier,
)
from .variational_mpo_classifiers import (  # noqa: F401
    classifier_predictions,
    ensemble_predictions,
    evaluate_classifier_top_k_accuracy,
    evaluate_hard_ensemble_top_k_accuracy,
    evaluate_soft_ensemble_top_k_accuracy,
    mps_encoding,
    padded_classifier_predictions,
    project_classifier,
)
