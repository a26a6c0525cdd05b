from .tensor_core import Tensor, TensorNetwork, rand_uuid  # noqa: F401
