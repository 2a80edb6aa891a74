from .registry import Registry  # noqa: F401
from .root import BACKBONE_REGISTRY, DATASET_REGISTRY, HOOK_REGISTRY, MODULE_REGISTRY  # noqa: F401
