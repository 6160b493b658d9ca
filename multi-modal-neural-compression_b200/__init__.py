"""multi-modal-neural-compression_b200 — B200-native rate path of the ScaleHyperprior multi-task codecs.

Python/PyTorch host code over hand-written sm_100a CUDA kernels behind a C ABI (include/mmnc_b200.h).
The directory name is not a Python identifier; import it through the `mmnc_b200` alias at the repository root
(`import mmnc_b200 as mm`).  All sub-modules are imported eagerly so attribute access works through the alias.
"""
from . import (_lib, ops, entropy_models, layers, models, metrics, container, compressors, parallel, synthetic,  # noqa: F401
               input_pipeline)
from .container import Container  # noqa: F401
from .input_pipeline import GpuBatchLoader  # noqa: F401
from ._lib import build, launch_count  # noqa: F401
from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional, LowerBound  # noqa: F401
from .layers import GDN, NonNegativeParametrizer, conv, deconv  # noqa: F401
from .models import ScaleHyperprior, get_scale_table  # noqa: F401
from .compressors import (  # noqa: F401
    MultiTaskCompressor, MultiTaskMixedLatentCompressor, MultiTaskDisjointLatentCompressor,
    MultiTaskSharedLatentCompressor, SingleTaskCompressor, UncertaintyWeightingStrategy, NoWeightingStrategy,
    DummyModule, build_compressor, task_parameters)
from .parallel import DataParallel, FlatGradBucket  # noqa: F401
from .synthetic import synthetic_batch  # noqa: F401

__version__ = "0.1.0"
