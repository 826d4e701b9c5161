"""lctgan: hand-written sm_100a CUDA kernels (liblctgan_sm100.so, C ABI in include/lctgan.h) and
their autograd bindings for the LCT-GAN adversarial training step.

    import sys; sys.path.insert(0, "<repo>/lct-gan_b200")
    from models.generator import LCTEnhancer, LCTGeneratorConfig      # same API as the reference
    from models.discriminators import MultiPeriodDiscriminator, MultiScaleDiscriminator
    from datasets.tf_features import TFFeatures, TFFeaturesConfig
    import losses

Importing this package loads the shared library and fails loudly if it has not been built.
"""
from . import _lib

_lib.lib()   # no library, no product: raise at import

from . import ops, functional  # noqa: E402

__all__ = ["ops", "functional"]
