"""rdg_b200 -- B200-native (sm_100a) RainDisaggGAN hot path.

Generator forward / critic forward / WGAN-GP training step as hand-written CUDA behind the
reference's own Python surface.  See DESIGN.md and include/rdg_b200.h.
"""
from . import weights
from ._lib import MODE_BF16, MODE_FP16, MODE_FP32, NonFiniteError, RdgError, load as load_library

__all__ = ["weights", "load_library", "NonFiniteError", "RdgError", "MODE_FP32", "MODE_BF16", "MODE_FP16",
           "Context", "Generator", "Critic"]


def __getattr__(name):
    # engine imports torch; keep `import rdg_b200` cheap for host-only users (weights, hdf5)
    if name in ("Context", "Generator", "Critic"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)
