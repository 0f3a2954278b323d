"""grf_b200 -- host side of the B200-native GRF hot path.

Thin Python over the C ABI in ``include/grf_b200.h`` (hand-written sm_100a CUDA
kernels in ``csrc/``).  PyTorch is used for device memory, streams and
``torch.distributed`` only.  There is no CPU fallback: importing the engine
without the compiled library, or calling it without a CUDA device, raises.
"""

from . import _lib  # noqa: F401
from .engine import (  # noqa: F401
    DeviceGraph,
    PhiBlocks,
    StepMatrices,
    WalkConfig,
    build_phi_blocks,
    build_step_matrices,
    phi_blocks_from_scipy,
    run_walker,
)
