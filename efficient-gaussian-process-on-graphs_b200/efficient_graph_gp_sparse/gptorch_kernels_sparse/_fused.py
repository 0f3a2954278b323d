"""Shared by the two kernels: one fused Phi-blocks object for a list of step operators."""

import torch

from grf_b200.engine import PhiBlocks, phi_blocks_from_torch_csr


def fused_blocks(step_matrices) -> PhiBlocks:
    """The Phi blocks of ``step_matrices`` (a list of SparseLinearOperator / sparse CSR tensors,
    one per walk length).  GraphPreprocessor attaches them; otherwise they are built once here."""
    blocks = getattr(step_matrices, "phi_blocks", None)
    first = step_matrices[0]
    dev = getattr(first, "sparse_csr_tensor", first).device
    if blocks is not None and blocks.device == dev:
        return blocks
    tensors = [getattr(m, "sparse_csr_tensor", m) for m in step_matrices]
    blocks = phi_blocks_from_torch_csr(tensors)
    try:
        step_matrices.phi_blocks = blocks
    except AttributeError:
        pass
    return blocks
