"""Drop-in for ``utils_sparse/sparse_lo.py:4-25`` -- the matvec boundary.

``SparseLinearOperator(sparse_csr_tensor)`` keeps the reference's protocol
(``_matmul``, ``_size``, ``_transpose_nonbatch``) but the product runs in the
grf_b200 CUDA SpMM on {int32 col, float32 val} blocks instead of cuSPARSE on
int64 indices, and the transpose is built once and cached instead of being
re-sorted on every call (sparse_lo.py:25).
"""

import torch

from grf_b200.linop import LinearOperator
from grf_b200.engine import PhiBlocks, phi_blocks_from_torch_csr


class SparseLinearOperator(LinearOperator):
    """A LinearOperator that wraps a sparse CSR tensor (GPU SpMM via grf_b200)."""

    def __init__(self, sparse_csr_tensor, _blocks: PhiBlocks = None, _transposed: bool = False):
        if not sparse_csr_tensor.is_sparse_csr:
            raise ValueError("Input tensor must be a sparse CSR tensor")
        self.sparse_csr_tensor = sparse_csr_tensor
        self._blocks = _blocks
        self._transposed = _transposed
        super().__init__(sparse_csr_tensor)

    @property
    def blocks(self) -> PhiBlocks:
        if self._blocks is None:
            self._blocks = phi_blocks_from_torch_csr(self.sparse_csr_tensor)
        return self._blocks

    def to(self, device):
        moved = self.sparse_csr_tensor.to(device)
        keep = self._blocks if (self._blocks is not None and self._blocks.device == moved.device) else None
        return SparseLinearOperator(moved, keep, self._transposed)

    def _matmul(self, rhs):
        one = torch.ones(1, dtype=torch.float32, device=self.blocks.device)
        squeeze = rhs.dim() == 1
        rhs2 = rhs[:, None] if squeeze else rhs
        out = self.blocks.apply_t(one, rhs2) if self._transposed else self.blocks.apply(one, rhs2)
        return out[:, 0] if squeeze else out

    def _size(self):
        return self.sparse_csr_tensor.size()

    def _base_tensor(self):
        return self.sparse_csr_tensor

    def _transpose_nonbatch(self):
        # The reference re-sorts to CSR here on every call (sparse_lo.py:25); the
        # Phi^T blocks are built once and shared, so transposing is free.
        return _TransposedSparseLinearOperator(self)


class _TransposedSparseLinearOperator(SparseLinearOperator):
    def __init__(self, base: SparseLinearOperator):
        self._base = base
        self._blocks = base.blocks
        self._transposed = not base._transposed
        self._csr = None
        LinearOperator.__init__(self, base._base_tensor())

    @property
    def sparse_csr_tensor(self):
        if self._csr is None:
            self._csr = self._base.sparse_csr_tensor.t().to_sparse_csr()
        return self._csr

    def _size(self):
        r, c = self._base._size()
        return torch.Size((c, r))

    def _transpose_nonbatch(self):
        return self._base
