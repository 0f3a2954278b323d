"""Fused lazy operators for the GRF kernel: Phi = sum_l f_l M_l and K = Phi[x1] Phi[x2]^T.

These replace what the reference assembles out of upstream pieces at every
forward (sparse_grf_kernel.py:24-62): L ``ConstantMulLinearOperator``s over
``SparseLinearOperator``s inside a ``SumLinearOperator``, wrapped in
``InterpolatedLinearOperator``s for the row selection and a
``MatmulLinearOperator`` for K.  Here one operator owns the Phi blocks and one
CUDA call does a whole half of the matvec; the modulator stays a live autograd
leaf: its gradient is the per-length reduction ``grf_phi_fgrad``.
"""

from __future__ import annotations

from typing import Optional

import torch

from .engine import PhiBlocks
from .linop import LinearOperator


def _reduce_partial(u: torch.Tensor, group) -> torch.Tensor:
    """Sum of the per-GPU partials Phi_g^T V_g (the one collective of the path)."""
    if group is None:
        return u
    import torch.distributed as dist

    u = u.contiguous()
    dist.all_reduce(u, group=None if group is True else group)
    return u


class PhiApply(torch.autograd.Function):
    """y = Phi[rows] x   (transposed=False)   or   y = Phi[rows]^T x   (transposed=True)."""

    @staticmethod
    def forward(ctx, f, x, blocks: PhiBlocks, rows, transposed: bool, group):
        ctx.blocks, ctx.rows, ctx.transposed, ctx.group = blocks, rows, transposed, group
        fd = f.detach()
        y = blocks.apply_t(fd, x, rows) if transposed else blocks.apply(fd, x, rows)
        if transposed:
            y = _reduce_partial(y, group)
        ctx.save_for_backward(fd, x.detach())
        return y

    @staticmethod
    def backward(ctx, g):
        f, x = ctx.saved_tensors
        blocks, rows = ctx.blocks, ctx.rows
        g = g.contiguous()
        grad_f = grad_x = None
        if ctx.transposed:
            # y = Phi^T x:  dL/dx = Phi g ; dL/df_l = sum_k <x[k], M_l[rows[k]] g>
            if ctx.needs_input_grad[1]:
                grad_x = blocks.apply(f, g, rows)
            if ctx.needs_input_grad[0]:
                grad_f = blocks.fgrad_half(rows, x, g)
        else:
            # y = Phi x:  dL/dx = Phi^T g (a partial sum when sharded) ; dL/df_l = sum_k <g[k], M_l[rows[k]] x>
            if ctx.needs_input_grad[1]:
                grad_x = _reduce_partial(blocks.apply_t(f, g, rows), ctx.group)
            if ctx.needs_input_grad[0]:
                grad_f = blocks.fgrad_half(rows, g, x)
        if grad_f is not None and ctx.group is not None:
            grad_f = _reduce_partial(grad_f, ctx.group)
        return grad_f, grad_x, None, None, None, None


class GRFFeatureOperator(LinearOperator):
    """Lazy ``Phi[rows]`` (or its transpose) with ``Phi = sum_l f[l] M_l``.

    ``group``: a torch.distributed group (or True for the default group) when
    the Phi blocks are a row shard; products with Phi^T are then all-reduced."""

    def __init__(self, blocks: PhiBlocks, modulator: torch.Tensor, rows: Optional[torch.Tensor] = None,
                 transposed: bool = False, group=None):
        self.blocks, self.modulator, self.rows, self.transposed, self.group = blocks, modulator, rows, transposed, group
        super().__init__(modulator)

    @property
    def device(self):
        return self.blocks.device

    def _n_sel(self):
        return self.blocks.n_rows if self.rows is None else int(self.rows.numel())

    def _size(self):
        n, m = self._n_sel(), self.blocks.n_cols
        return torch.Size((m, n) if self.transposed else (n, m))

    def _matmul(self, rhs):
        squeeze = rhs.dim() == 1
        x = rhs[:, None] if squeeze else rhs
        y = PhiApply.apply(self.modulator, x.to(torch.float32), self.blocks, self.rows, self.transposed, self.group)
        return y[:, 0] if squeeze else y

    def _transpose_nonbatch(self):
        return GRFFeatureOperator(self.blocks, self.modulator, self.rows, not self.transposed, self.group)

    def _select_rows(self, index):
        if self.transposed:
            raise IndexError("row selection is defined on Phi, not on Phi^T")
        if isinstance(index, slice):
            index = torch.arange(self._n_sel(), device=self.device)[index]
        index = torch.as_tensor(index, device=self.device).long().flatten()
        rows = index if self.rows is None else self.rows.long()[index]
        self.blocks.check_ids(rows)      # IndexError as ``phi[idx]`` in the reference, not a silent zero row
        return GRFFeatureOperator(self.blocks, self.modulator, rows.to(torch.int32), False, self.group)

    def __getitem__(self, index):
        if isinstance(index, tuple):
            rows, cols = index
            if not (isinstance(cols, slice) and cols == slice(None)):
                return super().__getitem__(index)
            if isinstance(rows, slice) and rows == slice(None):
                return self
            return self._select_rows(rows)
        return self._select_rows(index)

    def matmul(self, other):
        if isinstance(other, GRFFeatureOperator) and (not self.transposed) and other.transposed \
                and other.blocks is self.blocks:
            return GRFKernelOperator(self.blocks, self.modulator, self.rows, other.rows, self.group)
        return super().matmul(other)

    __matmul__ = matmul

    def to_dense(self):
        """Dense Phi[rows] (or its transpose): identity columns in chunks through the product, for tests and
        small row sets only -- nothing on the kernel path densifies."""
        n = self.shape[-1]
        return _dense_by_column_chunks(self._matmul, n, self.device)

    def row_dots_with(self, other: "GRFFeatureOperator") -> torch.Tensor:
        """sum(self * other, -1) for two row selections of the same Phi (the reference's diag=True branch,
        sparse_grf_kernel.py:55-57) without densifying either: per-pair sparse dot products on the device,
        differentiable w.r.t. the modulator."""
        if self.transposed or other.transposed or other.blocks is not self.blocks:
            raise ValueError("row_dots_with: both operands must be row selections of the same Phi")
        d = PhiRowDots.apply(self.modulator, self.blocks, self.rows, other.rows)
        if self.group is not None:
            d = _reduce_partial(d.contiguous(), self.group)
        return d


def _dense_by_column_chunks(matmul, n: int, device, chunk: int = 1024) -> torch.Tensor:
    cols = []
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        eye = torch.zeros((n, c1 - c0), dtype=torch.float32, device=device)
        eye[c0:c1] = torch.eye(c1 - c0, dtype=torch.float32, device=device)
        cols.append(matmul(eye))
    return torch.cat(cols, dim=1) if cols else matmul(torch.zeros((n, 0), dtype=torch.float32, device=device))


class PhiRowDots(torch.autograd.Function):
    """d[i] = <Phi_f[x1[i]], Phi_f[x2[i]]> = sum_l f[l] * D12[i, l], D12[i, l] = <M_l[x1[i]], Phi_f[x2[i]]>.
    dd[i]/df[l] = D12[i, l] + D21[i, l] (``grf_phi_row_dots`` with the roles swapped)."""

    @staticmethod
    def forward(ctx, f, blocks: PhiBlocks, x1, x2):
        fd = f.detach()
        d12 = blocks.row_dots(fd, x1, x2)
        ctx.blocks, ctx.x1, ctx.x2 = blocks, x1, x2
        ctx.same = x1 is x2
        ctx.save_for_backward(fd, d12)
        return d12 @ fd.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        fd, d12 = ctx.saved_tensors
        d21 = d12 if ctx.same else ctx.blocks.row_dots(fd, ctx.x2, ctx.x1)
        return (g.to(torch.float32) @ (d12 + d21)).to(fd.dtype), None, None, None


class GRFKernelOperator(LinearOperator):
    """Lazy ``K[x1, x2] = Phi[x1] Phi[x2]^T``; ``_matmul`` is the fused two-pass CUDA matvec."""

    def __init__(self, blocks: PhiBlocks, modulator: torch.Tensor, x1=None, x2=None, group=None):
        self.blocks, self.modulator, self.x1, self.x2, self.group = blocks, modulator, x1, x2, group
        super().__init__(modulator)

    @property
    def device(self):
        return self.blocks.device

    def _size(self):
        n1 = self.blocks.n_rows if self.x1 is None else int(self.x1.numel())
        n2 = self.blocks.n_rows if self.x2 is None else int(self.x2.numel())
        return torch.Size((n1, n2))

    def _matmul(self, rhs):
        squeeze = rhs.dim() == 1
        v = (rhs[:, None] if squeeze else rhs).to(torch.float32)
        u = PhiApply.apply(self.modulator, v, self.blocks, self.x2, True, self.group)
        y = PhiApply.apply(self.modulator, u, self.blocks, self.x1, False, self.group)
        return y[:, 0] if squeeze else y

    def _transpose_nonbatch(self):
        return GRFKernelOperator(self.blocks, self.modulator, self.x2, self.x1, self.group)

    def plan(self, t: int, merged: bool = True):
        """A fixed-shape fast path for CG: ``plan(v, out)`` = K v with the current modulator value."""
        return self.blocks.plan(self.modulator.detach(), t, x1=self.x1, x2=self.x2, group=self.group, merged=merged)

    def solve(self, rhs, sigma2: float = 0.0, tolerance: float = 1e-2, max_iter: int = 1000, eps: float = 1e-10,
              return_info: bool = False):
        """(K + sigma2 I)^-1 rhs by conjugate gradients (no gradient tracking)."""
        from .cg import linear_cg_fused

        with torch.no_grad():
            rhs2 = rhs[:, None] if rhs.dim() == 1 else rhs
            # single GPU or row-sharded: the fused vector kernels either way (sharded: dot products all-reduced)
            out, info = linear_cg_fused(self.plan(rhs2.shape[1]), rhs2, sigma2, tolerance, max_iter, eps,
                                        return_info=True)
            out = out[:, 0] if rhs.dim() == 1 else out
        return (out, info) if return_info else out

    def _bilinear_derivative(self, left_vecs, right_vecs):
        """Upstream protocol: gradients of sum(left * (K right)) w.r.t. the representation (= the modulator)."""
        if left_vecs.dim() == 1:
            left_vecs, right_vecs = left_vecs[:, None], right_vecs[:, None]
        grad = self.blocks.fgrad(self.modulator.detach(), left_vecs, right_vecs, self.x1, self.x2)
        return (_reduce_partial(grad, self.group),)

    def diagonal(self, offset=0, dim1=-2, dim2=-1):
        """diag(K[x1, x2])[i] = <Phi[x1[i]], Phi[x2[i]]>: per-pair sparse dot products, nothing densified."""
        if offset != 0:
            return self.to_dense().diagonal(offset)
        n = min(self._size())
        x1 = None if self.x1 is None and self.x2 is None else (
            torch.arange(n, device=self.device, dtype=torch.int32) + self.blocks.row_lo if self.x1 is None
            else self.x1[:n])
        x2 = x1 if self.x2 is self.x1 else (
            None if x1 is None else (torch.arange(n, device=self.device, dtype=torch.int32) + self.blocks.row_lo
                                     if self.x2 is None else self.x2[:n]))
        d = PhiRowDots.apply(self.modulator, self.blocks, x1, x2)
        return _reduce_partial(d.contiguous(), self.group) if self.group is not None else d

    def to_dense(self):
        """Dense K[x1, x2] through identity columns in chunks (tests / small index sets)."""
        return _dense_by_column_chunks(self._matmul, self._size()[1], self.device)

