"""Lazy linear-operator protocol the GPyTorch-facing drop-ins are written against.

The reference builds its kernels on the third-party ``linear_operator`` package
(``sparse_lo.py:2``, ``sparse_grf_kernel.py:3``).  When that package is
importable its ``LinearOperator`` is used as the base class, so the drop-ins
plug into GPyTorch unchanged.  It is not installed in this image, so this
module also carries a small stand-in that implements exactly the slice of the
protocol the reference's call sites use:

    scalar * op, sum(ops), op[idx], op[idx, :], op.T / op.transpose(-1, -2),
    op @ op, op @ tensor, tensor @ op, op + c * Identity, op._matmul,
    op.to_dense(), op.diagonal(), (op * op).sum(-1)

Host-side plumbing only: no arithmetic on the hot path happens here.
"""

from __future__ import annotations

import torch

try:  # pragma: no cover - not installed in the build image
    from linear_operator.operators import LinearOperator as _UpstreamLinearOperator

    HAVE_UPSTREAM = True
except Exception:  # ImportError or a broken install
    _UpstreamLinearOperator = None
    HAVE_UPSTREAM = False


class _StandInLinearOperator:
    """Minimal lazy operator: subclasses implement ``_matmul``, ``_size``,
    ``_transpose_nonbatch`` (the three methods ``sparse_lo.py`` implements)."""

    def __init__(self, *args, **kwargs):
        self._args = args
        self._kwargs = kwargs

    # ---- to be provided by subclasses ------------------------------------
    def _matmul(self, rhs):
        raise NotImplementedError

    def _size(self):
        raise NotImplementedError

    def _transpose_nonbatch(self):
        return TransposedLinearOperator(self)

    # ---- shape -----------------------------------------------------------
    @property
    def shape(self):
        return torch.Size(self._size())

    def size(self, dim=None):
        s = self.shape
        return s if dim is None else s[dim]

    def dim(self):
        return len(self.shape)

    @property
    def device(self):
        for a in self._args:
            if hasattr(a, "device"):
                return a.device
        return torch.device("cpu")

    @property
    def dtype(self):
        return torch.float32

    # ---- algebra ---------------------------------------------------------
    def matmul(self, other):
        if isinstance(other, _StandInLinearOperator):
            return MatmulLinearOperator(self, other)
        if other.dim() == 1:
            return self._matmul(other[:, None])[:, 0]
        return self._matmul(other)

    __matmul__ = matmul

    def __rmatmul__(self, other):
        # tensor @ op == (op^T @ tensor^T)^T
        if other.dim() == 1:
            return self._transpose_nonbatch()._matmul(other[:, None])[:, 0]
        return self._transpose_nonbatch()._matmul(other.transpose(-1, -2)).transpose(-1, -2)

    def transpose(self, dim1, dim2):
        nd = self.dim()
        if {dim1 % nd, dim2 % nd} != {nd - 2, nd - 1}:
            raise ValueError("only the last two dimensions can be transposed")
        return self._transpose_nonbatch()

    def t(self):
        return self._transpose_nonbatch()

    @property
    def T(self):
        return self._transpose_nonbatch()

    @property
    def mT(self):
        return self._transpose_nonbatch()

    def __mul__(self, other):
        if isinstance(other, _StandInLinearOperator):
            return DenseLinearOperator(self.to_dense() * other.to_dense())
        return ConstantMulLinearOperator(self, other)

    __rmul__ = __mul__

    def __add__(self, other):
        if isinstance(other, (int, float)) and other == 0:
            return self
        if isinstance(other, _StandInLinearOperator):
            return SumLinearOperator(self, other)
        return DenseLinearOperator(self.to_dense() + other)

    __radd__ = __add__

    def __getitem__(self, index):
        if isinstance(index, tuple):
            rows, cols = index
            out = self if _is_full_slice(rows) else RowSelectLinearOperator(self, rows)
            if not _is_full_slice(cols):
                out = RowSelectLinearOperator(out._transpose_nonbatch(), cols)._transpose_nonbatch()
            return out
        return RowSelectLinearOperator(self, index)

    def add_diagonal(self, diag):
        n = self.shape[-1]
        return SumLinearOperator(self, ConstantMulLinearOperator(IdentityLinearOperator(n, device=self.device), diag))

    def add_jitter(self, jitter_val=1e-3):
        return self.add_diagonal(torch.as_tensor(jitter_val, device=self.device))

    # ---- evaluation ------------------------------------------------------
    def to_dense(self):
        n = self.shape[-1]
        return self._matmul(torch.eye(n, dtype=torch.float32, device=self.device))

    def evaluate(self):
        return self.to_dense()

    def diagonal(self, offset=0, dim1=-2, dim2=-1):
        return self.to_dense().diagonal(offset, dim1, dim2)

    def sum(self, dim=None):
        d = self.to_dense()
        return d.sum() if dim is None else d.sum(dim)


def _is_full_slice(idx) -> bool:
    return isinstance(idx, slice) and idx == slice(None)


LinearOperator = _UpstreamLinearOperator if HAVE_UPSTREAM else _StandInLinearOperator


# The composite operators below are only defined for the stand-in; with the
# upstream package its own ConstantMul / Sum / Matmul / Interpolated operators
# are produced by the base-class algebra.
class DenseLinearOperator(_StandInLinearOperator):
    def __init__(self, tensor):
        super().__init__(tensor)
        self.tensor = tensor

    def _matmul(self, rhs):
        return self.tensor @ rhs

    def _size(self):
        return self.tensor.shape

    def _transpose_nonbatch(self):
        return DenseLinearOperator(self.tensor.transpose(-1, -2))

    def to_dense(self):
        return self.tensor


class IdentityLinearOperator(_StandInLinearOperator):
    def __init__(self, diag_shape, device=None, dtype=torch.float32):
        super().__init__()
        self.n = int(diag_shape)
        self._device = torch.device(device) if device is not None else torch.device("cpu")

    @property
    def device(self):
        return self._device

    def _matmul(self, rhs):
        return rhs

    def _size(self):
        return (self.n, self.n)

    def _transpose_nonbatch(self):
        return self


class TransposedLinearOperator(_StandInLinearOperator):
    def __init__(self, base):
        super().__init__(base)
        self.base = base

    def _matmul(self, rhs):
        raise NotImplementedError(f"{type(self.base).__name__} does not define a transposed product")

    def _size(self):
        s = self.base.shape
        return (s[1], s[0])

    def _transpose_nonbatch(self):
        return self.base


class ConstantMulLinearOperator(_StandInLinearOperator):
    def __init__(self, base, constant):
        super().__init__(base)
        self.base = base
        self.constant = constant

    @property
    def device(self):
        return self.base.device

    def _matmul(self, rhs):
        return self.constant * self.base._matmul(rhs)

    def _size(self):
        return self.base.shape

    def _transpose_nonbatch(self):
        return ConstantMulLinearOperator(self.base._transpose_nonbatch(), self.constant)


class SumLinearOperator(_StandInLinearOperator):
    def __init__(self, *ops):
        flat = []
        for o in ops:
            flat.extend(o.ops if isinstance(o, SumLinearOperator) else [o])
        super().__init__(*flat)
        self.ops = flat

    @property
    def device(self):
        return self.ops[0].device

    def _matmul(self, rhs):
        out = self.ops[0]._matmul(rhs)
        for o in self.ops[1:]:
            out = out + o._matmul(rhs)
        return out

    def _size(self):
        return self.ops[0].shape

    def _transpose_nonbatch(self):
        return SumLinearOperator(*[o._transpose_nonbatch() for o in self.ops])


class MatmulLinearOperator(_StandInLinearOperator):
    def __init__(self, left, right):
        super().__init__(left, right)
        self.left, self.right = left, right

    @property
    def device(self):
        return self.left.device

    def _matmul(self, rhs):
        return self.left._matmul(self.right._matmul(rhs))

    def _size(self):
        return (self.left.shape[0], self.right.shape[1])

    def _transpose_nonbatch(self):
        return MatmulLinearOperator(self.right._transpose_nonbatch(), self.left._transpose_nonbatch())


class RowSelectLinearOperator(_StandInLinearOperator):
    """``op[idx]``: gather after the product; its transpose scatters before it
    (what upstream's InterpolatedLinearOperator does for the reference)."""

    def __init__(self, base, index, transposed=False):
        super().__init__(base)
        self.base = base
        if isinstance(index, slice):
            index = torch.arange(base.shape[0], device=base.device)[index]
        self.index = torch.as_tensor(index, device=base.device).long().flatten()
        self.transposed = transposed

    @property
    def device(self):
        return self.base.device

    def _matmul(self, rhs):
        if not self.transposed:
            return self.base._matmul(rhs)[self.index]
        full = torch.zeros((self.base.shape[0], rhs.shape[1]), dtype=rhs.dtype, device=rhs.device)
        full.index_add_(0, self.index, rhs)
        return self.base._transpose_nonbatch()._matmul(full)

    def _size(self):
        n, m = self.index.numel(), self.base.shape[1]
        return (m, n) if self.transposed else (n, m)

    def _transpose_nonbatch(self):
        return RowSelectLinearOperator(self.base, self.index, not self.transposed)
