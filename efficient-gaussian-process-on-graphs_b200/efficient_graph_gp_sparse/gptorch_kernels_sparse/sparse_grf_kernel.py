"""Drop-in for ``gptorch_kernels_sparse/sparse_grf_kernel.py:5-62``.

Same constructor, the same learnable ``raw_modulator_vector ~ randn(L)``
(:14-17), ``modulator_vector``, ``forward(x1_idx, x2_idx, diag)`` and
``_get_feature_matrix()``.  What the reference assembles lazily out of 2L
SparseLinearOperators, ConstantMul / Sum / Interpolated / Matmul operators
(:51-62) is here ONE operator over the fused Phi blocks: a product with
``K[x1, x2] = Phi[x1] Phi[x2]^T`` is two CUDA launches, and the modulator
gradient is the per-length reduction of ``grf_phi_fgrad``.
"""

import torch

from grf_b200.gp_compat import Kernel
from grf_b200.operators import GRFFeatureOperator
from ._fused import fused_blocks


class SparseGRFKernel(Kernel):
    def __init__(self, max_walk_length, step_matrices_torch, **kwargs):
        super().__init__(**kwargs)
        self.register_parameter(
            name="raw_modulator_vector",
            parameter=torch.nn.Parameter(torch.randn(max_walk_length))
        )
        self.step_matrices = step_matrices_torch
        self.max_walk_length = max_walk_length
        self._blocks = None

    @property
    def modulator_vector(self):
        return self.raw_modulator_vector

    @property
    def phi_blocks(self):
        if self._blocks is None:
            self._blocks = fused_blocks(self.step_matrices)
            if self._blocks.n_steps != self.max_walk_length:
                raise ValueError("The length of the modulator vector must be equal to the max_walk_length.")
        return self._blocks

    def forward(self, x1_idx=None, x2_idx=None, diag=False, **params):
        """K[x1, x2] with K = Phi Phi^T, returned lazily."""
        phi = self._get_feature_matrix()
        if x1_idx is not None:
            x1_idx = x1_idx.long().flatten()
            phi_x1 = phi[x1_idx]
        else:
            phi_x1 = phi
        if x2_idx is not None:
            x2_idx = x2_idx.long().flatten()
            phi_x2 = phi[x2_idx]
        else:
            phi_x2 = phi
        if diag:
            # diag(A B^T) = sum(A * B, -1) as per-pair sparse dot products (the reference densifies both row sets)
            return phi_x1.row_dots_with(phi_x2)
        return phi_x1 @ phi_x2.transpose(-1, -2)

    def _get_feature_matrix(self):
        """Lazy Phi = sum_l modulator[l] * M_l (the i-th row is the GRF vector of node i)."""
        mod = self.modulator_vector
        return GRFFeatureOperator(self.phi_blocks, mod.to(self.phi_blocks.device))
