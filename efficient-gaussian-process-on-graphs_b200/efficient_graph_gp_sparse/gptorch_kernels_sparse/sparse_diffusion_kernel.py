"""Drop-in for ``gptorch_kernels_sparse/sparse_diffusion_kernel.py:6-96``:
``diffusion_modulator_torch`` and ``SparseDiffusionKernel`` (learnable ``beta`` and
``sigma_f`` under Positive constraints, modulator ``sigma_f (-beta)^l / (2^l l!)``), on the
fused Phi blocks; gradients reach ``raw_beta`` / ``raw_sigma_f`` through the modulator."""

import torch

from grf_b200.gp_compat import Kernel, Positive
from grf_b200.operators import GRFFeatureOperator
from ._fused import fused_blocks


def diffusion_modulator_torch(length: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    """(-beta)^length / (2^length * Gamma(length + 1)), dtype / device following ``beta``."""
    length = length.to(dtype=beta.dtype, device=beta.device)
    numerator = torch.pow(-beta, length)
    denominator = torch.pow(torch.tensor(2.0, dtype=beta.dtype, device=beta.device), length)
    denominator = denominator * torch.exp(torch.lgamma(length + 1.0))
    return numerator / denominator


class SparseDiffusionKernel(Kernel):
    def __init__(self, max_walk_length, step_matrices_torch, **kwargs):
        super().__init__(**kwargs)
        self.register_parameter(name="raw_beta", parameter=torch.nn.Parameter(torch.tensor(1.0)))
        self.register_constraint("raw_beta", Positive())
        self.register_parameter(name="raw_sigma_f", parameter=torch.nn.Parameter(torch.tensor(1.0)))
        self.register_constraint("raw_sigma_f", Positive())
        self.step_matrices = step_matrices_torch
        self.max_walk_length = max_walk_length
        self._blocks = None

    @property
    def beta(self):
        return self.raw_beta_constraint.transform(self.raw_beta)

    @property
    def sigma_f(self):
        return self.raw_sigma_f_constraint.transform(self.raw_sigma_f)

    @property
    def modulator_vector(self):
        walk_lengths = torch.arange(self.max_walk_length, dtype=self.raw_beta.dtype, device=self.raw_beta.device)
        return self.sigma_f * diffusion_modulator_torch(walk_lengths, self.beta)

    @property
    def phi_blocks(self):
        if self._blocks is None:
            self._blocks = fused_blocks(self.step_matrices)
        return self._blocks

    def forward(self, x1_idx=None, x2_idx=None, diag=False, **params):
        phi = self._get_feature_matrix()
        phi_x1 = phi if x1_idx is None else phi[x1_idx.long().flatten()]
        phi_x2 = phi if x2_idx is None else phi[x2_idx.long().flatten()]
        if diag:
            return phi_x1.row_dots_with(phi_x2)
        return phi_x1 @ phi_x2.transpose(-1, -2)

    def _get_feature_matrix(self):
        return GRFFeatureOperator(self.phi_blocks, self.modulator_vector.to(self.phi_blocks.device))
