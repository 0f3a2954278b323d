from .sparse_grf_kernel import SparseGRFKernel
from .sparse_diffusion_kernel import SparseDiffusionKernel, diffusion_modulator_torch

__all__ = ["SparseGRFKernel", "SparseDiffusionKernel", "diffusion_modulator_torch"]
