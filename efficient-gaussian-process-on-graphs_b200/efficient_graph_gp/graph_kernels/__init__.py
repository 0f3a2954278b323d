from .fast_grf_kernel_general import fast_general_grf_kernel
from .fast_grf_kernel_diffusion import fast_diffusion_grf_kernel
from .utils import get_normalized_laplacian

__all__ = ["get_normalized_laplacian", "fast_general_grf_kernel", "fast_diffusion_grf_kernel"]
