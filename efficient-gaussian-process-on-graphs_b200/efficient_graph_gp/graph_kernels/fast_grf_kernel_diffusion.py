"""Drop-in for ``graph_kernels/fast_grf_kernel_diffusion.py:7-21``."""

import numpy as np

from efficient_graph_gp.modulation_functions import diffusion_modulator
from efficient_graph_gp.graph_kernels.fast_grf_kernel_general import fast_general_grf_kernel


def fast_diffusion_grf_kernel(adj_matrix, walks_per_node=50, p_halt=0.1, max_walk_length=10, beta=1.0):
    """GRF estimate of the diffusion kernel: modulator f_l = (-beta)^l / (2^l l!)."""
    modulator_vector = np.array([diffusion_modulator(step, beta) for step in range(max_walk_length)])
    return fast_general_grf_kernel(adj_matrix, modulator_vector, walks_per_node, p_halt, max_walk_length)
