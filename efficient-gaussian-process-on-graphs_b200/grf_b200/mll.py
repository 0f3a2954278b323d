"""Exact-GP marginal log-likelihood gradients with CG + stochastic trace estimation.

This is what the reference gets from upstream GPyTorch when its training loops call
``loss = -mll(model(X), y); loss.backward()`` (run_scaling_experiment.py:603-612,
graph_bo/utils/bo_utils.py:244-251) with ``max_cholesky_size = 0`` and
``num_trace_samples`` probe vectors: one batched CG solve
``(K + s2 I)^-1 [y | z_1 .. z_p]`` and, per hyper-parameter,

    dL/dtheta = 1/2 a^T (dK/dtheta) a  -  1/2 * mean_j  w_j^T (dK/dtheta) z_j ,   a = Khat^-1 y, w_j = Khat^-1 z_j.

Here the solve is the fused CUDA CG on the Phi(Phi^T V) matvec and, for the modulator f
(K = Phi_f Phi_f^T, Phi_f = sum_l f_l M_l), the bilinear forms are the per-length reductions of
``grf_phi_fgrad``; the chain to the raw parameters (softplus constraints, the diffusion
formula) is left to autograd through ``modulator_vector`` / ``likelihood.noise``.
gpytorch is not installed in this image, so parity is against a float64 dense evaluation
(tests/test_gpu_gp.py), not against upstream.
"""

from __future__ import annotations

from typing import Optional

import torch


def neg_mll_backward(kernel, likelihood, x_train: torch.Tensor, y_train: torch.Tensor, num_probes: int = 16,
                     cg_tolerance: float = 1e-2, max_cg_iterations: int = 1000, probes: Optional[torch.Tensor] = None,
                     generator: Optional[torch.Generator] = None, cg_eps: float = 1e-10):
    """Accumulate d(-mll/n)/dtheta into ``.grad`` of the kernel's and the likelihood's parameters.

    Returns a dict with the data-fit term, the CG iteration count and the modulator gradient.
    ``probes`` ([n, p]) overrides the Gaussian probe vectors (used by the exactness test)."""
    blocks = kernel.phi_blocks
    dev = blocks.device
    x = x_train.long().flatten().to(dev)
    y = y_train.to(dev).to(torch.float32).reshape(-1)
    n = y.numel()
    f = kernel.modulator_vector            # differentiable function of the raw parameters
    noise = likelihood.noise               # idem
    s2 = float(noise.detach())
    if probes is None:
        probes = torch.randn(n, num_probes, device=dev, generator=generator)
    probes = probes.to(dev).to(torch.float32)
    p = probes.shape[1]
    rhs = torch.cat([y[:, None], probes], dim=1).contiguous()

    K = kernel(x, x)
    sol, info = K.solve(rhs, s2, tolerance=cg_tolerance, max_iter=max_cg_iterations, eps=cg_eps, return_info=True)
    alpha, w = sol[:, :1].contiguous(), sol[:, 1:].contiguous()

    fd = f.detach()
    with torch.no_grad():
        # d/df_l of a^T K a and of sum_j w_j^T K z_j
        g_fit = blocks.fgrad(fd, alpha, alpha, x1=x, x2=x)
        g_tr = blocks.fgrad(fd, w, probes, x1=x, x2=x)
        dL_df = 0.5 * g_fit - 0.5 * g_tr / p
        dL_ds2 = 0.5 * float((alpha * alpha).sum()) - 0.5 * float((w * probes).sum()) / p
        datafit = 0.5 * float((y * alpha[:, 0]).sum())
    # loss = -mll / n  (gpytorch's ExactMarginalLogLikelihood divides by the number of data points)
    f.backward((-dL_df / n).to(f.dtype).to(f.device), retain_graph=True)
    noise.backward(torch.as_tensor([-dL_ds2 / n], dtype=noise.dtype, device=noise.device).reshape(noise.shape))
    return {"datafit": datafit, "cg_iterations": info["iterations"], "grad_modulator": (-dL_df / n).cpu(),
            "grad_noise": -dL_ds2 / n}
