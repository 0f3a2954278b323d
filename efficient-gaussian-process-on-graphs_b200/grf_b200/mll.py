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


def lanczos_logdet(matmul, probes: torch.Tensor, iterations: int = 1) -> float:
    """log|A| by stochastic Lanczos quadrature: mean_j ||z_j||^2 e_1^T log(T_j) e_1 with T_j the
    ``iterations`` x ``iterations`` Lanczos tridiagonal of A started at z_j / ||z_j||.

    The reference sets ``max_lanczos_quadrature_iterations = 1`` (run_scaling_experiment.py:133,
    graph_bo/utils/gpytorch_config.py:8), for which T_j is the Rayleigh quotient z^T A z / z^T z --
    one batched matvec.  More iterations run the three-term recurrence with full
    re-orthogonalisation (small k)."""
    z = probes.to(torch.float32)
    norms2 = (z * z).sum(dim=0)
    q = z / norms2.sqrt()
    qs = [q]
    alphas, betas = [], []
    beta_prev, q_prev = None, None
    for k in range(iterations):
        w = matmul(q)
        if q_prev is not None:
            w = w - beta_prev * q_prev
        alpha = (w * q).sum(dim=0)
        w = w - alpha * q
        for qq in qs:                                   # re-orthogonalise (k is small)
            w = w - (w * qq).sum(dim=0) * qq
        alphas.append(alpha)
        beta = w.norm(dim=0)
        if k + 1 < iterations:
            betas.append(beta)
            q_prev, beta_prev = q, beta
            q = w / beta.clamp_min(1e-20)
            qs.append(q)
    a = torch.stack(alphas, dim=1).double().cpu()                      # [p, k]
    p, k = a.shape
    T = torch.diag_embed(a)
    if k > 1:
        b = torch.stack(betas, dim=1).double().cpu()                   # [p, k-1]
        T = T + torch.diag_embed(b, offset=1) + torch.diag_embed(b, offset=-1)
    evals, evecs = torch.linalg.eigh(T)
    quad = (evecs[:, 0, :] ** 2 * torch.log(evals.clamp_min(1e-30))).sum(dim=1)   # e_1^T log(T) e_1
    return float((norms2.double().cpu() * quad).mean())


def neg_mll_backward(kernel, likelihood, x_train: torch.Tensor, y_train: torch.Tensor, num_probes: int = 16,
                     cg_tolerance: float = 1e-2, max_cg_iterations: int = 1000, probes: Optional[torch.Tensor] = None,
                     generator: Optional[torch.Generator] = None, cg_eps: float = 1e-10,
                     lanczos_iterations: int = 1):
    """Accumulate d(-mll/n)/dtheta into ``.grad`` of the kernel's and the likelihood's parameters.

    Returns a dict with the loss estimate (-mll/n, log-determinant by ``lanczos_iterations`` steps of
    stochastic Lanczos quadrature on the same probes; 0 skips it), the CG iteration count and the gradients.
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
        loss = None
        if lanczos_iterations > 0:
            plan = K.plan(p)
            logdet = lanczos_logdet(lambda v: plan(v.contiguous()) + s2 * v, probes, lanczos_iterations)
            loss = (datafit + 0.5 * logdet + 0.5 * n * 1.8378770664093453) / n
    # loss = -mll / n  (gpytorch's ExactMarginalLogLikelihood divides by the number of data points)
    f.backward((-dL_df / n).to(f.dtype).to(f.device), retain_graph=True)
    noise.backward(torch.as_tensor([-dL_ds2 / n], dtype=noise.dtype, device=noise.device).reshape(noise.shape))
    return {"loss": loss, "datafit": datafit, "cg_iterations": info["iterations"],
            "grad_modulator": (-dL_df / n).cpu(), "grad_noise": -dL_ds2 / n}
