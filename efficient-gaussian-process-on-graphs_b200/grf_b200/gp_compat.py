"""GPyTorch surface the reference's kernels and model are written against.

``gpytorch`` is used when importable (``gpytorch.kernels.Kernel``,
``gpytorch.constraints.Positive``, ``gpytorch.models.ExactGP``,
``GaussianLikelihood``, ``settings.cg_tolerance``).  It is not installed in this
image, so minimal stand-ins with the same attribute names and parameter
transforms are provided; they carry no numerics of the hot path.
"""

from __future__ import annotations

import math

import torch

try:  # pragma: no cover - not installed in the build image
    import gpytorch as _gpytorch

    HAVE_GPYTORCH = True
except Exception:
    _gpytorch = None
    HAVE_GPYTORCH = False


class _Positive:
    """softplus transform, like gpytorch.constraints.Positive."""

    def transform(self, raw):
        return torch.nn.functional.softplus(raw)

    def inverse_transform(self, value):
        value = torch.as_tensor(value)
        return value + torch.log(-torch.expm1(-value))


class _GreaterThan(_Positive):
    def __init__(self, lower_bound):
        self.lower_bound = float(lower_bound)

    def transform(self, raw):
        return torch.nn.functional.softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return super().inverse_transform(torch.as_tensor(value) - self.lower_bound)


class _Kernel(torch.nn.Module):
    """Stand-in for gpytorch.kernels.Kernel: parameter / constraint registration and
    ``kernel(x1, x2)`` -> ``forward`` (lazy operator out)."""

    def __init__(self, **kwargs):
        super().__init__()

    def register_parameter(self, name, parameter=None, param=None):
        # gpytorch.Module.register_parameter(name, parameter); torch's keyword is `param`
        super().register_parameter(name, parameter if parameter is not None else param)

    def register_constraint(self, param_name, constraint):
        setattr(self, param_name + "_constraint", constraint)

    def __call__(self, x1=None, x2=None, diag=False, **params):
        return self.forward(x1, x1 if x2 is None else x2, diag=diag, **params)


class _GaussianLikelihood(torch.nn.Module):
    """Homoskedastic Gaussian noise with gpytorch's defaults (noise > 1e-4, raw_noise = 0)."""

    def __init__(self, noise_constraint=None):
        super().__init__()
        self.raw_noise = torch.nn.Parameter(torch.zeros(1))
        self.raw_noise_constraint = noise_constraint or _GreaterThan(1e-4)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        with torch.no_grad():
            self.raw_noise.copy_(self.raw_noise_constraint.inverse_transform(
                torch.as_tensor(value, dtype=self.raw_noise.dtype)).reshape(1))


class _ExactGP(torch.nn.Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.train_inputs = (train_inputs,)
        self.train_targets = train_targets
        self.likelihood = likelihood


class _Setting:
    """``settings.<name>._global_value = x`` / ``.value()``, as the reference configures gpytorch
    (run_scaling_experiment.py:129-135, graph_bo/utils/gpytorch_config.py:6-10)."""

    def __init__(self, default):
        self._global_value = default

    def value(self):
        return self._global_value


class _Settings:
    cg_tolerance = _Setting(1.0)                       # gpytorch default
    max_cg_iterations = _Setting(1000)
    max_cholesky_size = _Setting(800)
    max_lanczos_quadrature_iterations = _Setting(20)
    num_trace_samples = _Setting(10)
    min_preconditioning_size = _Setting(2000)


if HAVE_GPYTORCH:  # pragma: no cover
    Kernel = _gpytorch.kernels.Kernel
    Positive = _gpytorch.constraints.Positive
    ExactGP = _gpytorch.models.ExactGP
    GaussianLikelihood = _gpytorch.likelihoods.GaussianLikelihood
    settings = _gpytorch.settings
else:
    Kernel, Positive, ExactGP, GaussianLikelihood, settings = _Kernel, _Positive, _ExactGP, _GaussianLikelihood, _Settings
