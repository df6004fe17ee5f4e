"""logposterior with the reference's signature (sub_modules/loss_modules.py:5-13).

``logposterior(loglikelihood, params_optimization)`` returns f(theta, *args) = NLL + sum(theta)
(+ ridge_alpha * sum(exp(theta)^2) with loss_ridge_regression).  The returned object also carries the analytic
gradient (``.grad``, ``.value_and_grad``), which is what ``jit(grad(func, 0))`` is in the reference's scripts;
``stopro_b200.solver.optimizers`` picks it up, so one factorisation serves both value and gradient.
"""
import numpy as np


class LogPosterior:
    def __init__(self, loglikelihood, params_optimization):
        self.loglikelihood = loglikelihood
        self.ridge = bool(params_optimization.get("loss_ridge_regression"))
        self.ridge_alpha = params_optimization.get("ridge_alpha", 0.0) if self.ridge else 0.0
        self.model = getattr(loglikelihood, "__self__", None)

    def _prior(self, theta):
        v = np.sum(theta)
        if self.ridge:
            v = v + self.ridge_alpha * np.sum(np.square(np.exp(theta)))
        return v

    def __call__(self, theta, *args):
        theta = np.asarray(theta, dtype=np.float64)
        return self.loglikelihood(theta, *args) + self._prior(theta)

    def value_and_grad(self, theta, *args):
        if self.model is None or not hasattr(self.model, "value_and_grad"):
            raise TypeError("analytic gradient needs a stopro_b200 GP model's trainingFunction_all")
        theta = np.asarray(theta, dtype=np.float64)
        nll, g = self.model.value_and_grad(theta, *args)
        g = g + 1.0
        if self.ridge:
            g = g + 2.0 * self.ridge_alpha * np.square(np.exp(theta))
        return nll + self._prior(theta), g

    def grad(self, theta, *args):
        return self.value_and_grad(theta, *args)[1]


def logposterior(loglikelihood, params_optimization):
    return LogPosterior(loglikelihood, params_optimization)


def hessian(f):
    """The reference builds ``jit(jacfwd(jacrev(f)))`` in every script but never evaluates it
    (sub_modules/loss_modules.py:16-20); construction must not fail, evaluation is not on the path."""

    def _hessian(*args, **kwargs):
        raise NotImplementedError("the Hessian of the log posterior is not on the B200 hot path")

    return _hessian
