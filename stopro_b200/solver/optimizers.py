"""optimize_by_adam with the reference's signature and stopping rules (solver/optimizers.py:94-263).

What is kept: normalisation of loss and gradient by the number of training points (:131-139), "loss before optimize"
as first loss entry and NaN guard (:239-246), the Adam update (optax.adam defaults: b1 = 0.9, b2 = 0.999,
eps = 1e-8), history lists of theta / loss / gradient norm, the two-in-a-row plateau stop on |delta loss| < eps
(:224-229) and the NaN guards on gradient and theta (:182-183, :230-232).  What differs: value and gradient come from
ONE factorisation (``f.value_and_grad`` of stopro_b200's logposterior) instead of func + dfunc assembling and
factorising K twice per step (:148-150).  The optional scipy pre-stage is delegated to scipy.optimize.minimize.
"""
import numpy as np


def _value_and_grad(f, df):
    if hasattr(f, "value_and_grad"):
        return f.value_and_grad
    if df is None:
        raise ValueError("a gradient function is required")
    return lambda p, *args: (f(p, *args), df(p, *args))


def optimize_by_adam(f, df, hf, init, params_optimization, *args):
    maxiter_GD = params_optimization["maxiter_GD"]
    lr = params_optimization["lr"]
    eps = params_optimization["eps"]
    maxiter_scipy = params_optimization.get("maxiter_scipy", [0])
    method_scipy = params_optimization.get("method_scipy", [])
    print_process = params_optimization.get("print_process", False)
    index_fixed = params_optimization.get("index_fixed")
    if params_optimization.get("method_GD", "adam") != "adam":
        raise NotImplementedError("only method_GD == 'adam' is supported")

    r_train = args[0]
    ntraining = sum(np.shape(r)[0] for r in r_train)
    vg = _value_and_grad(f, df)
    init = np.asarray(init, dtype=np.float64).copy()
    free = np.ones(len(init), dtype=bool)
    if index_fixed:
        free[np.asarray(index_fixed, dtype=int)] = False

    def value_and_grad(theta):
        v, g = vg(theta, *args)
        return v / ntraining, np.where(free, g, 0.0) / ntraining

    loss, theta, norms = [], [init.copy()], []
    loss_before = f(init, *args) / ntraining
    print(f"loss before optimize: {loss_before}")
    loss.append(loss_before)
    if np.isnan(loss_before):
        raise Exception("loss is nan at the initial hyper-parameters")

    if maxiter_scipy and maxiter_scipy[0]:
        from scipy.optimize import minimize

        x = theta[-1]
        for method, maxiter in zip(method_scipy, maxiter_scipy):
            res = minimize(lambda p: value_and_grad(p)[0], x, jac=(lambda p: value_and_grad(p)[1]) if method != "Nelder-Mead" else None,
                           method=method, options={"maxiter": maxiter})
            x = np.where(free, res.x, init)
            theta.append(x.copy())
            loss.append(float(res.fun))

    m, v = np.zeros_like(init), np.zeros_like(init)
    b1, b2, adam_eps = 0.9, 0.999, 1e-8
    converged_once = False
    for t in range(maxiter_GD):
        cur = theta[-1]
        value, grads = value_and_grad(cur)
        if np.any(np.isnan(grads)):
            raise Exception("gradient of loss became nan")
        m = b1 * m + (1.0 - b1) * grads
        v = b2 * v + (1.0 - b2) * grads * grads
        mhat, vhat = m / (1.0 - b1 ** (t + 1)), v / (1.0 - b2 ** (t + 1))
        new = cur - lr * mhat / (np.sqrt(vhat) + adam_eps)
        norms.append(float(np.linalg.norm(grads)))
        theta.append(new)
        loss.append(value)
        if print_process:
            with open("run.out", "a") as fh:
                fh.write(f"step{t:4} loss: {value:.4f} max_grad: {np.max(np.abs(grads)):.5f}, arg={np.argmax(np.abs(grads))}\n")
                fh.write(f"norm_of_grads: {norms[-1]:.5f}\n")
        if t < 1:
            continue
        if abs(loss[-1] - loss[-2]) < eps:
            if converged_once:
                print("converged")
                break
            converged_once = True
        elif np.any(np.isnan(theta[-1])):
            print("diverged")
            raise Exception("theta became nan")
        else:
            converged_once = False
    return theta[-1], loss, theta, norms
