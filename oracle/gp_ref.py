"""CPU restatement of the reference's GPmodel (assembly, NLL, gradient, posterior).

Follows /root/reference/GP/gp.py line by line, with numpy / LAPACK standing in
for jax.numpy:
  _add_jiggle :23-42, _add_jiggle_noise :44-70, logpGP :72-89, postGP :91-120,
  calculate_K_training :122-154, calculate_K_test :156-189,
  calculate_K_asymmetric :191-211, trainingFunction_all :213-224,
  predictingFunction_all :226-256, calc_sec :258-261, set_constants :263-285,
  calc_K_given_theta_i :364-372, d_trainingFunction_all :412-488,
  d_logposterior :491-493;  sub_modules/loss_modules.py:5-13 (logposterior).
The dense linear algebra keeps the reference's op sequence (Cholesky, then
*general* ``solve`` calls on the triangular factor, explicit inverse through
two solves against I, one dense matmul per hyper-parameter) because that
sequence is what the CPU baseline times.

Two assembly back ends:
  backend="autodiff": oracle.autodiff_ops (nested torch.func grad/hessian,
      the literal restatement; dK/dtheta by jacfwd like gp.py:459) -- small N.
  backend="closed":   oracle.closed_form (numpy closed forms) -- any N.
Test infrastructure only (see oracle/__init__.py).
"""
import numpy as np

from . import blocks_ref, closed_form


class GPRef:
    def __init__(self, table, kernel_form="product", dim=None, lbox=None, index_optimize_noise=None,
                 backend="closed", dtype=np.float64, kernel_type="se"):
        self.table = blocks_ref.TABLES[table]
        self.dim = self.table["dim"] if self.table["dim"] is not None else dim
        self.form = kernel_form if self.dim == 2 else "product"  # kernels.py:419-426: 3-D ignores kernel_form
        self.lbox = None if lbox is None else np.asarray(lbox, dtype=np.float64)
        self.index_optimize_noise = index_optimize_noise if index_optimize_noise else False
        self.backend = backend
        self.kernel_type = kernel_type  # "se" or "mt52" / "mt72" / "mt92" (closed backend only; oracle/matern_ref.py)
        self.dtype = dtype  # np.longdouble: closed-form blocks in extended precision (oracle/extended.py); closed backend only
        self._ad = None

    # ------------------------------------------------------------------ operators
    def _autodiff(self):
        if self._ad is None:
            import torch
            from . import autodiff_ops, kernels_ref

            kern = kernels_ref.define_kernel_ref(
                {"kernel_type": "se", "kernel_form": self.form, "input_dim": self.dim, "distance_func": False})
            self._ad = (torch, autodiff_ops.AutodiffOps(kern, self.dim))
        return self._ad

    def _op_eval(self, op, r, rp, theta_g):
        if self.backend == "closed":
            if self.kernel_type != "se":
                from . import matern_ref
                return matern_ref.eval_operator(self.kernel_type, op, r, rp, theta_g, self.form, self.dim)
            return closed_form.eval_operator(op, r, rp, theta_g, self.form, self.dim, dtype=self.dtype)
        torch, ad = self._autodiff()
        t = lambda x: x if isinstance(x, torch.Tensor) else torch.as_tensor(np.asarray(x), dtype=torch.float64)
        return ad.block(op)(t(r), t(rp), t(theta_g))

    def _zeros(self, n, m):
        if self.backend == "closed":
            return np.zeros((n, m), dtype=self.dtype)
        return self._autodiff()[0].zeros((n, m), dtype=self._autodiff()[0].float64)

    def block(self, name, r, rp, theta):
        lbox = self.lbox
        if self.backend != "closed" and lbox is not None:
            lbox = self._autodiff()[0].as_tensor(lbox)
        return blocks_ref.eval_block(name, self.table, self._op_eval, r, rp, theta, lbox, self._zeros)

    # ------------------------------------------------------------------ gp.py:258-285
    @staticmethod
    def calc_sec(pts):
        return np.concatenate([np.zeros(1, dtype=int), np.cumsum([len(x) for x in pts])])

    def set_constants(self, *args, only_training=False):
        if only_training:
            r_train, _, _ = args
        else:
            r_test, _, r_train, _, _ = args
            self.num_te, self.sec_te = len(r_test), self.calc_sec(r_test)
        self.num_tr, self.sec_tr = len(r_train), self.calc_sec(r_train)

    def split_hyp_and_noise(self, theta):
        if self.index_optimize_noise:
            return theta[:-1], theta[-1]
        return theta, None

    # ------------------------------------------------------------------ gp.py:23-70
    def add_eps_to_sigma(self, S, eps, noise_parameter=None):
        xp = np if isinstance(S, np.ndarray) else self._autodiff()[0]
        n = S.shape[0]
        if not self.index_optimize_noise:
            return S + xp.diag(xp.ones(n, dtype=S.dtype) * eps)
        lo = int(self.sec_tr[self.index_optimize_noise[0]])
        hi = int(self.sec_tr[self.index_optimize_noise[-1] + 1])
        scale = xp.ones(n, dtype=S.dtype)
        pieces = [scale[:lo], scale[lo:hi] * xp.exp(noise_parameter), scale[hi:] * eps]
        d = np.concatenate(pieces) if xp is np else xp.cat(pieces)
        return S + xp.diag(d)

    # ------------------------------------------------------------------ gp.py:122-211
    def _cat(self, grid):
        if self.backend == "closed":
            return np.block(grid)
        torch = self._autodiff()[0]
        return torch.cat([torch.cat(row, dim=1) for row in grid], dim=0)

    def _symmetric(self, pts, rows, theta):
        # gp.py:133-154: upper-triangular list of lists, lower blocks are the transposes
        nb = len(pts)
        grid = [[None] * nb for _ in range(nb)]
        for i in range(nb):
            for j in range(i, nb):
                B = self.block(rows[i][j - i], pts[i], pts[j], theta)
                grid[i][j] = B
                if j != i:
                    grid[j][i] = B.T
        return self._cat(grid)

    def trainingK_all(self, theta, train_pts):
        return self._symmetric(train_pts, self.table["training"], theta)

    def testK_all(self, theta, test_pts):
        return self._symmetric(test_pts, self.table["test"], theta)

    def mixedK_all(self, theta, test_pts, train_pts):
        # gp.py:204-211: Ks[i][j](test_i, train_j)
        rows = self.table["mixed"]
        return self._cat([[self.block(rows[i][j], test_pts[i], train_pts[j], theta)
                           for j in range(len(train_pts))] for i in range(len(test_pts))])

    def _np(self, x):
        return x if isinstance(x, np.ndarray) else x.detach().numpy()

    def _theta(self, theta):
        if self.backend == "closed":
            return np.asarray(theta, dtype=self.dtype)
        torch = self._autodiff()[0]
        return theta if isinstance(theta, torch.Tensor) else torch.as_tensor(np.asarray(theta), dtype=torch.float64)

    def _pts(self, pts):
        if self.backend == "closed":
            return [np.asarray(p, dtype=self.dtype) for p in pts]
        torch = self._autodiff()[0]
        return [torch.as_tensor(np.asarray(p), dtype=torch.float64) for p in pts]

    def training_sigma(self, theta, r, eps):
        """trainingK_all + add_eps_to_sigma (gp.py:221-223), returned as numpy."""
        theta = self._theta(theta)
        th, noise = self.split_hyp_and_noise(theta)
        S = self.trainingK_all(th, self._pts(r))
        return self._np(self.add_eps_to_sigma(S, eps, noise_parameter=noise))

    # ------------------------------------------------------------------ gp.py:72-89, 213-224
    @staticmethod
    def logpGP(dy, S):
        n = len(dy)
        L = np.linalg.cholesky(S)
        v = np.linalg.solve(L, dy)  # general solve on a triangular matrix, as the reference does
        return 0.5 * np.dot(v, v) + np.sum(np.log(np.diag(L))) + 0.5 * n * np.log(2.0 * np.pi)

    def trainingFunction_all(self, theta, r, delta_y, eps):
        return self.logpGP(np.asarray(delta_y, dtype=np.float64), self.training_sigma(theta, r, eps))

    def logposterior(self, params_optimization=None):
        """sub_modules/loss_modules.py:5-13."""
        po = params_optimization or {"loss_ridge_regression": False}
        if po.get("loss_ridge_regression"):
            return lambda th, *a: (self.trainingFunction_all(th, *a) + np.sum(np.asarray(th))
                                   + po["ridge_alpha"] * np.sum(np.square(np.exp(np.asarray(th)))))
        return lambda th, *a: self.trainingFunction_all(th, *a) + np.sum(np.asarray(th))

    # ------------------------------------------------------------------ gp.py:91-120, 226-256
    @staticmethod
    def postGP(dyb, Kaa, Kab, Kbb):
        L = np.linalg.cholesky(Kbb)
        alpha = np.linalg.solve(L.T, np.linalg.solve(L, dyb))
        mu = Kab @ alpha
        V = np.linalg.solve(L, Kab.T)
        return mu, Kaa - np.einsum("ji,jk->ik", V, V)

    def predictingFunction_all(self, theta, r_test, mu_test, r_train, delta_y, eps):
        theta_t = self._theta(theta)
        th, noise = self.split_hyp_and_noise(theta_t)
        Sbb = self.training_sigma(theta, r_train, eps)
        Sab = self._np(self.mixedK_all(th, self._pts(r_test), self._pts(r_train)))
        Saa = self._np(self.testK_all(th, self._pts(r_test)))
        mus, covs = self.postGP(np.asarray(delta_y, dtype=np.float64), Saa, Sab, Sbb)
        mu_out, cov_out, lo = [], [], 0
        for i in range(len(r_test)):
            hi = lo + len(r_test[i])
            mu_out.append(mus[lo:hi] + np.asarray(mu_test[i]))
            cov_out.append(covs[lo:hi, lo:hi])
            lo = hi
        return mu_out, cov_out

    # ------------------------------------------------------------------ gp.py:364-372, 412-493
    def dK_dtheta(self, theta, r, eps):
        """List over theta_p of dSigma/dtheta_p (N,N); autodiff: jacfwd (gp.py:459), closed: formulas."""
        P = len(theta)
        if self.backend != "closed":
            torch, _ = self._autodiff()
            from torch.func import jacfwd

            pts = self._pts(r)

            def sigma(th_full):
                th, noise = self.split_hyp_and_noise(th_full)
                return self.add_eps_to_sigma(self.trainingK_all(th, pts), eps, noise_parameter=noise)

            J = jacfwd(sigma)(self._theta(theta))  # (N, N, P)
            return [J[:, :, p].numpy() for p in range(P)]
        return self._dK_closed(np.asarray(theta, dtype=self.dtype), r)

    def _dK_closed(self, theta, r):
        pts = self._pts(r)
        sec = self.calc_sec(pts)
        N = int(sec[-1])
        th, noise = self.split_hyp_and_noise(theta)
        out = [np.zeros((N, N), dtype=self.dtype) for _ in range(len(theta))]
        rows, blocks, groups = self.table["training"], self.table["blocks"], self.table["groups"]
        for i in range(len(pts)):
            for j in range(i, len(pts)):
                terms, shift = blocks_ref.parse_spec(blocks[rows[i][j - i]], blocks)
                for sign, op, group in terms:
                    sl = groups[group]
                    idx = list(range(len(th)))[sl]

                    def base(a, b):
                        if self.kernel_type != "se":
                            from . import matern_ref
                            return matern_ref.eval_operator(self.kernel_type, op, a, b, th[sl], self.form, self.dim, with_grad=True)[1]
                        return closed_form.eval_operator(op, a, b, th[sl], self.form, self.dim, with_grad=True, dtype=self.dtype)[1]

                    a, b, l = pts[i], pts[j], self.lbox
                    if shift is None:
                        g = base(a, b)
                    elif shift == "Xp":
                        g = base(a, b + l) - base(a, b)
                    elif shift == "X":
                        g = base(a + l, b) - base(a, b)
                    else:
                        g = base(a + l, b + l) - base(a + l, b) - base(a, b + l) + base(a, b)
                    for q, p in enumerate(idx):
                        out[p][sec[i]:sec[i + 1], sec[j]:sec[j + 1]] += sign * g[q]
                        if j != i:
                            out[p][sec[j]:sec[j + 1], sec[i]:sec[i + 1]] += sign * g[q].T
        if self.index_optimize_noise:
            lo = int(sec[self.index_optimize_noise[0]])
            hi = int(sec[self.index_optimize_noise[-1] + 1])
            d = np.zeros(N, dtype=self.dtype)
            d[lo:hi] = np.exp(noise)
            out[-1] = np.diag(d)
        return out

    def d_trainingFunction_all(self, theta, r, delta_y, eps):
        dy = np.asarray(delta_y, dtype=np.float64)
        S = self.training_sigma(theta, r, eps)
        L = np.linalg.cholesky(S)
        I = np.eye(len(dy))
        S_inv = np.linalg.solve(L.T, np.linalg.solve(L, I))
        alpha = np.linalg.solve(L.T, np.linalg.solve(L, dy))
        del L, S, I
        dKs = self.dK_dtheta(theta, r, eps)
        dloss = np.zeros(len(dKs))
        for p, dK in enumerate(dKs):
            first = alpha @ (dK @ alpha)
            second = np.sum(np.diagonal(S_inv @ dK))
            dloss[p] = (-first + second) / 2
        return dloss

    def d_logposterior(self, theta, *args):
        return self.d_trainingFunction_all(theta, *args) + 1.0
