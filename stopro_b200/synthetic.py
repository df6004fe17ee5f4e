"""Synthetic inputs with the shapes of the reference's configurations (pure numpy; no file I/O).

These restate the *mesh rules* of the reference's data generators so that the benchmark and the parity tests see
block structures of the right shape; they are inputs, not part of the hot path.
  C1 sin_1D naive      data_generator/sin_1D_naive.py:18-22 (x = linspace(0, 2pi, n), y = sin x + noise)
  C2 Poiseuille        data_generator/poiseuille.py:39-43, 62-142 and stokes_2D_generator.py:43-48 (x fastest mesh)
  C3 sinusoidal wall   data_generator/sinusoidal.py:989-1111 (walls y = +-(0.2 sin(2 pi x / 2.5) + 0.5), inlet
                       differences over lbox = [2.5, 0], constant body force 30/2.5)
  C4 drag3D            data_generator/drag3D.py:37-59, 82-95 (cube mesh minus a ball, Stokes-sphere velocity)
  C5 scaling sweep     SURVEY.md section 8(d): Poiseuille block structure, i.i.d. uniform points
Each generator returns a dict with r_train / f_train (lists per block), r_test / f_test, theta0, eps and the model
arguments.  delta_y = concatenate(f_train) because mu_train = 0 (sub_modules/load_modules.py:18).
"""
import numpy as np


def mesh2(x, y):
    xx, yy = np.meshgrid(x, y)  # x varies fastest, like StokesDataGenerator.make_r_mesh
    return np.stack([xx.reshape(-1), yy.reshape(-1)], axis=1)


def _pack(name, model, model_kwargs, kernel, r_train, f_train, r_test, f_test, theta0, eps=1e-6):
    return dict(name=name, model=model, model_kwargs=model_kwargs, kernel=kernel, r_train=r_train, f_train=f_train,
                r_test=r_test, f_test=f_test, mu_test=[np.zeros(len(r)) for r in r_test],
                delta_y=np.concatenate(f_train), theta0=np.asarray(theta0, dtype=np.float64), eps=eps)


def sin_1d_naive(n=32, n_test=300, seed=0):
    rng = np.random.RandomState(seed)
    x = np.linspace(0.0, 2.0 * np.pi, n)
    y = np.sin(x) + rng.normal(0.0, 0.1, n)
    xt = np.linspace(-0.5 * np.pi, 2.5 * np.pi, n_test)
    kernel = dict(kernel_type="se", kernel_form="product", input_dim=1, distance_func=False)
    return _pack("sin_1D_naive", "naive", dict(index_optimize_noise=[0]), kernel, [x], [y], [xt], [np.sin(xt)],
                 [0.0, 0.0, np.log(4e-4)])


def sin_1d_laplacian(ly_num=10, n_test=300):
    ry = np.array([0.0, 2.0 * np.pi])
    rl = np.linspace(0.0, 2.0 * np.pi, ly_num + 1)[:-1]
    xt = np.linspace(-np.pi, 3.0 * np.pi, n_test)
    kernel = dict(kernel_type="se", kernel_form="product", input_dim=1, distance_func=False)
    return _pack("sin_1D_laplacian", "laplacian1d", {}, kernel, [ry, rl], [np.sin(ry), -np.sin(rl)], [xt], [np.sin(xt)],
                 [0.0, 0.0])


def poiseuille(u_num=11, p_num=11, f_num=12, div_num=None, pad=0.03, n_test=33, kernel_form="additive"):
    div_num = f_num if div_num is None else div_num
    r_u = mesh2(np.linspace(0.0, 1.0, u_num), np.array([0.0, 1.0]))       # walls
    r_p = mesh2(np.array([0.0, 1.0]), np.linspace(0.0, 1.0, p_num))       # inlet / outlet
    r_f = mesh2(np.linspace(pad, 1.0 - pad, f_num), np.linspace(pad, 1.0 - pad, f_num))
    r_d = mesh2(np.linspace(pad, 1.0 - pad, div_num), np.linspace(pad, 1.0 - pad, div_num))
    ux = lambda r: 0.5 * r[:, 1] * (1.0 - r[:, 1])
    p = lambda r: 1.0 - r[:, 0]
    r_t = mesh2(np.linspace(0.0, 1.0, n_test), np.linspace(0.0, 1.0, n_test))
    kernel = dict(kernel_type="se", kernel_form=kernel_form, input_dim=2, distance_func=False)
    return _pack("poiseuille", "poiseuille", {}, kernel,
                 [r_u, r_u.copy(), r_p, r_f, r_f.copy(), r_d],
                 [ux(r_u), np.zeros(len(r_u)), p(r_p), np.zeros(len(r_f)), np.zeros(len(r_f)), np.zeros(len(r_d))],
                 [r_t, r_t.copy(), r_t.copy()], [ux(r_t), np.zeros(len(r_t)), p(r_t)], np.zeros(9))


def _wall(x, amp=0.2, period=2.5, half=0.5):
    return amp * np.sin(2.0 * np.pi * x / period) + half


def sinusoidal(u_num=31, f_nx=26, f_ny=13, dif_num=15, n_test=24, period=2.5):
    xs = np.linspace(0.0, period, u_num)
    r_wall = np.concatenate([np.stack([xs, _wall(xs)], 1), np.stack([xs, -_wall(xs)], 1)])  # 2 * u_num wall points
    # interior grid stretched between the walls (never on them)
    gx = np.linspace(0.0, period, f_nx + 1)[:-1] + 0.5 * period / f_nx
    gy = np.linspace(-1.0, 1.0, f_ny + 2)[1:-1]
    xx, yy = np.meshgrid(gx, gy)
    r_in = np.stack([xx.reshape(-1), (yy * _wall(xx)).reshape(-1)], 1)
    y_in = np.linspace(-1.0, 1.0, dif_num + 2)[1:-1] * _wall(0.0)
    r_dif = np.stack([np.zeros(dif_num), y_in], 1)  # inlet column; partner points are r + lbox
    tx = np.linspace(0.0, period, n_test + 1)[:-1] + 0.5 * period / n_test
    ty = np.linspace(-1.0, 1.0, n_test + 2)[1:-1]
    txx, tyy = np.meshgrid(tx, ty)
    r_t = np.stack([txx.reshape(-1), (tyy * _wall(txx)).reshape(-1)], 1)
    n_in = len(r_in)
    f_train = [np.zeros(len(r_wall)), np.zeros(len(r_wall)), np.zeros(dif_num), np.zeros(dif_num),
               np.full(n_in, 30.0 / period), np.zeros(n_in), np.zeros(n_in), np.zeros(dif_num)]
    r_train = [r_wall, r_wall.copy(), r_dif, r_dif.copy(), r_in, r_in.copy(), r_in.copy(), r_dif.copy()]
    kernel = dict(kernel_type="se", kernel_form="product", input_dim=2, distance_func=False)
    return _pack("sinusoidal", "sinusoidal", dict(lbox=np.array([period, 0.0]), use_difp=True, use_difu=True), kernel,
                 r_train, f_train, [r_t, r_t.copy()], [np.zeros(len(r_t)), np.zeros(len(r_t))],
                 np.tile([0.0, -1.0, -1.0], 3))


def _stokes_sphere(r, a=0.4, U0=1.0):
    """Velocity of uniform flow U0 e_x past a fixed sphere of radius a (Stokes solution)."""
    rr = np.linalg.norm(r, axis=1)
    x = r[:, 0]
    c1 = 1.0 - 0.75 * a / rr - 0.25 * a ** 3 / rr ** 3
    c2 = -0.75 * a / rr ** 3 + 0.75 * a ** 3 / rr ** 5
    u = U0 * (c1[:, None] * np.array([1.0, 0.0, 0.0])[None, :] + (c2 * x)[:, None] * r)
    return u[:, 0], u[:, 1], u[:, 2]


def drag3d(n_u=6, n_f=8, n_test=40, radius_cut=0.43):
    def cube_minus_ball(n):
        g = np.linspace(-0.97, 0.97, n)
        xx, yy, zz = np.meshgrid(g, g, g)  # default 'xy' indexing, like stokes_3D_generator.py:15-23
        r = np.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], 1)
        return r[np.linalg.norm(r, axis=1) > radius_cut]

    r_u, r_f = cube_minus_ball(n_u), cube_minus_ball(n_f)
    ux, uy, uz = _stokes_sphere(r_u)
    g = np.linspace(-0.97, 0.97, n_test)
    xx, yy = np.meshgrid(g, g)
    r_t = np.stack([xx.reshape(-1), yy.reshape(-1), np.zeros(n_test * n_test)], 1)
    r_t = r_t[np.linalg.norm(r_t, axis=1) > radius_cut]
    tx, ty, tz = _stokes_sphere(r_t)
    zf = np.zeros(len(r_f))
    kernel = dict(kernel_type="se", kernel_form="product", input_dim=3, distance_func=False)
    return _pack("drag3D", "stokes3d", {}, kernel, [r_u, r_u.copy(), r_u.copy(), r_f, r_f.copy(), r_f.copy(), r_f.copy()],
                 [ux, uy, uz, zf, zf.copy(), zf.copy(), zf.copy()], [r_t, r_t.copy(), r_t.copy()], [tx, ty, tz],
                 np.tile([0.0, -1.0, -1.0, -1.0], 4))


def sinusoidal_without_difp(variant="infer_difp", **kw):
    """The sinusoidal channel trained without the pressure-difference block (GP/gp_sinusoidal_infer_difp.py):
    variant in {"infer_difp", "infer_u_without_difp", "infer_gov_without_difp"}."""
    base = sinusoidal(**kw)
    r_train, f_train = base["r_train"][:7], base["f_train"][:7]
    r_t = base["r_test"][0]
    if variant == "infer_difp":
        r_test = [base["r_train"][7]]          # inlet column: p(r + lbox) - p(r)
        f_test = [np.full(len(r_test[0]), -30.0)]
    elif variant == "infer_u_without_difp":
        r_test, f_test = [r_t, r_t.copy()], [np.zeros(len(r_t)), np.zeros(len(r_t))]
    else:
        r_test = [r_t, r_t.copy(), r_t.copy()]
        f_test = [np.full(len(r_t), 12.0), np.zeros(len(r_t)), np.zeros(len(r_t))]
    return _pack("sinusoidal_" + variant, "sinusoidal_" + variant, dict(base["model_kwargs"], use_difp=False), base["kernel"],
                 r_train, f_train, r_test, f_test, base["theta0"])


def drag3d_variant(variant="stokes3d_naive", n_u=4, n_f=5, n_test=10):
    """The 3-D sphere-drag data arranged for the other live 3-D classes: "stokes3d_naive" (velocity only,
    GP/gp_stokes_3D_naive.py), "stokes2d2c" / "stokes2d2c_surface" (GP/gp_stokes_3D_2D2C.py) and
    "stokes3d_infer_difp" (GPStokes3D(infer_difp=True), GP/gp_stokes_3D.py:112-123)."""
    base = drag3d(n_u=n_u, n_f=n_f, n_test=n_test)
    r_u, r_f = base["r_train"][0], base["r_train"][3]
    ux, uy, uz = base["f_train"][:3]
    zf = np.zeros(len(r_f))
    kw = {}
    r_test, f_test = base["r_test"], base["f_test"]
    if variant == "stokes3d_naive":
        r_train, f_train = [r_u, r_u.copy(), r_u.copy()], [ux, uy, uz]
    elif variant == "stokes2d2c":
        r_train = [r_u, r_u.copy(), r_f, r_f.copy(), r_f.copy(), r_f.copy()]
        f_train = [ux, uy, zf, zf.copy(), zf.copy(), zf.copy()]
    elif variant == "stokes2d2c_surface":
        r_s = r_u[np.abs(r_u[:, 2]) > 0.9]      # "surface" points carrying all three components
        sx, sy, sz = _stokes_sphere(r_s)
        r_train = [r_u, r_u.copy(), r_s, r_s.copy(), r_s.copy(), r_f, r_f.copy(), r_f.copy(), r_f.copy()]
        f_train = [ux, uy, sx, sy, sz, zf, zf.copy(), zf.copy(), zf.copy()]
    elif variant == "stokes3d_infer_difp":
        r_train, f_train = base["r_train"], base["f_train"]
        kw = dict(lbox=np.array([1.94, 0.0, 0.0]), infer_difp=True)
        g = np.linspace(-0.9, 0.9, 5)
        yy, zz = np.meshgrid(g, g)
        r_in = np.stack([np.full(25, -0.97), yy.reshape(-1), zz.reshape(-1)], 1)
        r_test, f_test = [r_in], [np.zeros(25)]
    else:
        raise ValueError(variant)
    return _pack(variant, variant, kw, base["kernel"], r_train, f_train, r_test, f_test, base["theta0"])


def stokes2d_scaling(n_total=20000, seed=0, well_conditioned=True, n_test=1024):
    """C5: Poiseuille block structure [ux, uy, p, fx, fy, div] with fractions [.05, .05, .05, .2833, .2833, .2834],
    i.i.d. U[0,1]^2 points, product SE.  well_conditioned ties every length scale to the point spacing
    (logl = log(4 / sqrt(N))); otherwise the schema-faithful theta = [0, -1, -1] x 3."""
    rng = np.random.default_rng(seed)
    frac = np.array([0.05, 0.05, 0.05, 0.2833, 0.2833, 0.2834])
    counts = np.floor(frac * n_total).astype(int)
    counts[-1] += n_total - counts.sum()
    r_train = [rng.random((c, 2)) for c in counts]
    ux = lambda r: 0.5 * r[:, 1] * (1.0 - r[:, 1])
    p = lambda r: 1.0 - r[:, 0]
    f_train = [ux(r_train[0]), np.zeros(counts[1]), p(r_train[2]), np.zeros(counts[3]), np.zeros(counts[4]), np.zeros(counts[5])]
    r_t = rng.random((n_test, 2))
    logl = np.log(4.0 / np.sqrt(n_total)) if well_conditioned else -1.0
    kernel = dict(kernel_type="se", kernel_form="product", input_dim=2, distance_func=False)
    return _pack(f"stokes2d_N{n_total}", "poiseuille", {}, kernel, r_train, f_train, [r_t, r_t.copy(), r_t.copy()],
                 [ux(r_t), np.zeros(n_test), p(r_t)], np.tile([0.0, logl, logl], 3))


def make_model(cfg):
    """Instantiate the stopro_b200 model class for a generated configuration."""
    from .GP.kernels import define_kernel
    from .GP.gp_naive import GPmodelNaive
    from .GP.gp_1D_laplacian import GPmodel1DLaplacian
    from .GP.gp_poiseuille_independent import GPPoiseuilleIndependent
    from .GP.gp_sinusoidal_independent import GPSinusoidalWithoutPIndependent
    from .GP.gp_stokes_3D import GPStokes3D

    from .GP.gp_sinusoidal_infer_difp import (GPSinusoidalInferDifP, GPSinusoidalInferGovWithoutDifP,
                                              GPSinusoidalInferUWithoutDifP)
    from .GP.gp_stokes_3D_2D2C import GPStokes2D2C, GPStokes2D2CSurface
    from .GP.gp_stokes_3D_naive import GPStokes3DNaive

    cls = dict(naive=GPmodelNaive, laplacian1d=GPmodel1DLaplacian, poiseuille=GPPoiseuilleIndependent,
               sinusoidal=GPSinusoidalWithoutPIndependent, stokes3d=GPStokes3D,
               sinusoidal_infer_gov=GPSinusoidalWithoutPIndependent,
               sinusoidal_infer_difp=GPSinusoidalInferDifP,
               sinusoidal_infer_u_without_difp=GPSinusoidalInferUWithoutDifP,
               sinusoidal_infer_gov_without_difp=GPSinusoidalInferGovWithoutDifP,
               stokes3d_infer_difp=GPStokes3D, stokes3d_naive=GPStokes3DNaive, stokes2d2c=GPStokes2D2C,
               stokes2d2c_surface=GPStokes2D2CSurface)[cfg["model"]]
    return cls(Kernel=define_kernel(cfg["kernel"]), **cfg["model_kwargs"])


def from_golden(path):
    """A configuration whose inputs were produced by the REFERENCE's own data generator and stored in a golden fixture
    (tests/golden/ref_c*.npz, written by tests/golden/make_golden_ref.py): the BASELINE configurations C1-C4 at their
    true sizes (N = 32 / 498 / 1180 / 2640), including the 576 FEM test points of the sinusoidal channel."""
    import json

    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    nb_tr, nb_te = len(g["sec_tr"]) - 1, len(g["sec_te"]) - 1
    r_train = [g[f"r_train_{i}"] for i in range(nb_tr)]
    f_train = [g[f"f_train_{i}"] for i in range(nb_tr)]
    r_test = [g[f"r_test_{i}"] for i in range(nb_te)]
    f_test = [g[f"f_test_{i}"] for i in range(nb_te)]
    kw = {k: (np.asarray(v, dtype=np.float64) if k == "lbox" and v is not None else v)
          for k, v in meta["model_kwargs"].items()}
    return _pack(meta["model"], meta["model"], kw, meta["kernel"], r_train, f_train, r_test, f_test, g["theta0"],
                 eps=float(g["eps"]))
