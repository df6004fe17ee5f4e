"""Block library and per-class block tables of the reference, as spec strings.

Each entry restates one method of the reference's block library:
  2-D: /root/reference/GP/gp_2D_stokes_independent.py:11-246
       (theta slices ind_uxux=0:3, ind_uyuy=3:6, ind_pp=6:9, :17-19)
  3-D: /root/reference/GP/gp_3D_stokes_independent.py:12-120, :238-239
       (theta slices 0:4, 4:8, 8:12, 12:16 with pp last, :18-21)
A spec is a sum of ``<sign><operator>:<theta group>`` terms (operators as in
oracle.autodiff_ops), ``"0"`` for the reference's Kzero blocks, or one of the
periodic-difference wrappers of /root/reference/GP/gp.py:
  Xp(B) = B(r, r'+l) - B(r, r')                       gp.py:374-383
  X(B)  = B(r+l, r') - B(r, r')                       gp.py:385-394
  XX(B) = B(r+l,r'+l) - B(r+l,r') - B(r,r'+l) + B(r,r')   gp.py:396-410

The class tables restate
  GPPoiseuilleIndependent          gp_poiseuille_independent.py:21-42
  GPSinusoidalWithoutPIndependent  gp_sinusoidal_independent.py:49-181
  GPStokes3D                       gp_stokes_3D.py:51-172
  GPmodelNaive                     gp_naive.py:31-45
  GPmodel1DLaplacian               gp_1D_laplacian.py:48-59
Quirks are kept by name (e.g. Kuydifux wraps Kuxuy, the (uy,ux) mixed slot
holds Kuxuy) rather than "fixed".
"""

BLOCKS_2D = {
    # plain (:22-38)
    "Kuxux": "+K:ux", "Kuxuy": "0", "Kuyuy": "+K:uy", "Kuxp": "0", "Kuyp": "0", "Kpp": "+K:pp",
    # governing x governing (:41-63)
    "Kfxfx": "+d0d0:pp +LL:ux", "Kfxfy": "+d0d1:pp", "Kfyfy": "+d1d1:pp +LL:uy",
    "Kfxdiv": "-Ld0:ux", "Kfydiv": "-Ld1:uy", "Kdivdiv": "+d0d0:ux +d1d1:uy",
    # variable x governing (:66-109)
    "Kuxfx": "-L1:ux", "Kuyfx": "0", "Kpfx": "+d10:pp", "Kuxfy": "0", "Kuyfy": "-L1:uy", "Kpfy": "+d11:pp",
    "Kuxdiv": "+d10:ux", "Kuydiv": "+d11:uy", "Kpdiv": "0",
    # governing x variable (:112-155)
    "Kfxp": "+d00:pp", "Kfyp": "+d01:pp", "Kdivp": "0", "Kuyux": "0", "Kpux": "0", "Kpuy": "0",
    "Kfxux": "-L0:ux", "Kfxuy": "0", "Kfyux": "0", "Kfyuy": "-L0:uy",
    "Kdivfx": "-d0L:ux", "Kdivfy": "-d1L:uy", "Kfyfx": "+d1d0:pp", "Kdivux": "+d00:ux", "Kdivuy": "+d01:uy",
    # periodic differences (:158-246)
    "Kuxdifux": "Xp(Kuxux)", "Kuxdifuy": "Xp(Kuxuy)", "Kuydifux": "Xp(Kuxuy)", "Kuydifuy": "Xp(Kuyuy)",
    "Kuxdifp": "Xp(Kuxp)", "Kuydifp": "Xp(Kuyp)", "Kpdifp": "Xp(Kpp)",
    "Kfxdifp": "Xp(Kfxp)", "Kfydifp": "Xp(Kfyp)", "Kdivdifp": "Xp(Kdivp)",
    "Kdifuxdifux": "XX(Kuxux)", "Kdifuxdifuy": "XX(Kuxuy)", "Kdifuxfx": "X(Kuxfx)", "Kdifuxfy": "X(Kuxfy)",
    "Kdifuxdiv": "X(Kuxdiv)", "Kdifuydifuy": "XX(Kuyuy)", "Kdifuyfx": "X(Kuyfx)", "Kdifuyfy": "X(Kuyfy)",
    "Kdifuydiv": "X(Kuydiv)", "Kdifuxp": "X(Kuxp)", "Kdifuxdifp": "XX(Kuxp)", "Kdifuyp": "X(Kuyp)",
    "Kdifuydifp": "XX(Kuyp)", "Kdifpdifp": "XX(Kpp)",
    "Kfxdifux": "Xp(Kfxux)", "Kfxdifuy": "Xp(Kfxuy)", "Kfydifux": "Xp(Kfyux)", "Kfydifuy": "Xp(Kfyuy)",
    "Kdivdifux": "Xp(Kdivux)", "Kdivdifuy": "Xp(Kdivuy)",
}
GROUPS_2D = {"ux": slice(0, 3), "uy": slice(3, 6), "pp": slice(6, 9)}

# 3-D: the class inherits the 2-D library (MRO) and adds / overrides these.
BLOCKS_3D = dict(BLOCKS_2D)
BLOCKS_3D.update({
    "Kuxuz": "0", "Kuyuz": "0", "Kuzuz": "+K:uz", "Kuzp": "0",                    # :25-35
    "Kfxfz": "+d0d2:pp", "Kfyfz": "+d1d2:pp", "Kfzfz": "+d2d2:pp +LL:uz",          # :39-48
    "Kfzdiv": "-Ld2:uz", "Kdivdiv": "+d0d0:ux +d1d1:uy +d2d2:uz",                  # :50-58
    "Kuxfz": "0", "Kuyfz": "0", "Kuzfx": "0", "Kuzfy": "0", "Kuzfz": "-L1:uz", "Kuzdiv": "+d12:uz",  # :61-77
    "Kpux": "0", "Kpuy": "0", "Kpuz": "0", "Kpfx": "+d10:pp", "Kpfy": "+d11:pp", "Kpfz": "+d12:pp",
    "Kpdiv": "0",                                                                   # :80-99
    "Kdifpux": "X(Kpux)", "Kdifpuy": "X(Kpuy)", "Kdifpuz": "X(Kpuz)", "Kdifpfx": "X(Kpfx)",
    "Kdifpfy": "X(Kpfy)", "Kdifpfz": "X(Kpfz)", "Kdifpdiv": "X(Kpdiv)",            # :101-120
    "Kdifpdifp": "XX(Kpp)",                                                         # :238-239
})
GROUPS_3D = {"ux": slice(0, 4), "uy": slice(4, 8), "uz": slice(8, 12), "pp": slice(12, 16)}

# single-block models: theta is used whole
BLOCKS_NAIVE = {"Kyy": "+K:all"}
BLOCKS_1D_LAPLACIAN = {"Kyy": "+K:all", "Kyly": "+L1K:all", "Klyly": "+LLK:all"}
GROUPS_ALL = {"all": slice(None)}


def _rows(text):
    return [row.split() for row in text.strip().splitlines()]


TABLES = {
    "poiseuille": {
        "dim": 2, "blocks": BLOCKS_2D, "groups": GROUPS_2D,
        "training": _rows("""
            Kuxux Kuxuy Kuxp Kuxfx Kuxfy Kuxdiv
            Kuyuy Kuyp Kuyfx Kuyfy Kuydiv
            Kpp Kpfx Kpfy Kpdiv
            Kfxfx Kfxfy Kfxdiv
            Kfyfy Kfydiv
            Kdivdiv"""),
        "mixed": _rows("""
            Kuxux Kuxuy Kuxp Kuxfx Kuxfy Kuxdiv
            Kuyux Kuyuy Kuyp Kuyfx Kuyfy Kuydiv
            Kpux Kpuy Kpp Kpfx Kpfy Kpdiv"""),
        "test": _rows("""
            Kuxux Kuxuy Kuxp
            Kuyuy Kuyp
            Kpp"""),
    },
    # use_difp=True, use_difu=True, infer_governing_eqs=False (test_0 / test_1 configuration)
    "sinusoidal": {
        "dim": 2, "blocks": BLOCKS_2D, "groups": GROUPS_2D,
        "training": _rows("""
            Kuxux Kuxuy Kuxdifux Kuxdifuy Kuxfx Kuxfy Kuxdiv Kuxdifp
            Kuyuy Kuydifux Kuydifuy Kuyfx Kuyfy Kuydiv Kuydifp
            Kdifuxdifux Kdifuxdifuy Kdifuxfx Kdifuxfy Kdifuxdiv Kdifuxdifp
            Kdifuydifuy Kdifuyfx Kdifuyfy Kdifuydiv Kdifuydifp
            Kfxfx Kfxfy Kfxdiv Kfxdifp
            Kfyfy Kfydiv Kfydifp
            Kdivdiv Kdivdifp
            Kdifpdifp"""),
        "mixed": _rows("""
            Kuxux Kuxuy Kuxdifux Kuxdifuy Kuxfx Kuxfy Kuxdiv Kuxdifp
            Kuxuy Kuyuy Kuydifux Kuydifuy Kuyfx Kuyfy Kuydiv Kuydifp"""),
        "test": _rows("""
            Kuxux Kuxuy
            Kuyuy"""),
    },
    # infer_governing_eqs=True variant (gp_sinusoidal_independent.py:92-124, :171-176)
    "sinusoidal_infer_gov": {
        "dim": 2, "blocks": BLOCKS_2D, "groups": GROUPS_2D,
        "training": None,  # same as "sinusoidal"
        "mixed": _rows("""
            Kfxux Kfxuy Kfxdifux Kfxdifuy Kfxfx Kfxfy Kfxdiv Kfxdifp
            Kfyux Kfyuy Kfydifux Kfydifuy Kfyfx Kfyfy Kfydiv Kfydifp
            Kdivux Kdivuy Kdivdifux Kdivdifuy Kdivfx Kdivfy Kdivdiv Kdivdifp"""),
        "test": _rows("""
            Kfxfx Kfxfy Kfxdiv
            Kfyfy Kfydiv
            Kdivdiv"""),
    },
    "stokes3d": {
        "dim": 3, "blocks": BLOCKS_3D, "groups": GROUPS_3D,
        "training": _rows("""
            Kuxux Kuxuy Kuxuz Kuxfx Kuxfy Kuxfz Kuxdiv
            Kuyuy Kuyuz Kuyfx Kuyfy Kuyfz Kuydiv
            Kuzuz Kuzfx Kuzfy Kuzfz Kuzdiv
            Kfxfx Kfxfy Kfxfz Kfxdiv
            Kfyfy Kfyfz Kfydiv
            Kfzfz Kfzdiv
            Kdivdiv"""),
        "mixed": _rows("""
            Kuxux Kuxuy Kuxuz Kuxfx Kuxfy Kuxfz Kuxdiv
            Kuxuy Kuyuy Kuyuz Kuyfx Kuyfy Kuyfz Kuydiv
            Kuxuz Kuyuz Kuzuz Kuzfx Kuzfy Kuzfz Kuzdiv"""),
        "test": _rows("""
            Kuxux Kuxuy Kuxuz
            Kuyuy Kuyuz
            Kuzuz"""),
    },
    "naive": {
        "dim": None, "blocks": BLOCKS_NAIVE, "groups": GROUPS_ALL,
        "training": [["Kyy"]], "mixed": [["Kyy"]], "test": [["Kyy"]],
    },
    "laplacian1d": {
        "dim": 1, "blocks": BLOCKS_1D_LAPLACIAN, "groups": GROUPS_ALL,
        "training": [["Kyy", "Kyly"], ["Klyly"]], "mixed": [["Kyy", "Kyly"]], "test": [["Kyy"]],
    },
}
TABLES["sinusoidal_infer_gov"]["training"] = TABLES["sinusoidal"]["training"]

# ---- the other live classes of the reference, restated name by name
# GP/gp_sinusoidal_infer_difp.py:9-41 adds these methods to the 2-D library
BLOCKS_2D_INFER_DIFP = dict(BLOCKS_2D)
BLOCKS_2D_INFER_DIFP.update({
    "Kpux": "0", "Kpuy": "0", "Kpfx": "+d10:pp", "Kpfy": "+d11:pp", "Kpdiv": "0",
    "Kdifpux": "X(Kpux)", "Kdifpuy": "X(Kpuy)", "Kdifpfx": "X(Kpfx)", "Kdifpfy": "X(Kpfy)", "Kdifpdiv": "X(Kpdiv)",
    "Kdifpdifux": "XX(Kpux)", "Kdifpdifuy": "XX(Kpuy)",
})
_SIN7 = _rows("""
    Kuxux Kuxuy Kuxdifux Kuxdifuy Kuxfx Kuxfy Kuxdiv
    Kuyuy Kuydifux Kuydifuy Kuyfx Kuyfy Kuydiv
    Kdifuxdifux Kdifuxdifuy Kdifuxfx Kdifuxfy Kdifuxdiv
    Kdifuydifuy Kdifuyfx Kdifuyfy Kdifuydiv
    Kfxfx Kfxfy Kfxdiv
    Kfyfy Kfydiv
    Kdivdiv""")  # gp_sinusoidal_infer_difp.py:44-59
TABLES.update({
    "sinusoidal_infer_difp": {  # GPSinusoidalInferDifP :60-68
        "dim": 2, "blocks": BLOCKS_2D_INFER_DIFP, "groups": GROUPS_2D, "training": _SIN7,
        "mixed": _rows("Kdifpux Kdifpuy Kdifpdifux Kdifpdifuy Kdifpfx Kdifpfy Kdifpdiv"),
        "test": _rows("Kdifpdifp"),
    },
    "sinusoidal_infer_u_without_difp": {  # GPSinusoidalInferUWithoutDifP :71-82
        "dim": 2, "blocks": BLOCKS_2D, "groups": GROUPS_2D, "training": _SIN7,
        "mixed": _rows("""
            Kuxux Kuxuy Kuxdifux Kuxdifuy Kuxfx Kuxfy Kuxdiv
            Kuyux Kuyuy Kuydifux Kuydifuy Kuyfx Kuyfy Kuydiv"""),
        "test": _rows("""
            Kuxux Kuxuy
            Kuyuy"""),
    },
    "sinusoidal_infer_gov_without_difp": {  # GPSinusoidalInferGovWithoutDifP :85-101 (Kfxuy in the (fx, fy) test slot: :97)
        "dim": 2, "blocks": BLOCKS_2D, "groups": GROUPS_2D, "training": _SIN7,
        "mixed": _rows("""
            Kfxux Kfxuy Kfxdifux Kfxdifuy Kfxfx Kfxfy Kfxdiv
            Kfyux Kfyuy Kfydifux Kfydifuy Kfyfx Kfyfy Kfydiv
            Kdivux Kdivuy Kdivdifux Kdivdifuy Kdivfx Kdivfy Kdivdiv"""),
        "test": _rows("""
            Kfxfx Kfxuy Kfxdiv
            Kfyfy Kfydiv
            Kdivdiv"""),
    },
    "stokes3d_infer_difp": {  # GPStokes3D(infer_difp=True): gp_stokes_3D.py:112-123, :163-166
        "dim": 3, "blocks": BLOCKS_3D, "groups": GROUPS_3D, "training": TABLES["stokes3d"]["training"],
        "mixed": _rows("Kdifpux Kdifpuy Kdifpuz Kdifpfx Kdifpfy Kdifpfz Kdifpdiv"),
        "test": _rows("Kdifpdifp"),
    },
    "stokes3d_naive": {  # GPStokes3DNaive: gp_stokes_3D_naive.py:49-128
        "dim": 3, "blocks": BLOCKS_3D, "groups": GROUPS_3D,
        "training": _rows("""
            Kuxux Kuxuy Kuxuz
            Kuyuy Kuyuz
            Kuzuz"""),
        "mixed": _rows("""
            Kuxux Kuxuy Kuxuz
            Kuxuy Kuyuy Kuyuz
            Kuxuz Kuyuz Kuzuz"""),
        "test": TABLES["stokes3d"]["test"],
    },
    "stokes2d2c": {  # GPStokes2D2C: gp_stokes_3D_2D2C.py:14-78
        "dim": 3, "blocks": BLOCKS_3D, "groups": GROUPS_3D,
        "training": _rows("""
            Kuxux Kuxuy Kuxfx Kuxfy Kuxfz Kuxdiv
            Kuyuy Kuyfx Kuyfy Kuyfz Kuydiv
            Kfxfx Kfxfy Kfxfz Kfxdiv
            Kfyfy Kfyfz Kfydiv
            Kfzfz Kfzdiv
            Kdivdiv"""),
        "mixed": _rows("""
            Kuxux Kuxuy Kuxfx Kuxfy Kuxfz Kuxdiv
            Kuxuy Kuyuy Kuyfx Kuyfy Kuyfz Kuydiv
            Kuxuz Kuyuz Kuzfx Kuzfy Kuzfz Kuzdiv"""),
        "test": TABLES["stokes3d"]["test"],
    },
    "stokes2d2c_surface": {  # GPStokes2D2CSurface: gp_stokes_3D_2D2C.py:86-188
        "dim": 3, "blocks": BLOCKS_3D, "groups": GROUPS_3D,
        "training": _rows("""
            Kuxux Kuxuy Kuxux Kuxuy Kuxuz Kuxfx Kuxfy Kuxfz Kuxdiv
            Kuyuy Kuyux Kuyuy Kuyuz Kuyfx Kuyfy Kuyfz Kuydiv
            Kuxux Kuxuy Kuxuz Kuxfx Kuxfy Kuxfz Kuxdiv
            Kuyuy Kuyuz Kuyfx Kuyfy Kuyfz Kuydiv
            Kuzuz Kuzfx Kuzfy Kuzfz Kuzdiv
            Kfxfx Kfxfy Kfxfz Kfxdiv
            Kfyfy Kfyfz Kfydiv
            Kfzfz Kfzdiv
            Kdivdiv"""),
        "mixed": _rows("""
            Kuxux Kuxuy Kuxux Kuxuy Kuxuz Kuxfx Kuxfy Kuxfz Kuxdiv
            Kuyux Kuyuy Kuyux Kuyuy Kuyuz Kuyfx Kuyfy Kuyfz Kuydiv
            Kuxuz Kuyuz Kuxuz Kuyuz Kuzuz Kuzfx Kuzfy Kuzfz Kuzdiv"""),
        "test": TABLES["stokes3d"]["test"],
    },
})


def parse_spec(spec, blocks):
    """-> (terms, shift) with terms = [(sign, op, group)], shift in {None,'Xp','X','XX'}."""
    spec = spec.strip()
    for wrap in ("XX", "Xp", "X"):
        if spec.startswith(wrap + "("):
            inner, shift = parse_spec(blocks[spec[len(wrap) + 1:-1]], blocks)
            assert shift is None
            return inner, wrap
    if spec == "0":
        return [], None
    terms = []
    for tok in spec.split():
        op, group = tok[1:].split(":")
        terms.append((1.0 if tok[0] == "+" else -1.0, op, group))
    return terms, None


def eval_block(name, table, op_eval, r, rp, theta, lbox=None, zeros=None):
    """Evaluate one named block.

    op_eval(op_name, r, rp, theta_group) -> dense (n, m) block; ``zeros(n, m)`` makes the Kzero block.
    """
    terms, shift = parse_spec(table["blocks"][name], table["blocks"])

    def base(a, b):
        acc = zeros(len(a), len(b))
        for sign, op, group in terms:
            acc = acc + sign * op_eval(op, a, b, theta[table["groups"][group]])
        return acc

    if shift is None:
        return base(r, rp)
    if shift == "Xp":
        return base(r, rp + lbox) - base(r, rp)
    if shift == "X":
        return base(r + lbox, rp) - base(r, rp)
    return base(r + lbox, rp + lbox) - base(r + lbox, rp) - base(r, rp + lbox) + base(r, rp)
