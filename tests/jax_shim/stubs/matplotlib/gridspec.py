from _stub import install as _install
_install(globals(), "matplotlib.gridspec")
