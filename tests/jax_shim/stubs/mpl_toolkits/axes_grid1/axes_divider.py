from _stub import install as _install
_install(globals(), "mpl_toolkits.axes_grid1.axes_divider")
