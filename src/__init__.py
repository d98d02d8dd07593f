"""Drop-in ``src`` package: the reference's transform entry points
(src/cwt.py, src/wct.py, src/xwt.py, src/dwt.py, src/modwt.py,
src/utils/wavelet_helpers.py) re-exported from the B200 engine.  Copy this
directory over the reference's files of the same name and
its app, helpers and tests keep importing ``from src import cwt`` unchanged."""
