"""keras_unsupervised_b200 - the RBM/DBN contrastive-divergence path of tonandr/keras_unsupervised
(ku.ebm) on B200: hand-written sm_100a kernels behind a C ABI (include/kucd.h, libkucd.so), driven
from Python with the reference's class surface.

    from keras_unsupervised_b200.ebm import RBM, DBN      # instead of: from ku.ebm import RBM, DBN

There is no CPU implementation in this package.  Importing it is cheap; the first object that needs
the engine loads libkucd.so and raises if it is absent or if no sm_100 GPU is visible.
"""
__version__ = "0.1.0"

from . import ebm  # noqa: F401


def install_as_ku(force: bool = False) -> None:
    """Make the reference's own import lines resolve to this engine:

        import keras_unsupervised_b200 as kucd; kucd.install_as_ku()
        from ku.ebm.rbm import RBM          # examples/rbm/rbm_softmax_mnist.py:24, unchanged
        from ku.ebm import RBM, DBN         # ku/ebm/__init__.py:1-2

    Registers `ku`, `ku.ebm`, `ku.ebm.rbm` and `ku.ebm.dbn` in sys.modules as aliases of the modules of this
    package.  Only the RBM/DBN path exists here: any other `ku.*` import fails as it would without the package.
    Refuses to shadow a real `ku` that is already imported unless force=True."""
    import sys
    import types

    from . import ebm as _ebm
    from .ebm import dbn as _dbn, rbm as _rbm

    present = sys.modules.get("ku")
    if present is not None and not getattr(present, "_kucd_alias", False) and not force:
        raise RuntimeError("a real `ku` package is already imported; pass force=True to replace its ku.ebm")
    ku = types.ModuleType("ku")
    ku.__doc__ = "alias of keras_unsupervised_b200 (RBM/DBN path only), installed by install_as_ku()"
    ku.__path__ = []  # a package with nothing else inside
    ku._kucd_alias = True
    ku.ebm = _ebm
    sys.modules.update({"ku": ku, "ku.ebm": _ebm, "ku.ebm.rbm": _rbm, "ku.ebm.dbn": _dbn})
