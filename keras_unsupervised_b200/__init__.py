"""keras_unsupervised_b200 - the RBM/DBN contrastive-divergence path of tonandr/keras_unsupervised
(ku.ebm) on B200: hand-written sm_100a kernels behind a C ABI (include/kucd.h, libkucd.so), driven
from Python with the reference's class surface.

    from keras_unsupervised_b200.ebm import RBM, DBN      # instead of: from ku.ebm import RBM, DBN

There is no CPU implementation in this package.  Importing it is cheap; the first object that needs
the engine loads libkucd.so and raises if it is absent or if no sm_100 GPU is visible.
"""
__version__ = "0.1.0"

from . import ebm  # noqa: F401
