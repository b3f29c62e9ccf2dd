"""Same exports as /root/reference/ku/ebm/__init__.py:1-2."""
from .dbn import DBN  # noqa: F401
from .rbm import RBM, MODE_VISIBLE_BERNOULLI, MODE_VISIBLE_GAUSSIAN, MODE_COMPLEX  # noqa: F401
