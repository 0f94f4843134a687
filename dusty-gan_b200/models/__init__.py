"""Point-drop heads (reference models/__init__.py exposes define_G; only the heads are rebuilt here,
the DCGAN-eqlr backbone stays the reference's PyTorch module and is passed in as ``backbone``)."""
from . import dusty  # noqa: F401
from .dusty import DUSty1, DUSty2, GumbelSigmoid  # noqa: F401
