"""Import alias for the package directory ``dusty-gan_b200/``.

A hyphen cannot appear in a Python identifier, so ``import dusty_gan_b200`` resolves here and this
stub points the package search path at the real sources next door; every sub-module
(``dusty_gan_b200.models.dusty``, ``dusty_gan_b200.utils.lidar`` ...) is loaded from there.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "dusty-gan_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _fh
