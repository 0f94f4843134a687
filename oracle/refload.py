"""Load the compiled reference extensions from oracle/_ref (built by oracle/build_ref.py) --
TEST INFRASTRUCTURE ONLY. Returns None when an artefact is absent (e.g. a fresh checkout)."""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


def load(name):
    """name in {"dustyref_cd", "dustyref_cd_o3", "dustyref_fps"} -> python extension module or None."""
    if name in _cache:
        return _cache[name]
    path = os.path.join(HERE, "_ref", name, name + ".so")
    mod = None
    if os.path.exists(path):
        import torch  # noqa: F401  (the extension links against libtorch)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod
