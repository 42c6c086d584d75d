"""TEST INFRASTRUCTURE — import pieces of the *Python* reference in place.

Only works where /root/reference exists (this container).  Used by
tests/golden/make_golden.py to regenerate the committed fixtures and by the
optional cross-checks in tests/ (skipped when the reference is absent, e.g. on
the GPU box).  Nothing is copied: a stub package object named ``kaolin`` is
given the reference directory as its ``__path__`` so that individual
sub-modules (camera maths, obj reader, the naive DefTet oracle, mask_iou, and
the reference's own rasterize/dibr wrappers) import without running
kaolin/__init__.py; ``kaolin._C`` and a few absent third-party packages are
replaced by inert stubs.
"""
import importlib
import os
import sys
import types

REF = os.environ.get("KAOLIN_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        s = _Stub(self.__name__ + "." + k)
        setattr(self, k, s)
        return s

    def __call__(self, *a, **k):
        return None


def available():
    return os.path.isdir(os.path.join(REF, "kaolin", "render", "mesh"))


def setup(c_module=None):
    """Install the stub ``kaolin`` package; ``c_module`` (optional) is an object
    exposing ``render.mesh.<op>`` that plays the role of ``kaolin._C``."""
    if not available():
        raise RuntimeError(f"reference not found under {REF}")
    if "kaolin" in sys.modules and not getattr(sys.modules["kaolin"], "_is_ref_stub", False):
        raise RuntimeError("a real 'kaolin' package is already imported")
    pkg = sys.modules.get("kaolin")
    if pkg is None:
        pkg = types.ModuleType("kaolin")
        pkg.__path__ = [os.path.join(REF, "kaolin")]
        pkg._is_ref_stub = True
        sys.modules["kaolin"] = pkg
        for name in ["wget", "pxr", "warp", "pygltflib", "plyfile", "usd", "tornado", "flask"]:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if c_module is None:
        c_module = _Stub("kaolin._C")
    sys.modules["kaolin._C"] = c_module
    pkg._C = c_module
    # modules that did `from kaolin import _C` at import time keep their own name
    for name in ("kaolin.render.mesh.rasterization", "kaolin.render.mesh.dibr",
                 "kaolin.render.mesh.deftet"):
        mod = sys.modules.get(name)
        if mod is not None:
            mod._C = c_module
    return pkg


def module(name):
    setup(sys.modules.get("kaolin._C") if "kaolin" in sys.modules else None)
    return importlib.import_module(name)
