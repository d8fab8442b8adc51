"""
Load the UNMODIFIED reference package (accbpg) for use as a checker / CPU arm  --  TEST INFRASTRUCTURE.

Two places can hold it:
  * /root/reference (the build container; read-only), or
  * oracle/_ref/ (git-ignored; `python oracle/build_ref.py` copies the reference's accbpg/*.py there unmodified so the
    package travels to the GPU box with the snapshot, like a built .so would; nothing from it enters the history).
The reference imports cvxpy, jax and matplotlib at module scope although the hot path never touches them; they are
absent from the image, so permissive stub modules are registered first (recipe from SURVEY.md section 8c).

Only tests/, bench.py's CPU legs, oracle/gen_golden*.py and __graft_entry__ may import this file.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LOCAL = os.path.join(HERE, "_ref")
REF_ROOT = os.environ.get("ACCBPG_REFERENCE", "/root/reference")


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]          # used as a decorator: hand the function back
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__getattr__ = lambda attr: _Anything()
    for key, val in attrs.items():
        setattr(mod, key, val)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def register_stubs():
    _stub("cvxpy")
    jax = _stub("jax")
    jax.numpy = _stub("jax.numpy")
    jax.scipy = _stub("jax.scipy")
    jax.scipy.linalg = _stub("jax.scipy.linalg", cholesky=_Anything())
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot", __all__=[])


def locate():
    """Directory that contains the reference's `accbpg` package, or None."""
    for root in (REF_LOCAL, REF_ROOT):
        if os.path.exists(os.path.join(root, "accbpg", "algorithms.py")):
            return root
    return None


def import_reference(root=None):
    """Import the reference package with stubbed optional dependencies; raises ImportError when it is nowhere."""
    root = root or locate()
    if root is None:
        raise ImportError("the reference package is neither under oracle/_ref nor under " + REF_ROOT)
    register_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    import accbpg  # noqa
    return accbpg
