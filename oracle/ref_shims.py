"""ORACLE — TEST INFRASTRUCTURE ONLY (authoring container only).

Imports the UNMODIFIED reference (`nets.py`, `main.py`) under the shim list of
SURVEY.md §8c so its own classes can generate golden vectors, be timed by
`bench.py --impl reference`, and run their loops on swapped-in classes
(tests/test_gpu_boundary.py).  Source: `/root/reference` in the authoring container,
else the byte-identical copies `oracle/make_ref.py` put into the git-ignored
`oracle/_ref/` (which travels to the GPU box with the snapshot).  Never imported
by the product package.
"""
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CGS_REFERENCE", "/root/reference")
if not os.path.isfile(os.path.join(REF, "nets.py")):
    REF = os.path.join(_HERE, "_ref")


def available():
    return os.path.isfile(os.path.join(REF, "nets.py"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__file__ = f"<stub {name}>"
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load():
    """Return (nets, main) modules of the reference."""
    import numpy as np
    import torch  # noqa: F401  (must be imported before the stubs)
    import torchvision  # noqa: F401
    if "main" in sys.modules and getattr(sys.modules["main"], "__cgs_ref__", False):
        return sys.modules["nets"], sys.modules["main"]
    noop = lambda *a, **k: None
    ident = lambda a: a
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib")
        mpl.colors = _stub("matplotlib.colors", rgb_to_hsv=ident, hsv_to_rgb=ident)
        mpl.pyplot = _stub("matplotlib.pyplot", clf=noop, plot=noop, legend=noop, savefig=noop,
                           hist=noop, imsave=noop, figure=noop, close=noop, title=noop,
                           ylim=lambda *a, **k: (0, 1))
    for missing in ("minerl", "ffmpeg"):
        try:
            __import__(missing)
        except ImportError:
            _stub(missing)
    if "numpy.core.defchararray" not in sys.modules:
        try:
            import numpy.core.defchararray  # noqa: F401
        except Exception:
            _stub("numpy.core.defchararray", join=noop)
    for alias, typ in (("int", int), ("float", float), ("bool", bool)):
        if not hasattr(np, alias):
            setattr(np, alias, typ)
    from PIL import ImageFont
    _tt = ImageFont.truetype
    _default = ImageFont.load_default()

    def truetype(font=None, *a, **k):
        try:
            return _tt(font, *a, **k)
        except OSError:
            return _default
    ImageFont.truetype = truetype
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import nets
    import main
    main.__cgs_ref__ = True
    return nets, main


def make_handler(argv, workdir):
    """Build the reference `Handler` from `main()`'s own argparse defaults plus `argv`."""
    nets, main = load()
    captured = {}

    class _Capture(Exception):
        pass

    def fake_handler(args):
        captured["args"] = args
        raise _Capture()
    real = main.Handler
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = ["main.py"] + list(argv)
    main.Handler = fake_handler
    try:
        main.main()
    except _Capture:
        pass
    finally:
        main.Handler = real
        sys.argv = old_argv
    os.makedirs(workdir, exist_ok=True)
    os.chdir(workdir)
    try:
        H = real(captured["args"])
    finally:
        os.chdir(old_cwd)
    return H
