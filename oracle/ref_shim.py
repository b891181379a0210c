"""TEST INFRASTRUCTURE ONLY — loader for the UNMODIFIED reference package.

Imports ``vame`` from /root/reference (read-only mount, dev container only)
after registering stand-ins for the optional third-party packages the
reference imports at module scope but the hot path never calls
(matplotlib, umap, h5py, hmmlearn, ruamel.yaml).  Nothing in the reference is
edited or copied.  Used by ``gen_golden.py`` and ``tests/test_oracle_pinned.py``.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VAME_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vame"))


def _dummy(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def _install_shims():
    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return _Anything()

        def __getattr__(self, k):
            return _Anything()

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.cm",
                 "matplotlib.patches", "matplotlib.animation",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "umap", "h5py", "hmmlearn",
                 "hmmlearn.hmm", "networkx", "cv2"):
        try:
            importlib.import_module(name)
        except Exception:
            mod = _dummy(name)

            def _ga(k, _A=_Anything):
                if k.startswith("__"):
                    raise AttributeError(k)
                return _A()
            mod.__getattr__ = _ga  # type: ignore[attr-defined]
    try:
        importlib.import_module("ruamel.yaml")
    except Exception:
        import yaml as _pyyaml

        class YAML:
            def __init__(self, *a, **k):
                pass

            def load(self, f):
                return _pyyaml.safe_load(f)

            def dump(self, d, f):
                _pyyaml.safe_dump(dict(d), f, sort_keys=False)

        ru = _dummy("ruamel")
        ry = _dummy("ruamel.yaml", YAML=YAML)
        ru.yaml = ry


def load_reference():
    """Return the imported reference ``vame`` package (hot-path submodules loaded)."""
    if not reference_available():
        raise RuntimeError("reference not mounted at %s" % REFERENCE_ROOT)
    import torch  # noqa: F401  (import before the stand-ins are registered)
    import sklearn.cluster  # noqa: F401
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import vame  # noqa: F401
    rv = sys.modules["vame.model.rnn_vae"]
    # torch >= 2.4 removed ReduceLROnPlateau(verbose=...) which rnn_vae.py:337 passes.
    orig = rv.ReduceLROnPlateau
    if not getattr(orig, "_b200_wrapped", False):
        def _wrapped(*a, verbose=None, **k):
            return orig(*a, **k)
        _wrapped._b200_wrapped = True
        rv.ReduceLROnPlateau = _wrapped
    return vame


def reference_modules():
    """(rnn_model, rnn_vae, pose_segmentation) modules of the reference."""
    load_reference()
    return (sys.modules["vame.model.rnn_model"], sys.modules["vame.model.rnn_vae"],
            sys.modules["vame.analysis.pose_segmentation"])
