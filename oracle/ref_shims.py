"""TEST INFRASTRUCTURE — import the *unmodified* reference (read-only, /root/reference/codes) in THIS container.

Only used by oracle/make_golden.py and the CPU tests that pin oracle/ against the real reference; nothing on the
GPU box reads /root/reference (it does not exist there).  The reference needs third-party modules that are not
installed offline; they are replaced by inert stubs before import:

  clip        (OpenAI CLIP, un-vendored, unpinned)   -> stub module; cap_id models never call it, text-path
                                                        tests pass xf_proj/xf_out explicitly
  mmcv        (mmcv-full 1.3.17-1.5.3, un-vendored)  -> get_dist_info, Registry, build_from_cfg, DDP alias
  matplotlib                                         -> empty modules
  np.float / np.int (removed in numpy >= 1.24)       -> aliases
"""
import os
import sys
import types

REFERENCE_CODES = os.environ.get("HIG_REFERENCE_CODES", "/root/reference/codes")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_CODES, "models"))


def install_shims():
    import numpy as np
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "int"):
        np.int = int

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def _no_clip(*a, **k):
        raise RuntimeError("clip is stubbed: the real OpenAI CLIP package/weights are not available offline")

    try:
        import clip  # noqa: F401
    except Exception:
        stub("clip", load=_no_clip, tokenize=_no_clip)

    try:
        import mmcv  # noqa: F401
    except Exception:
        import torch

        class Registry:
            def __init__(self, name):
                self.name = name

            def register_module(self, *a, **k):
                return lambda cls: cls

        mm = stub("mmcv")
        mm.runner = stub("mmcv.runner", get_dist_info=lambda: (0, 1))
        mm.utils = stub("mmcv.utils", Registry=Registry, build_from_cfg=lambda *a, **k: None)
        mm.parallel = stub("mmcv.parallel", MMDistributedDataParallel=torch.nn.parallel.DistributedDataParallel,
                           collate=None)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = stub("matplotlib", use=lambda *a, **k: None)
        mpl.pyplot = stub("matplotlib.pyplot")
        for sub in ("animation", "patches", "colors", "cm"):
            setattr(mpl, sub, stub("matplotlib." + sub, FuncAnimation=None, FFMpegFileWriter=None))
        stub("mpl_toolkits")
        stub("mpl_toolkits.mplot3d", Axes3D=None)
        stub("mpl_toolkits.mplot3d.art3d", Poly3DCollection=None)
        sys.modules["mpl_toolkits.mplot3d"].art3d = sys.modules["mpl_toolkits.mplot3d.art3d"]


def import_reference():
    """Returns (interaction_transformer module, gaussian_diffusion module) of the real reference."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_CODES}")
    install_shims()
    if REFERENCE_CODES not in sys.path:
        sys.path.insert(0, REFERENCE_CODES)
    import importlib
    it = importlib.import_module("models.interaction_transformer")
    gd = importlib.import_module("models.gaussian_diffusion")
    return it, gd


def import_reference_trainer():
    it, gd = import_reference()
    import importlib
    tr = importlib.import_module("trainers.mul_ddpm_trainer")
    return tr
