"""The C-ABI shared library builds without a GPU (nvcc cross-compiles sm_100a), loads, and exports every symbol
include/hig_b200.h declares; the ctypes prototypes cover exactly that set.  No compute calls here."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _declared():
    src = open(os.path.join(ROOT, "include", "hig_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hig_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import hig_b200  # noqa: F401
    from hig_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hig_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)
    assert lib.hig_version() >= 100
    assert lib.hig_launch_count() == 0 or lib.hig_launch_count() > 0


def test_ctypes_prototypes_match_header_arity_and_kinds():
    """Header / binding drift check: every prototype in include/hig_b200.h has as many parameters as its ctypes
    argtypes entry, pointers are bound as c_void_p and scalars as the matching C integer / float type."""
    import ctypes
    import hig_b200  # noqa: F401
    from hig_b200 import _lib
    src = open(os.path.join(ROOT, "include", "hig_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|unsigned long long|const char\*)\s+(hig_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert len(protos) == len(_lib.SIGNATURES)
    for name, params in protos:
        params = [p.strip() for p in params.split(",") if p.strip() and p.strip() != "void"]
        argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for p_decl, at in zip(params, argtypes):
            if "*" in p_decl:
                assert at is ctypes.c_void_p, (name, p_decl)
            elif p_decl.startswith("unsigned long long"):
                assert at is ctypes.c_ulonglong, (name, p_decl)
            elif p_decl.startswith("long long"):
                assert at is ctypes.c_longlong, (name, p_decl)
            elif p_decl.startswith("float"):
                assert at is ctypes.c_float, (name, p_decl)
            else:
                assert p_decl.startswith("int ") and at is ctypes.c_int, (name, p_decl)


def test_product_path_refuses_cpu_tensors():
    import pytest
    import torch
    import hig_b200  # noqa: F401
    from hig_b200 import ops
    from hig_b200.interaction_transformer import MotionInteractionTransformer
    with pytest.raises(RuntimeError):
        ops.gemm(torch.randn(8, 8).bfloat16(), torch.randn(8, 8).bfloat16())
    m = MotionInteractionTransformer(263, num_frames=196, num_layers=1, cap_id=True)
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 4, 263), torch.zeros(2, dtype=torch.long), length=[4, 4],
          text=[torch.tensor([1]), torch.tensor([2])])


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "human-interaction-generation_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S).replace("# oracle", ""), fn
