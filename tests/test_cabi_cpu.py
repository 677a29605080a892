"""CPU: the C-ABI library builds/loads without a GPU and exports exactly what include/mmf_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mmf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(?:int|int64_t|void)\s+(mmf_[a-z0-9_]+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    from incomplete_multimodal_fusion_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert declared, "no symbols parsed from the header"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.mmf_abi_version() == 1


def test_struct_layouts_match_header_field_order():
    """ctypes mirrors must list the same fields in the same order as the C structs."""
    from incomplete_multimodal_fusion_b200 import _lib
    src = open(os.path.join(ROOT, "include", "mmf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for cname, cls in (("MmfGemmArgs", _lib.GemmArgs), ("MmfAttnArgs", _lib.AttnArgs),
                       ("MmfSlotAttnArgs", _lib.SlotAttnArgs), ("MmfPoolAttnArgs", _lib.PoolAttnArgs),
                       ("MmfAdamWTensor", _lib.AdamWTensor)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, flags=re.S).group(1)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert names == [f[0] for f in cls._fields_], (cname, names, [f[0] for f in cls._fields_])


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from incomplete_multimodal_fusion_b200 import kernels
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kernels.gemm(a, a, torch.zeros(8, 8))
