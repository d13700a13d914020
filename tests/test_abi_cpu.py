"""CPU-side checks of the C ABI: the shared library loads and exports every symbol the header
declares; without a GPU context creation fails loudly (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pcdgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcdgpu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import pcd_b200.lib as L
    if not os.path.exists(L.LIB_PATH):
        pytest.skip("libpcdgpu.so not built yet (run __graft_entry__.build())")
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(L.EXPORTS) == names


def test_no_cpu_fallback():
    import torch
    import pcd_b200.lib as L
    if not os.path.exists(L.LIB_PATH):
        with pytest.raises(ImportError):
            L.load()
        return
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(L.PcdGpuError) as e:
        L.Context(0)
    assert e.value.code == -2


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pcd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "c_oracle" not in text and "pcd_oracle" not in text and "liboracle" not in text, f


def test_domain_size_rule_matches_oracle():
    """GeneralEvaluationDomain::new: host-only entry point, no GPU needed"""
    import pcd_b200.lib as L
    if not os.path.exists(L.LIB_PATH):
        pytest.skip("libpcdgpu.so not built yet")
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as co
    for field in (0, 1):
        for m in (1, 2, 3, 1000, 1 << 16, (1 << 17) - 1, 1 << 17, (1 << 17) + 1, 200704, 200705, 7 << 17, (7 << 17) + 1,
                  49 << 17, (49 << 17) + 1, 1 << 20, (1 << 34) + 1):
            assert L.domain_size(field, m) == co.domain_size(field, m), (field, m)


def test_rust_sys_crate_declares_every_header_function():
    """rust/pcdgpu-sys/src/lib.rs (uncompiled source) is generated from include/pcdgpu.h: every exported function is
    declared, and the committed file is what tools/gen_rust_sys.py produces"""
    import subprocess
    import sys
    src = open(os.path.join(ROOT, "rust", "pcdgpu-sys", "src", "lib.rs")).read()
    for n in _header_functions():
        assert re.search(r"pub fn %s\(" % n, src), n
    assert subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py"), "--check"]) == 0
