"""CPU tests: the C-ABI library builds, loads and exports every symbol include/bpgpu.h declares;
without a GPU the product refuses to run (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for hdr, prefix in (("bpgpu.h", "bpgpu_"), ("bphost.h", "bph_")):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(" + prefix + r"[A-Za-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol(bp):
    if not os.path.exists(bp.library_path()):
        bp.build_library()
    lib = bp.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    from bulletproofs_amcl_b200 import binding
    assert sorted(s[0] for s in binding.SYMBOLS + binding.SYMBOLS_HOST) == names


def test_constants_without_gpu(bp):
    lib = bp.lib()
    assert lib.bpgpu_modbytes(bp.BLS12_381) == 48
    assert lib.bpgpu_modbytes(bp.BN254) == 32
    assert lib.bpgpu_modbytes(7) < 0
    assert lib.bpgpu_strerror(-4).decode().startswith("verification failed")
    assert 3 <= lib.bpgpu_msm_window_bits(1 << 20) <= 18


def test_no_cpu_fallback(bp):
    if bp.lib().bpgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(bp.BpgpuError):
        bp.Context(bp.BLS12_381, 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bulletproofs-amcl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
