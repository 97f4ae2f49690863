"""The C-ABI library loads on a CPU-only box, exports everything include/radk.h declares, and
refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "radk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(radk_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(hr):
    lib = ctypes.CDLL(hr.library_path())
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/radk.h but not exported by libradk.so"
    assert sorted(hr.api.EXPORTS) == names


def test_version_and_defaults(hr):
    lib = hr.load_library()
    assert lib.radk_version() == 1
    o = hr.default_opts()
    assert (o.converge, o.parallel_src, o.tri_limit, o.area_limit, o.world) == (1200.0, 1, 1000000, 5.0, 1)
    assert abs(o.delta_ff - 0.000005) < 1e-12


def test_no_cpu_fallback(hr, models_dir):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = hr.Model(str(models_dir / "model2s.tri"))
    with pytest.raises(hr.RadkError) as e:
        hr.Context(m)
    assert e.value.code == 4 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "hw-radiosity_b200")
    for d, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(d, fn), errors="replace").read()
                assert "oracle/" not in text.replace("oracle/ref_harness/harness_api.h", "") and "pyoracle" not in text \
                    and "hostemu" not in text.replace("tests/hostemu", ""), fn
