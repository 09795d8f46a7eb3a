"""The C-ABI library without a GPU: it loads, exports exactly the entry points include/shf_b200.h declares, the
device-free part of the buffer API behaves like the reference's STPFilterBuffer, and compute entry points fail loudly
(no CPU fallback)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "shf_b200.h")).read()
    return sorted(set(re.findall(r"SHF_API\s+[\w\s\*]+?\b(shf_\w+)\s*\(", text)))


def test_header_and_library_agree(shf):
    from superterrainplus_b200.api import C_ABI_SYMBOLS

    declared = declared_symbols()
    assert declared == sorted(C_ABI_SYMBOLS)
    lib = shf.library()
    for name in declared:
        assert getattr(lib, name) is not None
    exported = subprocess.run(["nm", "-D", "--defined-only", shf.library_path()], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r"\bT (shf_\w+)", exported)))
    assert exported == declared, "library exports differ from the header"


def test_library_is_sm100a_only(shf):
    out = subprocess.run(["cuobjdump", "--list-elf", shf.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_buffer_api_without_device(shf):
    FB = shf.STPSingleHistogramFilter.STPFilterBuffer
    for exec_type in (FB.STPExecutionType.Serial, FB.STPExecutionType.Parallel):
        buf = FB(exec_type)
        assert buf.type() == exec_type
        assert buf.size() == (0, 0)
        hist = buf.readHistogram()
        assert hist.Bin is None and hist.HistogramStartOffset is None
        buf.close()
    with pytest.raises(shf.STPInvalidEnum):
        FB(0x42)


def test_no_cpu_fallback(shf):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is for GPU-less hosts")
    with pytest.raises(shf.STPCUDAError):
        shf.STPSingleHistogramFilter()


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "superterrainplus_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text and "shf_oracle" not in text, f
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read()
    # the measurement scripts run the product and read reports; none of them may execute the checker either
    for f in os.listdir(os.path.join(ROOT, "scripts")):
        if f.endswith((".py", ".sh")):
            text = open(os.path.join(ROOT, "scripts", f), errors="replace").read()
            assert "import oracle" not in text and "from oracle" not in text, f"scripts/{f}"
