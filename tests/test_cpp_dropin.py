"""The C++ drop-in class (include/SuperAlgorithm+Host/STPSingleHistogramFilter.h over the C ABI): it compiles with
g++ everywhere; on a GPU it passes the reference's own test scenario re-expressed in tests/cpp/test_histogram.cpp."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BINARY = os.path.join(HERE, "cpp", "test_histogram")


def build():
    out = subprocess.run(["bash", os.path.join(HERE, "cpp", "build.sh")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr


def test_dropin_compiles_and_fails_loudly_without_gpu(shf):
    import torch

    shf.library()
    build()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    run = subprocess.run([BINARY], capture_output=True, text=True)
    assert run.returncode != 0
    assert "cudaGetDeviceCount" in run.stderr or "CUDA" in run.stderr


def test_device_free_surface_of_the_cpp_classes(shf):
    """tests/cpp/test_surface.cpp: copy / move rules, enum values, bin layout, exception hierarchy (static_asserts) and the
    behaviour of a fresh STPFilterBuffer (type / size echo, null views, moves, STPInvalidEnum); the filter constructor
    either succeeds (GPU host) or throws STPCUDAError (no CPU path)."""
    shf.library()
    build()
    run = subprocess.run([os.path.join(HERE, "cpp", "test_surface")], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "all C++ surface checks passed" in run.stdout


@pytest.mark.gpu
def test_reference_scenario_in_cpp(shf):
    shf.library()
    build()
    run = subprocess.run([BINARY], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "all C++ histogram checks passed" in run.stdout


BATCHER = os.path.join(HERE, "cpp", "test_batcher")


@pytest.mark.gpu
def test_concurrent_callers_through_the_batcher(shf):
    """five worker threads, one buffer each, two geometries in flight -- every caller gets the
    histogram a direct call gives (checked against the oracle), calls are coalesced into fewer device passes."""
    shf.library()
    build()
    run = subprocess.run([BATCHER], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "all batcher checks passed" in run.stdout
    print(run.stdout)


@pytest.mark.gpu
def test_device_biome_factory_feeds_the_filter_in_cpp(shf):
    """SURVEY.md section 8 row f4 through the C++ classes: STPBiomeFactoryDevice produces the maps in device memory,
    STPSingleHistogramFilter::filterDevice reads them in place; maps and histograms against the CPU restatements."""
    shf.library()
    build()
    run = subprocess.run([os.path.join(HERE, "cpp", "test_biome")], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "all C++ biome factory checks passed" in run.stdout
