"""CPU: the __host__ __device__ FFT building blocks of the kernels (in-register butterflies, the
conjugate-symmetric odd-prime DFTs, the mixed-radix Stockham stages with their skewed buffers and
multiply-shift index arithmetic) compiled for the host with nvcc and checked against a naive
double-precision DFT (tests/host/fft_host_check.cu).  No GPU involved."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_fft_building_blocks_on_host(tmp_path):
    exe = str(tmp_path / "fft_host_check")
    src = os.path.join(ROOT, "tests", "host", "fft_host_check.cu")
    cmd = [NVCC, "-O1", "-std=c++17", "-o", exe, src, "-I", os.path.join(ROOT, "babe_b200", "csrc"),
           "-I", os.path.join(ROOT, "include")]
    subprocess.run(cmd, check=True, capture_output=True, timeout=600)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=600).stdout
    rows = [l.split() for l in out.strip().splitlines()]
    assert len(rows) == 27
    for name, n, err in rows:
        assert float(err) < 2e-6, (name, n, err)        # fp32 transforms of <= 2048 points


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_band_core_passes_on_host(tmp_path):
    """BandCore<R3> (csrc/bandfft.cuh): the three passes the CQT band kernels call, emulated thread by
    thread on the host, for M = 256 ... 4096."""
    exe = str(tmp_path / "bandfft_host_check")
    src = os.path.join(ROOT, "tests", "host", "bandfft_host_check.cu")
    cmd = [NVCC, "-O1", "-std=c++17", "-o", exe, src, "-I", os.path.join(ROOT, "babe_b200", "csrc"),
           "-I", os.path.join(ROOT, "include")]
    subprocess.run(cmd, check=True, capture_output=True, timeout=600)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=600).stdout
    rows = [l.split() for l in out.strip().splitlines()]
    assert [int(r[1]) for r in rows] == [256, 512, 1024, 2048, 4096]
    for name, n, err in rows:
        assert float(err) < 2e-6, (name, n, err)


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_core3_passes_on_host(tmp_path):
    """Core3 (csrc/stft_cores.cuh): forward / mirror / inverse passes of the 4096-point core of
    k_apply_filter, k_stft_stats and k_fir_filter, emulated thread by thread on the host."""
    exe = str(tmp_path / "core3_host_check")
    src = os.path.join(ROOT, "tests", "host", "core3_host_check.cu")
    cmd = [NVCC, "-O1", "-std=c++17", "-o", exe, src, "-I", os.path.join(ROOT, "babe_b200", "csrc"),
           "-I", os.path.join(ROOT, "include")]
    subprocess.run(cmd, check=True, capture_output=True, timeout=600)
    out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=600).stdout
    vals = dict(l.split() for l in out.strip().splitlines())
    assert float(vals["fwd"]) < 5e-7 and float(vals["roundtrip"]) < 5e-7 and float(vals["mirror"]) == 0.0
