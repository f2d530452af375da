"""The product's own device and host sources through the CPU emulation build (oracle/hostsim: test infrastructure - ga_kernels.cu
compiled by g++ against a stand-in CUDA runtime, same C ABI, loaded with GA_LIB) against the golden vectors: the host logic around
the kernels (batch planning, seed rounds, split launches, pipeline, result assembly, device-written mapping records) and the
kernels' arithmetic get exercised in the GPU-less container: goldens, differentials against the reference run here (cyclic graphs,
IUPAC reads, GFA overlaps, -B ramp redo), splitting, pipeline, seed rounds, the other BASELINE configs at test size.  The parity tests
proper are the -m gpu tests on the real library."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTSIM = os.path.join(ROOT, "oracle", "_ref", "libga_hostsim.so")


@pytest.mark.skipif(not os.path.exists(HOSTSIM), reason="oracle/_ref/libga_hostsim.so not built")
def test_goldens_and_host_logic_through_the_cpu_emulation():
    # a fresh interpreter: the library path is fixed when graphaligner_b200.api is first imported
    env = dict(os.environ, GA_LIB=HOSTSIM)
    # everything but the tests that need the device itself (page-locked host memory, the device-against-emulation comparison)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"), os.path.join(ROOT, "tests", "test_other_configs_gpu.py"),
                        "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider", "-k", "not page_locked and not big_batch"],
                       capture_output=True, text=True, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-500:]
    assert " passed" in r.stdout and "failed" not in r.stdout
