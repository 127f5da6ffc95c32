import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_SO = os.path.join(ROOT, "tests", "simt_emu", "libb200foam_emu.so")
CSRC = os.path.join(ROOT, "shallowinterfoam_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


_LIBS = {}


def get_lib(kind):
    """kind: 'gpu' = the product libb200foam.so (nvcc sm_100a) ; 'emu' = tests/simt_emu build of the same kernel
    sources under the SIMT emulator (test infrastructure only, exercised in the GPU-less container)."""
    from shallowinterfoam_b200.capi import B200Lib, load
    if kind not in _LIBS:
        if kind == "gpu":
            lib = load()
            lib.init(0, 0, 1)
        else:
            subprocess.check_call(["make", "-C", CSRC, "-j8", "emu"], stdout=subprocess.DEVNULL)
            lib = B200Lib(EMU_SO)
            lib.init(0, 0, 1)
        _LIBS[kind] = lib
    return _LIBS[kind]


BACKENDS = [pytest.param("emu", id="emu"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


@pytest.fixture(params=BACKENDS)
def lib(request):
    return get_lib(request.param)


@pytest.fixture(scope="session")
def oracle():
    from oracle import foam_oracle
    foam_oracle.lib()
    return foam_oracle
