import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
EMU_SO = os.path.join(ROOT, "tests", "emu", "_build", "libeig_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs the read-only reference tree (build container only)")


def _emu_sources_newer():
    src = os.path.join(ROOT, "evolutionary_illusion_generator_b200", "csrc")
    if not os.path.isfile(EMU_SO):
        return True
    t = os.path.getmtime(EMU_SO)
    files = [os.path.join(src, f) for f in os.listdir(src)] + [os.path.join(ROOT, "tests", "emu", "cuda_emu.h"),
                                                                os.path.join(ROOT, "include", "eig.h")]
    return any(os.path.getmtime(f) > t for f in files)


@pytest.fixture(scope="session")
def emu_lib():
    """The kernel sources compiled with g++ against tests/emu/cuda_emu.h (TEST-ONLY, see that header)."""
    from evolutionary_illusion_generator_b200 import _lib
    if _emu_sources_newer():
        subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build_emu.sh")])
    return _lib.EigLibrary(EMU_SO)


@pytest.fixture(scope="session")
def gpu_engine_factory():
    """Engines bound to the real libeig.so on cuda:0 (fails loudly if the library or the GPU is missing)."""
    from evolutionary_illusion_generator_b200 import engine as E
    made = []

    def make(w, h, channels, max_genomes):
        e = E.Engine(w, h, channels, max_genomes)
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()


def ref_available():
    return os.path.isfile("/root/reference/generate_illusion.py")
