import os
import random
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): a fresh checkout has none.  Build it before collection
    (the test modules import vimz_b200, which refuses to load without it) -- the same `make` target __graft_entry__.build() uses."""
    lib = os.path.join(ROOT, "vimz_b200", "libvimz_gpu.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-j8", "vimz_b200/libvimz_gpu.so"], cwd=ROOT, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def coracle():
    from oracle import c
    return c()


_engines = {}


@pytest.fixture(scope="session", params=["direct", "buckets"])
def engines(request):
    """curve name -> vimz_b200.Engine on cuda:0 (created lazily, shared by the session).  Every GPU test runs twice:
    with short commitment keys on the direct multiples table (the default, option msm_direct_max = 32768) and with the
    bucket pipeline forced for every key (msm_direct_max = 0), so both MSM paths see all the small-size cases."""
    import vimz_b200

    class Lazy(dict):
        def __missing__(self, name):
            e = vimz_b200.Engine(name, 0)
            if request.param == "buckets":
                e.set_option("msm_direct_max", 0)
            self[name] = e
            return e

    d = Lazy()
    yield d
    for e in d.values():
        e.close()


def make_bases(curve, n, seed):
    """n affine points k_i*G with known discrete logs (python big-int oracle)."""
    from oracle import pyref as P
    rng = random.Random(seed)
    G = P.generator(curve)
    ks = [rng.randrange(1, curve.q) for _ in range(n)]
    # walk: P_i = P_{i-1} + D keeps this O(n) additions instead of n scalar muls
    pts = []
    k0, dk = ks[0], rng.randrange(1, curve.q)
    cur = P.scalar_mul(curve, k0, G)
    D = P.scalar_mul(curve, dk, G)
    logs = []
    for i in range(n):
        pts.append(cur)
        logs.append((k0 + i * dk) % curve.q)
        cur = P.aff_add(curve, cur, D)
    return pts, logs
