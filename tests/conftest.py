import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU should fail loudly rather than silently pass; plain runs skip them.
    if _has_gpu():
        return
    markexpr = config.getoption("-m") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_fwd_golden.npz"))
    meta = {}
    for line in g["meta"]:
        parts = str(line).split(",")
        meta[parts[0]] = parts[1:]
    return g, meta


@pytest.fixture(scope="session")
def golden_u64():
    """Outputs of the reference's own code on 50-/60-/62-/63-bit primes (tests/golden/make_golden.py: make_u64).
    Returns (npz, cases): cases[name] = dict(N, q, psi, frames, seed_in, seed_in2, lazy, sha256)."""
    import numpy as np
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_fwd_golden_u64.npz"))
    cases = {}
    for line in g["meta"]:
        name, N, q, psi, frames, s1, s2, lazy, sha = str(line).split(",")
        cases[name] = dict(N=int(N), q=int(q), psi=int(psi), frames=int(frames), seed_in=int(s1), seed_in2=int(s2),
                           lazy=bool(int(lazy)), sha256=sha)
    return g, cases
