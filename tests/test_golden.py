"""Golden vectors generated from the compiled reference (tests/golden/make_golden.py) and committed:
the oracle must reproduce them on the CPU, the CUDA library on the GPU.  This pins both even where
oracle/_ref is absent."""
import json
import os
import sys

import pytest

import support as S

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_golden as G  # noqa: E402

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))


def check(name, make_engine):
    c = next(c for c in G.CASES if c["name"] == name)
    got = G.run_case(c, make_engine)
    want = GOLDEN[name]
    assert got["scalars"] == want["scalars"], name
    for key in ("analog", "inp", "out"):
        assert got[key] == want[key], "%s: %s digest differs from the reference's" % (name, key)


@pytest.mark.parametrize("name", [c["name"] for c in G.CASES])
def test_oracle_reproduces_reference_goldens(name):
    check(name, lambda v, w, h, f: S.OracleEngine(v, w, h, f, seed=1))


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c["name"] for c in G.CASES])
def test_cuda_library_reproduces_reference_goldens(name):
    import ctypes as C
    C.CDLL(None).srand(1)  # the VHS drop-in draws from libc rand() like the reference
    check(name, lambda v, w, h, f: S.ProductEngine(v, w, h, f))
