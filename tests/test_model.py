"""CPU test: the executable specification of the parallel algorithms (the same maths the
CUDA kernels implement, kernel by kernel) agrees with the oracle."""
import numpy as np

import helpers
import model_gpu_algorithms as model


def _cases():
    rng = np.random.default_rng(1)
    out = []
    for n in (1, 2, 3, 5, 17, 64, 150):
        out += [(f"{k}_{n}", v) for k, v in helpers.families(n).items()]
    for _ in range(60):
        n = int(rng.integers(1, 160))
        s = int(rng.choice([1, 2, 3, 4, 256]))
        out.append(("rnd", rng.integers(0, s, size=n, dtype=np.uint8).tobytes()))
    return out


def test_chunked_lyndon_boundaries(oracle):
    for name, x in _cases():
        want = oracle.lyndon_starts(x).tolist()
        for chunk in (1, 4, 16, 64):
            got = model.lyndon_starts_chunked(np.frombuffer(x, np.uint8), chunk).tolist()
            assert got == want, (name, chunk)


def test_doubling_forward(oracle):
    for name, x in _cases():
        assert model.forward(x, chunk=8) == oracle.forward(x), name


def test_splitter_inverse(oracle):
    for name, x in _cases():
        for shift in (26, 30, 31):
            assert model.inverse(x, shift=shift) == oracle.inverse(x), (name, shift)
