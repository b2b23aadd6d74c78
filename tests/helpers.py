"""Shared test plumbing: oracle / generator / product loaders and input families.

The oracle (oracle/liboracle.so, oracle/_ref/*) is the CHECKER.  Nothing in the
product imports it.
"""
import ctypes
import hashlib
import importlib.util
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
PKG = REPO / "bijective-bwt_b200"
ORACLE_DIR = REPO / "oracle"
REF_DIR = ORACLE_DIR / "_ref"


def _u8(buf):
    return (ctypes.c_ubyte * len(buf)).from_buffer_copy(buf)


class Oracle:
    """ctypes view of oracle/liboracle.so (built on demand with gcc)."""

    def __init__(self):
        so = ORACLE_DIR / "liboracle.so"
        if not so.exists():
            subprocess.check_call(["make", "-C", str(ORACLE_DIR), "all"], stdout=subprocess.DEVNULL)
        self.lib = ctypes.CDLL(str(so))
        for name in ("oracle_bwts_forward", "oracle_bwts_inverse"):
            f = getattr(self.lib, name)
            f.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
            f.restype = ctypes.c_int
        self.lib.oracle_lyndon_starts.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        self.lib.oracle_lyndon_starts.restype = ctypes.c_long
        self.lib.oracle_suffix_array.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        self.lib.oracle_suffix_array.restype = ctypes.c_int
        self.lib.oracle_lf_map.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p]
        self.lib.oracle_lf_map.restype = ctypes.c_int

    def _run(self, fn, data):
        data = bytes(data)
        n = len(data)
        src = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(n, dtype=np.uint8)
        rc = fn(src.ctypes.data, n, out.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"oracle returned {rc}")
        return out.tobytes()

    def forward(self, data):
        return self._run(self.lib.oracle_bwts_forward, data)

    def inverse(self, data):
        return self._run(self.lib.oracle_bwts_inverse, data)

    def lyndon_starts(self, data):
        data = bytes(data)
        src = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(len(data), dtype=np.int32)
        cnt = self.lib.oracle_lyndon_starts(src.ctypes.data, len(data), out.ctypes.data)
        return out[:cnt].copy()

    def suffix_array(self, data):
        data = bytes(data)
        src = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(len(data), dtype=np.int32)
        rc = self.lib.oracle_suffix_array(src.ctypes.data if len(data) else None, len(data), out.ctypes.data)
        assert rc == 0
        return out

    def lf_map(self, data):
        data = bytes(data)
        src = np.frombuffer(data, dtype=np.uint8)
        out = np.empty(len(data), dtype=np.int32)
        rc = self.lib.oracle_lf_map(src.ctypes.data, len(data), out.ctypes.data)
        assert rc == 0
        return out


def ref_available():
    return all((REF_DIR / b).exists() for b in ("mk_bwts", "mbwt_new", "unbwts"))


def ref_run(tool, data):
    """Run an unmodified reference binary from oracle/_ref on `data`."""
    with tempfile.TemporaryDirectory() as td:
        src = Path(td) / "in"
        dst = Path(td) / "out"
        src.write_bytes(bytes(data))
        subprocess.check_call([str(REF_DIR / tool), str(src), str(dst)], stdout=subprocess.DEVNULL)
        return dst.read_bytes()


class Generator:
    """ctypes view of libbwts_gen.so (bijective-bwt_b200/host/gen_input.c)."""

    KINDS = {"random": 1, "text": 2, "tiled": 3, "dna": 4, "fibonacci": 6}

    def __init__(self):
        so = PKG / "libbwts_gen.so"
        if not so.exists():
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", str(so),
                                   str(PKG / "host" / "gen_input.c")])
        self.lib = ctypes.CDLL(str(so))
        self.lib.bwts_gen.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_long]
        self.lib.bwts_gen.restype = ctypes.c_int

    def make(self, kind, seed, n):
        out = np.empty(n, dtype=np.uint8)
        rc = self.lib.bwts_gen(self.KINDS[kind] if isinstance(kind, str) else kind, seed,
                               out.ctypes.data, n)
        assert rc == 0
        return out.tobytes()


def load_product():
    """Import bijective-bwt_b200/bwts_b200.py (the ctypes mirror of the C ABI)."""
    spec = importlib.util.spec_from_file_location("bwts_b200", PKG / "bwts_b200.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bwts_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def sha256(b):
    return hashlib.sha256(bytes(b)).hexdigest()


# ---------------------------------------------------------------------------
# input families
# ---------------------------------------------------------------------------

def fibonacci_word(n):
    a, b = b"a", b"ab"
    while len(b) < n:
        a, b = b, b + a
    return b[:n]


def thue_morse(n):
    out = bytearray(n)
    for i in range(n):
        out[i] = ord("a") + (bin(i).count("1") & 1)
    return bytes(out)


def de_bruijn(k, order):
    """de Bruijn sequence B(k, order) over the first k lowercase letters."""
    a = [0] * (k * order)
    seq = []

    def db(t, p):
        if t > order:
            if order % p == 0:
                seq.extend(a[1:p + 1])
        else:
            a[t] = a[t - p]
            db(t + 1, p)
            for j in range(a[t - p] + 1, k):
                a[t] = j
                db(t + 1, t)

    db(1, 1)
    return bytes(ord("a") + x for x in seq)


def families(n, seed=0):
    """name -> bytes of length n (exactly), adversarial shapes for BWTS."""
    rng = np.random.default_rng(seed + n)
    fam = {}
    fam["all_a"] = b"a" * n
    fam["descending"] = bytes((255 - (i * 256 // max(n, 1))) & 255 for i in range(n))
    fam["ascending"] = bytes((i * 256 // max(n, 1)) & 255 for i in range(n))
    fam["abab"] = (b"ab" * (n // 2 + 1))[:n]
    fam["abcabd"] = (b"abcabd" * (n // 6 + 1))[:n]
    half = rng.integers(0, 256, size=max(n // 2, 1), dtype=np.uint8).tobytes()
    fam["ww"] = (half + half + b"\x00")[:n]
    fam["fibonacci"] = fibonacci_word(n)
    fam["thue_morse"] = thue_morse(n)
    db = de_bruijn(4, 6)
    fam["de_bruijn"] = (db * (n // len(db) + 1))[:n]
    fam["xky"] = ((b"ab" * n)[: max(n - 1, 0)] + b"c")[:n]
    fam["edge_bytes"] = bytes(rng.choice(np.array([0x00, 0x7F, 0x80, 0xFF], dtype=np.uint8), size=n))
    fam["random256"] = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
    fam["random2"] = bytes(rng.integers(97, 99, size=n, dtype=np.uint8))
    fam["random4"] = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n))
    fam["runs"] = bytes(np.repeat(rng.integers(97, 101, size=n // 7 + 1, dtype=np.uint8), 7)[:n])
    fam["ab_descending_len"] = b"".join(b"a" + b"b" * k for k in range(60, 0, -1))[:n].ljust(n, b"a")
    return {k: v for k, v in fam.items() if len(v) == n}
