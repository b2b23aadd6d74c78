#!/usr/bin/env python
"""Generate tests/golden/vectors.json from the UNMODIFIED reference binaries.

Run in the build container (where /root/reference exists):
    make -C oracle ref && python tests/golden/make_golden.py

Every vector is produced by oracle/_ref/mk_bwts (mk_bwts_sa.c), cross-checked
against oracle/_ref/mbwt_new (mk_bwts_sa_new.c) and inverted back with
oracle/_ref/unbwts (unbwts.c).  `inv` is unbwts applied to the INPUT read as an
arbitrary BWTS string (the transform is a bijection, so every string is valid).
Small vectors carry hex; large seeded ones carry the generator spec + SHA-256.
"""
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import helpers  # noqa: E402

KAT = [
    b"banana", b"^BANANA|", b"abracadabra", b"mississippi", b"a", b"aaaa", b"abab", b"ba",
    b"cbacbacba", b"zyxwv", b"abcabcabd", bytes.fromhex("00ff00ff80"), b"bananabanana",
]


def main():
    assert helpers.ref_available(), "run `make -C oracle ref` first"
    gen = helpers.Generator()
    vectors = []

    def add(name, data, spec=None):
        fwd = helpers.ref_run("mk_bwts", data)
        assert fwd == helpers.ref_run("mbwt_new", data), name
        assert helpers.ref_run("unbwts", fwd) == data, name
        inv = helpers.ref_run("unbwts", data)
        v = {"name": name, "n": len(data)}
        if spec is None:
            v.update(input=data.hex(), fwd=fwd.hex(), inv=inv.hex())
        else:
            v.update(spec=spec, input_sha256=helpers.sha256(data),
                     fwd_sha256=helpers.sha256(fwd), inv_sha256=helpers.sha256(inv))
        vectors.append(v)

    for i, k in enumerate(KAT):
        add(f"kat{i}", k)
    for n in (1, 2, 3, 7, 64, 255, 256, 257, 1000):
        for name, data in sorted(helpers.families(n).items()):
            add(f"{name}_{n}", data)
    for kind, seed, n in (("random", 1, 1 << 16), ("text", 2, 1 << 16), ("tiled", 3, 200_000),
                          ("dna", 4, 1 << 17), ("fibonacci", 0, 100_000),
                          ("random", 1, 1 << 20), ("text", 2, 1 << 20), ("dna", 4, 1 << 20),
                          ("tiled", 3, 3 << 20)):
        add(f"gen_{kind}_{seed}_{n}", gen.make(kind, seed, n), spec={"kind": kind, "seed": seed, "n": n})
    for name, data in sorted(helpers.families(1 << 16).items()):
        add(f"{name}_65536", data, spec={"family": name, "n": 1 << 16})

    (HERE / "vectors.json").write_text(json.dumps(vectors, indent=0))
    print(f"wrote {len(vectors)} vectors")


if __name__ == "__main__":
    main()
