"""Brute-force statement of the BWTS definition, small n only.

Independent of both the reference's SA fix-up and the GPU's prefix doubling:
Duval factorisation, then every rotation of every factor sorted by comparing
the infinite periodic strings (Fine-Wilf: |u|+|v| characters decide), output =
the byte cyclically preceding each rotation inside its own factor.
"""
from functools import cmp_to_key


def duval(s):
    """Start offsets of the Lyndon factors of s (bytes)."""
    n, f, out = len(s), 0, []
    while f < n:
        i, k = f, f + 1
        while k < n and s[i] <= s[k]:
            i = f if s[i] < s[k] else i + 1
            k += 1
        p = k - i
        while f <= i:
            out.append(f)
            f += p
    return out


def forward(s):
    s = bytes(s)
    n = len(s)
    starts = duval(s) + [n]
    rots = []  # (factor start, factor length, offset)
    for a, b in zip(starts[:-1], starts[1:]):
        for o in range(b - a):
            rots.append((a, b - a, o))

    def cmp(x, y):
        (a, la, oa), (b, lb, ob) = x, y
        for t in range(la + lb):
            ca = s[a + (oa + t) % la]
            cb = s[b + (ob + t) % lb]
            if ca != cb:
                return -1 if ca < cb else 1
        return 0

    rots.sort(key=cmp_to_key(cmp))
    return bytes(s[a + (o - 1) % l] for a, l, o in rots)


def inverse(b):
    """Closed form of the reference's cycle walk (unbwts.c:62-86)."""
    b = bytes(b)
    n = len(b)
    order = sorted(range(n), key=lambda i: (b[i], i))
    lf = [0] * n
    for r, i in enumerate(order):
        lf[i] = r
    out = bytearray(n)
    seen = [False] * n
    w = n - 1
    for st in range(n):
        if seen[st]:
            continue
        p = st
        while not seen[p]:
            seen[p] = True
            out[w] = b[p]
            w -= 1
            p = lf[p]
    return bytes(out)
