"""Seeded test inputs shared by the CPU (emulation) and GPU parity tests."""
import numpy as np

MIXED = ["a", "b", "é", "あ", "𝄞", "c", "ß", "漢"]


def mixed_width_case(seed):
    """Charwise patterns / haystacks over chars of 1-4 bytes, empty patterns included; haystacks also
    contain unmapped chars."""
    rng = np.random.default_rng(5000 + seed)
    kind = seed % 3
    alpha = int(rng.integers(1, 8))
    npat, mx = int(rng.integers(1, 80)), int(rng.integers(1, 9))
    pats = ["".join(MIXED[int(i)] for i in rng.integers(0, alpha, size=int(rng.integers(0 if seed % 5 == 0 else 1, mx + 1))))
            for _ in range(npat)]
    syms = MIXED[:alpha] + (["z", "語"] if seed % 3 else [])
    hays = ["".join(syms[int(i)] for i in rng.integers(0, len(syms), size=int(rng.integers(0, 200)))).encode()
            for _ in range(60)]
    offs = np.zeros(len(hays) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(h) for h in hays])
    return kind, pats, np.frombuffer(b"".join(hays), dtype=np.uint8), offs
