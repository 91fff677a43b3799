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


def nul_heavy_case(kind, n_patterns=120000, n_hay=1500, max_len=600):
    """Binary patterns of 3..12 bytes and haystacks drawn mostly from 0x00: with more than ~16k patterns the
    automaton is larger than the default hot region of the compact image, so many states sit in the shifted
    part next to the holes re-placed families leave behind (dev_image.cpp)."""
    rng = np.random.default_rng(31337 + kind)
    pats = sorted(set(bytes(rng.choice([0, 0, 1, 2, 3, 255], size=int(rng.integers(3, 13))).tolist())
                      for _ in range(n_patterns)))
    lens = rng.integers(0, max_len, size=n_hay)
    offs = np.zeros(n_hay + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    text = rng.choice(np.array([0, 0, 0, 1, 2, 3, 255], dtype=np.uint8), size=int(offs[-1]))
    return pats, np.ascontiguousarray(text), offs
