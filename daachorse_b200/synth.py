"""Seeded synthetic workloads for the configs of BASELINE.json.

Everything is a pure function of its seed (numpy PCG64), so the GPU box, the build container
and the CPU oracle all see the same patterns and the same haystacks.

  C2  10 000 distinct lowercase-ASCII patterns, length clip(N(7, 2.5), 3, 12);
      262 144 haystacks x 256 B over [a-z ] with planted patterns.
  C3  675 000 distinct UTF-8 patterns "UniDic-like": 1-8 code points drawn Zipf(s) from a
      ~6 k-symbol kana + kanji (+ a few ASCII) alphabet; 1 Mi haystacks x 4 KiB.
  C4  100 000 distinct CJK patterns, 1-6 code points (charwise, LeftmostLongest); 512 Ki x 1 KiB,
      cut at char boundaries and padded with ASCII spaces.
  C5  1 000 000 patterns from the C3 generator; long records.

Large batches are materialised as fixed-size windows into a text *pool* (tens to hundreds of
MiB, generated here on the CPU): haystack i = pool[starts[i] : starts[i] + hay_len].  The pool
and the starts are seeded, so any haystack can be regenerated on the CPU for the oracle, and
the full batch is materialised on the GPU with one gather (see ``materialise_on_device``).
"""
import numpy as np


# ---- alphabets -------------------------------------------------------------------------------

def cjk_alphabet(n_symbols=6000, n_ascii=40):
    """Code points: a few ASCII, hiragana, katakana, then CJK unified ideographs."""
    ascii_part = list(range(0x30, 0x3A)) + list(range(0x61, 0x7B))
    ascii_part = ascii_part[:n_ascii]
    kana = list(range(0x3041, 0x3097)) + list(range(0x30A1, 0x30FB))
    rest = n_symbols - len(ascii_part) - len(kana)
    kanji = list(range(0x4E00, 0x4E00 + max(rest, 0)))
    return np.array(kana + kanji + ascii_part, dtype=np.uint32)[:n_symbols]


def zipf_probs(n, s):
    w = 1.0 / np.power(np.arange(1, n + 1, dtype=np.float64), s)
    return w / w.sum()


def utf8_encode(cps):
    """Vectorised UTF-8 encoding.  Returns (bytes uint8 array, per-code-point byte lengths)."""
    cps = np.asarray(cps, dtype=np.uint32)
    n1 = cps < 0x80
    n2 = (cps >= 0x80) & (cps < 0x800)
    n3 = (cps >= 0x800) & (cps < 0x10000)
    n4 = cps >= 0x10000
    lens = (n1 * 1 + n2 * 2 + n3 * 3 + n4 * 4).astype(np.int64)
    offs = np.zeros(len(cps) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    out = np.zeros(int(offs[-1]), dtype=np.uint8)
    o = offs[:-1]
    out[o[n1]] = cps[n1]
    out[o[n2]] = 0xC0 | (cps[n2] >> 6)
    out[o[n2] + 1] = 0x80 | (cps[n2] & 0x3F)
    out[o[n3]] = 0xE0 | (cps[n3] >> 12)
    out[o[n3] + 1] = 0x80 | ((cps[n3] >> 6) & 0x3F)
    out[o[n3] + 2] = 0x80 | (cps[n3] & 0x3F)
    out[o[n4]] = 0xF0 | (cps[n4] >> 18)
    out[o[n4] + 1] = 0x80 | ((cps[n4] >> 12) & 0x3F)
    out[o[n4] + 2] = 0x80 | ((cps[n4] >> 6) & 0x3F)
    out[o[n4] + 3] = 0x80 | (cps[n4] & 0x3F)
    return out, lens


class PatternSet:
    """Patterns as one byte blob + n+1 offsets (the layout the C ABI takes)."""

    def __init__(self, blob, offs):
        self.blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self.offs = np.ascontiguousarray(offs, dtype=np.uint64)

    def __len__(self):
        return len(self.offs) - 1

    def get(self, i):
        return self.blob[int(self.offs[i]):int(self.offs[i + 1])].tobytes()

    def as_list(self):
        b = self.blob.tobytes()
        o = self.offs
        return [b[int(o[i]):int(o[i + 1])] for i in range(len(self))]


def _distinct_rows(rows, lens, n, rng):
    """rows: (m, W) uint32 zero-padded symbol rows with lengths; keep n distinct, random order."""
    key = np.concatenate([lens[:, None].astype(np.uint32), rows], axis=1)
    _, first = np.unique(key, axis=0, return_index=True)
    first.sort()  # keep generation order among the distinct ones
    if len(first) < n:
        raise ValueError("generator produced only %d distinct patterns (< %d)" % (len(first), n))
    keep = first[:n]
    perm = rng.permutation(n)
    return rows[keep][perm], lens[keep][perm]


def _encode_rows(rows, lens):
    """(n, W) zero-padded code point rows -> PatternSet."""
    n, W = rows.shape
    flat = rows[np.arange(W)[None, :] < lens[:, None]]
    blob, blens = utf8_encode(flat)
    cp_offs = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=cp_offs[1:])
    csum = np.zeros(len(flat) + 1, dtype=np.int64)
    np.cumsum(blens, out=csum[1:])
    return PatternSet(blob, csum[cp_offs].astype(np.uint64))


# share of the dictionary at each length (code points), 1..8
CJK_LEVEL_SHARE = (0.0, 0.26, 0.28, 0.20, 0.12, 0.07, 0.04, 0.03)


def patterns_cjk(n, seed=1, max_cp=8, zipf_s=1.0, n_symbols=6000, level_share=CJK_LEVEL_SHARE):
    """C3 / C4 / C5 generator: n distinct UTF-8 "dictionary words" of 2..max_cp code points.

    Words are grown level by level like a morphological dictionary: a word of k+1 characters
    extends an existing k-character word by one Zipf-drawn character, so longer words share
    their prefixes with shorter ones (that sharing is what makes 675 k words fit in ~2 M
    double-array states, the UniDic scale of figures/memory.txt).  Registration order is a
    seeded shuffle."""
    rng = np.random.default_rng(seed)
    alpha = cjk_alphabet(n_symbols)
    cdf = np.cumsum(zipf_probs(len(alpha), zipf_s))
    share = np.array(level_share[:max_cp], dtype=np.float64)
    share /= share.sum()
    want = np.floor(share * n).astype(np.int64)
    want[1] += n - want.sum()
    W = max_cp
    levels = []
    prev = None
    for k in range(1, W):  # k+1 = word length in code points
        need = int(want[k])
        if need == 0:
            continue
        got = np.zeros((0, W), dtype=np.uint32)
        while len(got) < need:
            m = int((need - len(got)) * 1.3) + 64
            rows = np.zeros((m, W), dtype=np.uint32)
            if prev is None:
                for j in range(k):
                    rows[:, j] = alpha[np.searchsorted(cdf, rng.random(m))]
            else:
                rows[:, :k] = prev[rng.integers(0, len(prev), size=m)][:, :k]
            rows[:, k] = alpha[np.minimum(np.searchsorted(cdf, rng.random(m)), len(alpha) - 1)]
            got = np.unique(np.concatenate([got, rows]), axis=0)
        got = got[rng.permutation(len(got))[:need]]
        levels.append((got, k + 1))
        prev = got
    rows = np.concatenate([g for g, _ in levels])
    lens = np.concatenate([np.full(len(g), L, dtype=np.int64) for g, L in levels])
    perm = rng.permutation(len(rows))
    return _encode_rows(rows[perm], lens[perm])


def patterns_ascii(n=10000, seed=1, alphabet=b"abcdefghijklmnopqrstuvwxyz", mean=7.0, sd=2.5, lo=3, hi=12):
    """C2 generator: distinct lowercase strings, length clip(N(mean, sd), lo, hi)."""
    rng = np.random.default_rng(seed)
    alpha = np.frombuffer(alphabet, dtype=np.uint8).astype(np.uint32)
    m = int(n * 1.3) + 100
    lens = np.clip(np.rint(rng.normal(mean, sd, size=m)), lo, hi).astype(np.int64)
    rows = alpha[rng.integers(0, len(alpha), size=(m, hi))]
    rows[np.arange(hi)[None, :] >= lens[:, None]] = 0
    rows, lens = _distinct_rows(rows, lens, n, rng)
    flat = rows[np.arange(hi)[None, :] < lens[:, None]].astype(np.uint8)
    offs = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=offs[1:])
    return PatternSet(flat, offs)


# ---- text pools ---------------------------------------------------------------------------------

def _ragged_gather(blob, starts, lens):
    """Concatenate blob[starts[i] : starts[i]+lens[i]] for all i (vectorised; pure index
    arithmetic, so the torch/CUDA route below produces the same bytes as numpy)."""
    try:
        import torch

        if torch.cuda.is_available() and len(lens) > (1 << 20):
            dev = torch.device("cuda")
            tl = torch.from_numpy(np.ascontiguousarray(lens, dtype=np.int64)).to(dev)
            ts = torch.from_numpy(np.ascontiguousarray(starts, dtype=np.int64)).to(dev)
            tb = torch.from_numpy(np.ascontiguousarray(blob)).to(dev)
            offs = torch.zeros(len(lens) + 1, dtype=torch.int64, device=dev)
            torch.cumsum(tl, 0, out=offs[1:])
            seg = torch.repeat_interleave(torch.arange(len(lens), device=dev), tl)
            src = ts[seg] + (torch.arange(seg.numel(), device=dev) - offs[seg])
            return tb[src].cpu().numpy(), offs.cpu().numpy()
    except ImportError:
        pass
    total = int(lens.sum())
    out_offs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=out_offs[1:])
    idx = np.arange(total, dtype=np.int64)
    seg = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    src = starts[seg] + (idx - out_offs[seg])
    return blob[src], out_offs


def text_pool(patterns, n_bytes, seed=2, rho=0.05, alphabet_cps=None, alphabet_probs=None):
    """A pool of ~n_bytes of text: a stream of tokens, each either one alphabet symbol (Zipf) or,
    with probability rho, one whole dictionary pattern (uniform over the dictionary).
    Returns (uint8 array of exactly n_bytes, boundaries int64 array of token start offsets)."""
    rng = np.random.default_rng(seed)
    if alphabet_cps is None:
        alphabet_cps = cjk_alphabet()
    if alphabet_probs is None:
        alphabet_probs = zipf_probs(len(alphabet_cps), 1.0)
    sym_bytes, sym_lens = utf8_encode(alphabet_cps)
    sym_offs = np.zeros(len(alphabet_cps) + 1, dtype=np.int64)
    np.cumsum(sym_lens, out=sym_offs[1:])
    p_offs = patterns.offs.astype(np.int64)
    p_lens = np.diff(p_offs)
    mean_tok = (1 - rho) * float((sym_lens * alphabet_probs).sum()) + rho * float(p_lens.mean())
    parts, bounds, have = [], [], 0
    while have < n_bytes:
        t = int((n_bytes - have) / mean_tok * 1.05) + 64
        is_pat = rng.random(t) < rho
        sym = rng.choice(len(alphabet_cps), size=t, p=alphabet_probs)
        pat = rng.integers(0, len(patterns), size=t)
        starts = np.where(is_pat, p_offs[pat] + len(sym_bytes), sym_offs[sym])
        lens = np.where(is_pat, p_lens[pat], sym_lens[sym])
        both = np.concatenate([sym_bytes, patterns.blob])
        chunk, offs = _ragged_gather(both, starts, lens)
        parts.append(chunk)
        bounds.append(offs[:-1] + have)
        have += len(chunk)
    pool = np.concatenate(parts)[:n_bytes]
    b = np.concatenate(bounds)
    return np.ascontiguousarray(pool), b[b < n_bytes]


def window_starts(boundaries, pool_len, n, hay_len, seed=3):
    """n window starts, each on a token boundary, with start + hay_len <= pool_len."""
    rng = np.random.default_rng(seed)
    ok = boundaries[boundaries <= pool_len - hay_len]
    return ok[rng.integers(0, len(ok), size=n)].astype(np.int64)


def pad_to_char_boundary(rows, pad=0x20):
    """rows: (n, hay_len) uint8 windows that start on a char boundary.  A window may end inside a
    multi-byte char; the incomplete tail is overwritten with ASCII spaces (C4: "cut at char
    boundaries and padded with ASCII space").  Returns a new array."""
    rows = np.array(rows, dtype=np.uint8, copy=True)
    n, L = rows.shape
    done = np.zeros(n, dtype=bool)
    for t in range(1, 4):  # the lead byte of an incomplete char is among the last 3 bytes
        if t > L:
            break
        b = rows[:, L - t]
        is_cont = (b & 0xC0) == 0x80
        need = np.where(b < 0x80, 1, np.where(b < 0xE0, 2, np.where(b < 0xF0, 3, 4)))
        cut = (~done) & (~is_cont) & (need > t)
        for k in range(1, t + 1):
            rows[cut, L - k] = pad
        done |= ~is_cont
    return rows


def materialise_host(pool, starts, hay_len):
    """(n * hay_len) uint8 batch + offsets on the host (small n only)."""
    n = len(starts)
    text = np.empty(n * hay_len, dtype=np.uint8)
    rows = text.reshape(n, hay_len) if n else text.reshape(0, hay_len)
    step = max(1, (64 << 20) // max(hay_len, 1))  # index matrices of at most 64 Mi entries at a time
    ar = np.arange(hay_len, dtype=np.int64)[None, :]
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        rows[lo:hi] = pool[starts[lo:hi, None] + ar]
    offs = (np.arange(n + 1, dtype=np.uint64) * np.uint64(hay_len))
    return text, offs


def materialise_on_device(pool_t, starts_t, hay_len, chunk=1 << 16):
    """Device batch: pool_t (uint8 CUDA tensor), starts_t (int64 CUDA tensor) -> (n*hay_len,) uint8."""
    import torch

    n = starts_t.numel()
    out = torch.empty((n, hay_len), dtype=torch.uint8, device=pool_t.device)
    view = pool_t.unfold(0, hay_len, 1)
    chunk = max(1, min(chunk, (1 << 30) // hay_len))  # at most 1 GiB of gathered rows at a time
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        out[lo:hi] = view[starts_t[lo:hi]]
    offs = torch.arange(n + 1, dtype=torch.int64, device=pool_t.device) * hay_len
    return out.reshape(-1), offs


def pad_to_char_boundary_device(text_t, n, hay_len, pad=0x20):
    """pad_to_char_boundary on a device batch of n windows x hay_len bytes, in place."""
    import torch

    rows = text_t.view(n, hay_len)
    done = torch.zeros(n, dtype=torch.bool, device=rows.device)
    for t in range(1, 4):
        if t > hay_len:
            break
        bt = rows[:, hay_len - t]
        is_cont = (bt & 0xC0) == 0x80
        need = torch.where(bt < 0x80, 1, torch.where(bt < 0xE0, 2, torch.where(bt < 0xF0, 3, 4)))
        cut = (~done) & (~is_cont) & (need > t)
        for k in range(1, t + 1):
            rows[cut, hay_len - k] = pad
        done |= ~is_cont
    return text_t


# ---- named configs --------------------------------------------------------------------------------

def config(name, scale=1.0):
    """Parameters of BASELINE.json's configs.  ``scale`` < 1 shrinks the batch (tests)."""
    if name == "C2":
        return dict(name="C2", variant="bytewise", n_patterns=10000, pattern_gen="ascii", mode="find_overlapping_iter",
                    match_kind=0, n_haystacks=int(262144 * scale), hay_len=256, rho=0.05, pool_bytes=32 << 20)
    if name == "C3":
        return dict(name="C3", variant="bytewise", n_patterns=675000, pattern_gen="cjk", mode="find_overlapping_iter",
                    match_kind=0, n_haystacks=int((1 << 20) * scale), hay_len=4096, rho=0.02, text_zipf_s=0.5,
                    pool_bytes=128 << 20)
    if name == "C4":
        return dict(name="C4", variant="charwise", n_patterns=100000, pattern_gen="cjk6", mode="leftmost_find_iter",
                    match_kind=1, n_haystacks=int((1 << 19) * scale), hay_len=1024, rho=0.02, text_zipf_s=0.5,
                    pool_bytes=64 << 20)
    if name == "C5":
        return dict(name="C5", variant="bytewise", n_patterns=1000000, pattern_gen="cjk", mode="find_overlapping_iter",
                    match_kind=0, n_haystacks=int(12800 * scale), hay_len=1 << 20, rho=0.02, text_zipf_s=0.5,
                    pool_bytes=256 << 20)
    raise KeyError(name)


def make_patterns(cfg, n=None, seed=1):
    n = n or cfg["n_patterns"]
    if cfg["pattern_gen"] == "ascii":
        return patterns_ascii(n, seed)
    if cfg["pattern_gen"] == "cjk6":
        return patterns_cjk(n, seed, max_cp=6, level_share=(0.0, 0.46, 0.32, 0.14, 0.05, 0.03))
    return patterns_cjk(n, seed)


def make_pool(cfg, patterns, pool_bytes=None, seed=2):
    nb = pool_bytes or cfg["pool_bytes"]
    if cfg["pattern_gen"] == "ascii":
        cps = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz ", dtype=np.uint8).astype(np.uint32)
        return text_pool(patterns, nb, seed, cfg["rho"], cps, np.full(len(cps), 1.0 / len(cps)))
    cps = cjk_alphabet()
    return text_pool(patterns, nb, seed, cfg["rho"], cps, zipf_probs(len(cps), cfg.get("text_zipf_s", 1.0)))
