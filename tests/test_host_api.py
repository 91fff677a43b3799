"""Host side of the product (construction, wire format, C ABI surface) -- no GPU needed.

The product's builder (daachorse_b200/csrc/host_build.cpp) and the oracle's builder
(oracle/dach_oracle.c) are two independent restatements of the reference; the crate's wire
format is the common ground, so "byte-identical serialize()" is the parity check here.
"""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import daachorse_b200 as D
import oracle_api as O
from daachorse_b200 import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "search_tests.json"), encoding="utf-8"))


def product(pats, cw=False, kind=0, nfb=16, values=None):
    B = D.CharwiseDoubleArrayAhoCorasickBuilder if cw else D.DoubleArrayAhoCorasickBuilder
    b = B.new().match_kind(kind).num_free_blocks(nfb)
    return b.build(pats) if values is None else b.build_with_values(list(zip(pats, values)))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "daachorse_b200.h")).read()
    declared = set(re.findall(r"^(?:int|size_t|void|uint8_t|uint32_t|uint64_t|double|const char \*)\s*\*?(dach_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert declared == set(_lib.SYMBOLS)
    L = C.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(L, s), s
    assert _lib.load().dach_abi_version() == 2


@pytest.mark.parametrize("cw", [False, True])
@pytest.mark.parametrize("kind", [0, 1, 2])
def test_golden_pattern_sets_serialize_identically(cw, kind):
    seen = set()
    for g in GOLD["groups"].values():
        for t in g:
            key = tuple(t["patterns"])
            if key in seen:
                continue
            seen.add(key)
            a = product(t["patterns"], cw, kind).serialize()
            b = O.OraclePma.build(t["patterns"], charwise=cw, match_kind=kind).serialize()
            assert a == b, t["patterns"]


def _rand_sets(rng, cw):
    n = int(rng.integers(1, 400))
    alpha = int(rng.integers(2, 40))
    maxlen = int(rng.integers(1, 12))
    pats = []
    for _ in range(n):
        L = int(rng.integers(0, maxlen + 1))
        if cw:
            pats.append("".join(chr(0x3041 + int(x)) if x % 3 else chr(97 + int(x) % 26)
                                for x in rng.integers(0, alpha, size=L)))
        else:
            pats.append(bytes(rng.integers(0, 256 if alpha > 30 else 97 + alpha, size=L).astype(np.uint8).tolist()))
    return pats


@pytest.mark.parametrize("seed", range(12))
@pytest.mark.parametrize("cw", [False, True])
def test_random_sets_serialize_identically(seed, cw):
    rng = np.random.default_rng(seed * 2 + cw)
    pats = _rand_sets(rng, cw)
    kind = seed % 3
    nfb = [16, 1, 2, 3][seed % 4]
    values = None if seed % 2 else rng.integers(0, 2**32, size=len(pats), dtype=np.uint64).tolist()
    a = product(pats, cw, kind, nfb, values)
    b = O.OraclePma.build(pats, charwise=cw, match_kind=kind, num_free_blocks=nfb, values=values)
    assert a.serialize() == b.serialize()
    assert (a.num_states(), a.heap_bytes(), a.num_elements()) == (b.num_states(), b.heap_bytes(), b.num_elements())
    assert a.max_pattern_len() == b.max_pattern_len()


def test_many_blocks_ring_closing():
    """> 16 blocks of 256 slots so that blocks drop out of the vacancy ring
    (src/build_helper.rs:138-146, src/bytewise/builder.rs:377-379)."""
    rng = np.random.default_rng(3)
    pats = [bytes(rng.integers(97, 123, size=int(rng.integers(3, 10))).astype(np.uint8).tolist()) for _ in range(4000)]
    for nfb in (1, 4, 16):
        a = product(pats, False, 0, nfb)
        b = O.OraclePma.build(pats, num_free_blocks=nfb)
        assert a.num_elements() > 16 * 256
        assert a.serialize() == b.serialize()
    cpats = ["".join(chr(0x4E00 + int(x)) for x in rng.integers(0, 300, size=int(rng.integers(1, 5)))) for _ in range(3000)]
    for nfb in (1, 16):
        assert product(cpats, True, 1, nfb).serialize() == O.OraclePma.build(
            cpats, charwise=True, match_kind=1, num_free_blocks=nfb).serialize()


def test_doc_constants():
    """src/bytewise.rs:761,782; src/charwise.rs:793,810."""
    p = D.DoubleArrayAhoCorasick.new(["bcd", "ab", "a"])
    assert (p.heap_bytes(), p.num_states(), p.match_kind()) == (4132, 6, D.MatchKind.Standard)
    c = D.CharwiseDoubleArrayAhoCorasick.new(["bcd", "ab", "a"])
    assert (c.heap_bytes(), c.num_elements(), c.num_states()) == (568, 8, 6)


def test_deserialize_roundtrip_and_rest():
    for cls, pats in ((D.DoubleArrayAhoCorasick, ["abba", "baaba", "ababa"]),
                      (D.CharwiseDoubleArrayAhoCorasick, ["全世界", "世界", "に"])):
        p = cls.new(pats)
        blob = p.serialize()
        q, rest = cls.deserialize(blob + b"tail")
        assert rest == b"tail"
        assert q.serialize() == blob
        assert q.num_states() == p.num_states()


def test_deserialize_rejects_invalid():
    """src/bytewise.rs:1495-1507, src/charwise.rs:1502-1514 + truncation + allocation guard."""
    for cls in (D.DoubleArrayAhoCorasick, D.CharwiseDoubleArrayAhoCorasick):
        with pytest.raises(D.DaachorseError) as e:
            cls.deserialize(bytes(21))
        assert e.value.code == _lib.INVALID_AUTOMATON
        blob = cls.new(["abc", "b"]).serialize()
        with pytest.raises(D.DaachorseError):
            cls.deserialize(blob[:-1])
        with pytest.raises(D.DaachorseError):
            cls.deserialize(b"\xff\xff\xff\xff" + bytes(64))
    # corrupt one field at a time: base / fail / output_pos out of range, parent not smaller
    blob = bytearray(D.DoubleArrayAhoCorasick.new(["ab", "b"]).serialize())
    n = int.from_bytes(blob[0:4], "little")
    for field_off in (0, 4):
        bad = bytearray(blob)
        bad[4 + 12 * 5 + field_off: 4 + 12 * 5 + field_off + 4] = (n + 5).to_bytes(4, "little")
        with pytest.raises(D.DaachorseError):
            D.DoubleArrayAhoCorasick.deserialize(bytes(bad))
    bad = bytearray(blob)
    outs = 4 + 12 * n + 4 + 4 + 4
    bad[outs + 8: outs + 12] = (1).to_bytes(4, "little")  # outputs[0].parent = 1 (not < its index)
    with pytest.raises(D.DaachorseError):
        D.DoubleArrayAhoCorasick.deserialize(bytes(bad))


def test_builder_errors():
    """tests/invalid_option_test.rs:1-9 and the builder's argument checks."""
    with pytest.raises(D.DaachorseError) as e:
        D.DoubleArrayAhoCorasickBuilder.new().num_free_blocks(0xFFFFFFFF).build(["pattern"])
    assert e.value.code == _lib.AUTOMATON_SCALE
    with pytest.raises(AssertionError):
        D.DoubleArrayAhoCorasickBuilder.new().num_free_blocks(0)
    with pytest.raises(D.DaachorseError) as e:
        D.DoubleArrayAhoCorasick.with_values([("a", 2**32)])
    assert e.value.code == _lib.INVALID_CONVERSION
    with pytest.raises(Exception):
        D.CharwiseDoubleArrayAhoCorasick.new([b"\xff\xfe"])  # not UTF-8
    L = _lib.load()
    h = C.c_void_p()
    offs = np.array([0, 2], dtype=np.uint64)
    blob = np.frombuffer(b"\xe3\x81", dtype=np.uint8)  # truncated UTF-8
    rc = L.dach_charwise_build(C.c_void_p(blob.ctypes.data), C.c_void_p(offs.ctypes.data), None, 1, 0, 16, C.byref(h))
    assert rc == _lib.INVALID_ARGUMENT
    rc = L.dach_bytewise_build(C.c_void_p(blob.ctypes.data), C.c_void_p(offs.ctypes.data), None, 1, 7, 16, C.byref(h))
    assert rc == _lib.INVALID_ARGUMENT


def test_match_kind_gating_before_any_device_work():
    """tests/matchkind_mismatch_test.rs / _charwise_test.rs: the crate panics when the iterator is
    created; here an AssertionError is raised before the device is touched."""
    for B in (D.DoubleArrayAhoCorasickBuilder, D.CharwiseDoubleArrayAhoCorasickBuilder):
        for kind in (D.MatchKind.LeftmostLongest, D.MatchKind.LeftmostFirst):
            p = B.new().match_kind(kind).build(["pattern"])
            for f in (p.find_iter, p.find_overlapping_iter, p.find_overlapping_no_suffix_iter):
                with pytest.raises(AssertionError):
                    f("")
        p = B.new().build(["pattern"])
        with pytest.raises(AssertionError):
            p.leftmost_find_iter("")


def test_scan_without_gpu_fails_loudly():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = D.DoubleArrayAhoCorasick.new(["a"])
    with pytest.raises(D.DaachorseError) as e:
        list(p.find_iter("a"))
    assert e.value.code == _lib.CUDA_ERROR


def test_upload_rejects_fail_cycle():
    """A failure chain that never reaches ROOT would spin the kernel (the crate documents the
    same hazard, src/bytewise.rs:824-830); the image builder refuses it.  Checked through the
    CPU emulation harness, which runs the same build_image()."""
    import emu_api as E

    blob = bytearray(D.DoubleArrayAhoCorasick.new(["ab", "b"]).serialize())
    n = int.from_bytes(blob[0:4], "little")
    # find two live slots and point their fails at each other
    live = [i for i in range(2, n) if int.from_bytes(blob[4 + 12 * i: 8 + 12 * i], "little") != 0
            or (int.from_bytes(blob[12 + 12 * i: 16 + 12 * i], "little") >> 8)]
    a, b = live[0], live[1]
    blob[8 + 12 * a: 12 + 12 * a] = b.to_bytes(4, "little")
    blob[8 + 12 * b: 12 + 12 * b] = a.to_bytes(4, "little")
    rc, *_ = E.scan(bytes(blob), False, 1, np.zeros(0, np.uint8), np.array([0, 0], dtype=np.uint64))
    assert rc == _lib.INVALID_AUTOMATON
