"""The kernels' per-lane logic (daachorse_b200/csrc/scan_lane.cuh), compiled for the CPU by
tests/emu, against the oracle and the reference's golden vectors.  No GPU needed; the same
checks run against the real kernels in test_gpu_parity.py."""
import json
import os

import numpy as np
import pytest

import emu_api as E
from cases import damage_standard_wire, hand_made_case, mixed_width_case, nul_heavy_case
import oracle_api as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "search_tests.json"), encoding="utf-8"))
MODE = {"find_iter": 0, "find_overlapping_iter": 1, "find_overlapping_no_suffix_iter": 2, "leftmost_find_iter": 3}
ORC_MODE = {0: O.FIND, 1: O.FIND_OVERLAPPING, 2: O.FIND_OVERLAPPING_NO_SUFFIX, 3: O.LEFTMOST_FIND}


def triples(m):
    return [(int(a), int(b), int(c)) for a, b, c in zip(m["start"], m["end"], m["value"])]


def _cases():
    for variant, iterator, coll, kind in GOLD["configs"]:
        if iterator not in MODE:
            continue
        for g in GOLD["collections"][coll]:
            for t in GOLD["groups"][g]:
                yield pytest.param(variant, iterator, kind, t, id="%s-%s-%s-%s" % (variant, iterator, kind, t["name"]))


@pytest.mark.parametrize("variant,iterator,kind,t", list(_cases()))
def test_golden_vectors(variant, iterator, kind, t):
    cw = variant == "charwise"
    wire = O.OraclePma.build(t["patterns"], charwise=cw, match_kind=O.KIND[kind]).serialize()
    hay = t["haystack"].encode()
    text = np.frombuffer(hay, dtype=np.uint8)
    for hot, kernel in ((0, 1), (3, 0), (0, 2), (0, 3), (256, 3), (1 << 16, 3), (0, 4)):
        rc, m, oo, need = E.scan(wire, cw, MODE[iterator], text, np.array([0, len(hay)], dtype=np.uint64), hot_n=hot,
                                 kernel=kernel)
        assert rc == 0
        assert triples(m) == [(s, e, v) for v, s, e in t["matches"]]


def rand_patterns(rng, n, alpha, maxlen, allow_empty=False):
    return [bytes(rng.integers(97, 97 + alpha, size=int(rng.integers(0 if allow_empty else 1, maxlen + 1))).tolist())
            for _ in range(n)]


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("cw", [False, True])
def test_random_batches(seed, kind, cw):
    rng = np.random.default_rng(100 * seed + 10 * kind + cw)
    alpha = int(rng.integers(2, 5))
    pats = rand_patterns(rng, int(rng.integers(1, 60)), alpha, 7, allow_empty=(seed == 0))
    if cw:  # mix in multi-byte symbols
        table = ["a", "b", "é", "あ", "𝄞"]
        pats = ["".join(table[b - 97] for b in p) for p in pats]
        hay_syms = [table[i] for i in range(alpha)] + ["z"]
    pma = O.OraclePma.build(pats, charwise=cw, match_kind=kind)
    wire = pma.serialize()
    n = 25
    hays = []
    for _ in range(n):
        L = int(rng.integers(0, 120))
        if cw:
            hays.append("".join(hay_syms[int(i)] for i in rng.integers(0, len(hay_syms), size=L)).encode())
        else:
            hays.append(bytes(rng.integers(97, 97 + alpha + 1, size=L).tolist()))
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(h) for h in hays])
    text = np.frombuffer(b"".join(hays), dtype=np.uint8)
    modes = [3] if kind else [0, 1, 2]
    for mode in modes:
        ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
        for hot, kernel in ((0, 1), (5, 0), (0, 2), (0, 3), (256, 3), (512, 3), (1 << 16, 3), (0, 4)):
            rc, m, oo, need = E.scan(wire, cw, mode, text, offs, hot_n=hot, kernel=kernel)
            assert rc == 0
            assert need == ref["total"]
            assert np.array_equal(np.diff(oo.astype(np.int64)), ref["counts"].astype(np.int64))
            assert m.tobytes() == ref["matches"].tobytes(), (pats, mode, hot)


def test_unaligned_offsets_and_long_chains():
    """Haystack starts at every alignment of the 16-byte text window; many outputs per position
    so that matches span several 20-match blocks."""
    pats = [b"a" * k for k in range(1, 30)] + [b"a"] * 5
    pma = O.OraclePma.build(pats)
    wire = pma.serialize()
    for shift in range(0, 33):
        text = np.frombuffer(b"x" * shift + b"a" * 70 + b"b" + b"a" * 40, dtype=np.uint8)
        offs = np.array([shift, shift + 50, shift + 50, shift + 111], dtype=np.uint64)
        ref = pma.scan_batch(O.FIND_OVERLAPPING, text, offs, want_matches=True)
        for kernel in (0, 1, 2, 3, 4):
            rc, m, oo, need = E.scan(wire, False, 1, text, offs, kernel=kernel, hot_n=256 if kernel == 3 else 0)
            assert rc == 0 and m.tobytes() == ref["matches"].tobytes()
            assert list(oo) == [0] + list(np.cumsum(ref["counts"]))


def test_overflow_protocol():
    pma = O.OraclePma.build([b"a", b"aa"])
    wire = pma.serialize()
    text = np.frombuffer(b"a" * 100, dtype=np.uint8)
    offs = np.array([0, 100], dtype=np.uint64)
    rc, m, oo, need = E.scan(wire, False, 1, text, offs, out_cap=10)
    assert rc == 6 and need == 199
    rc, m, oo, need = E.scan(wire, False, 1, text, offs, out_cap=199, pool_blocks=3)
    assert rc == 6 and need == 199  # pool exhausted: counting continues
    rc, m, oo, need = E.scan(wire, False, 1, text, offs, out_cap=199)
    assert rc == 0 and len(m) == 199


def test_mode_gating():
    wire = O.OraclePma.build([b"a"], match_kind=1).serialize()
    rc, *_ = E.scan(wire, False, 1, np.zeros(0, np.uint8), np.array([0, 0], dtype=np.uint64))
    assert rc == 5


@pytest.mark.parametrize("seed", range(5))
def test_segments_reproduce_the_sequential_scan(seed):
    """Intra-haystack segments with an (L-1)-byte warm-up: every segment
    length gives the sequential result, including empty patterns, empty haystacks, haystacks that
    end exactly on a segment boundary and patterns longer than a segment."""
    rng = np.random.default_rng(400 + seed)
    alpha = int(rng.integers(2, 5))
    pats = rand_patterns(rng, int(rng.integers(1, 60)), alpha, 9, allow_empty=(seed == 0))
    pma = O.OraclePma.build(pats)
    wire = pma.serialize()
    lens = list(rng.integers(0, 400, size=30)) + [0, 64, 128, 1, 63, 65]
    hays = [bytes(rng.integers(97, 97 + alpha + 1, size=int(L)).tolist()) for L in lens]
    offs = np.zeros(len(hays) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(h) for h in hays])
    text = np.frombuffer(b"".join(hays), dtype=np.uint8)
    for mode in (1, 2):
        ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
        for seg_len in (1, 3, 16, 64, 100, 1000):
            for kernel in (1, 2, 3, 4):
                for seg_from in (0, 11, len(hays) - 3):  # > 0: only the tail of the batch is cut
                    rc, m, oo, need = E.scan(wire, False, mode, text, offs, seg_len=seg_len, kernel=kernel, seg_from=seg_from,
                                             hot_n=256 if kernel == 3 else 0)
                    assert rc == 0 and need == ref["total"], (seg_len, mode)
                    assert m.tobytes() == ref["matches"].tobytes(), (seg_len, mode)
                    assert np.array_equal(np.diff(oo.astype(np.int64)), ref["counts"].astype(np.int64))


@pytest.mark.parametrize("seed", range(45))
def test_charwise_mixed_width_chars(seed):
    kind, pats, text, offs = mixed_width_case(seed)
    pma = O.OraclePma.build(pats, charwise=True, match_kind=kind)
    wire = pma.serialize()
    for mode in ([3] if kind else [0, 1, 2]):
        ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
        for kernel in (1, 0):
            rc, m, oo, need = E.scan(wire, True, mode, text, offs, kernel=kernel)
            assert rc == 0 and need == ref["total"]
            assert m.tobytes() == ref["matches"].tobytes(), (pats, mode, kernel)
            assert np.array_equal(np.diff(oo.astype(np.int64)), ref["counts"].astype(np.int64))


def test_charwise_leftmost_empty_pattern_never_stops_inside_a_char():
    """The reference advances self.pos by the length of the char that fell back to ROOT
    (src/charwise/iter.rs:345-346); with an empty pattern and chars of mixed widths that lands inside
    a char (undefined behaviour in the crate).  Oracle and kernels move on to the next boundary."""
    pats = ["ca", "", "𝄞𝄞c"]
    pma = O.OraclePma.build(pats, charwise=True, match_kind=2)
    hay = "bc𝄞".encode()
    text = np.frombuffer(hay, dtype=np.uint8)
    offs = np.array([0, len(hay)], dtype=np.uint64)
    ref = pma.scan_batch(O.LEFTMOST_FIND, text, offs, want_matches=True)
    assert triples(ref["matches"]) == [(0, 0, 1), (1, 1, 1), (6, 6, 1)]
    for kernel in (1, 0):
        rc, m, oo, need = E.scan(pma.serialize(), True, 3, text, offs, kernel=kernel)
        assert rc == 0 and triples(m) == triples(ref["matches"])


def _stream_case(seed):
    rng = np.random.default_rng(31000 + seed)
    alpha = int(rng.integers(1, 5))
    pats = rand_patterns(rng, int(rng.integers(1, 60)), alpha, int(rng.integers(1, 10)), allow_empty=(seed % 6 == 0))
    n_streams = 9
    streams = [bytes(rng.integers(97, 97 + alpha + (seed % 2), size=int(rng.integers(0, 500))).tolist()) for _ in range(n_streams)]
    return pats, streams, rng


@pytest.mark.parametrize("seed", range(24))
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("kernel", [3, 2, 4])
def test_stream_chunks_equal_the_stepper_over_the_whole_stream(seed, mode, kernel):
    """dach_dev_scan_stream (lane machine, emulated): streams cut into ragged chunks and scanned round
    by round with the state carried over report exactly what the crate's stepper reports over the
    whole stream (consume + matches(), tests/aho_corasick_crate_test.rs:422-520) -- minus the
    matches() of the initial state at position 0, which no consume() produced."""
    pats, streams, rng = _stream_case(seed)
    E.lib().emu_stream_config(kernel, 256 if seed % 2 else 0)
    pma = O.OraclePma.build(pats)
    wire = pma.serialize()
    orc_mode = O.FIND_STEPPER if mode == 0 else O.FIND_OVERLAPPING_STEPPER
    want = []
    for sbytes in streams:
        t = np.frombuffer(sbytes, dtype=np.uint8)
        ref = pma.scan_batch(orc_mode, t, np.array([0, len(sbytes)], dtype=np.uint64), want_matches=True)
        want.append([x for x in triples(ref["matches"]) if x[1] != 0])
    state = np.zeros(len(streams), dtype=np.uint32)
    pos = np.zeros(len(streams), dtype=np.uint32)
    got = [[] for _ in streams]
    while any(int(pos[i]) < len(s) for i, s in enumerate(streams)):
        chunks = []
        for i, s in enumerate(streams):
            k = int(rng.integers(0, 70))  # empty chunks included
            chunks.append(s[int(pos[i]): int(pos[i]) + k])
        offs = np.zeros(len(streams) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(c) for c in chunks])
        text = np.frombuffer(b"".join(chunks), dtype=np.uint8)
        rc, m, oo, need = E.scan_stream(wire, mode, text, offs, state, pos)
        if mode == 0 and any(len(p) == 0 for p in pats):
            assert rc == 1  # DACH_INVALID_ARGUMENT: find stepper with an empty pattern keeps the simple kernel
            return
        assert rc == 0
        tr = triples(m)
        for i in range(len(streams)):
            got[i] += tr[int(oo[i]): int(oo[i + 1])]
            pos[i] += len(chunks[i])
    assert got == want
    # the carried state is the reference's state id: one more round from a CPU-side walk agrees
    for i, s in enumerate(streams):
        assert int(state[i]) == pma.state_after(s, find_mode=(mode == 0))


@pytest.mark.parametrize("kind", [0, 1, 2])
@pytest.mark.parametrize("hot_slots", [0, 256, 4096, 65536])
def test_device_image_transition_function_is_the_crates(kind, hot_slots):
    """Exhaustive, scan-independent check of the compact bytewise image (dev_image.cpp): for EVERY state and
    byte, walking the records the way the lane machines do -- child signature, BASE ^ c with CHECK, the
    pre-resolved failure base, CF_FROOT / CF_F2ROOT / CF_F2DEAD -- lands where the crate's next_state
    (src/bytewise.rs:1063-1128) lands, for every size of the hot region (none, one block, some, all)."""
    rng = np.random.default_rng(4200 + kind)
    for pats in (sorted(set(rand_patterns(rng, 2500, 5, 10))), [b"a"], [b"ab", b"b", bytes(range(256))],
                 sorted(set(bytes(rng.integers(0, 256, size=int(rng.integers(1, 6))).tolist()) for _ in range(3000)))):
        pma = O.OraclePma.build(pats, match_kind=kind)
        bad, hs, used = E.check_image_transitions(pma.serialize(), hot_slots)
        assert bad == 0, (kind, hot_slots, len(pats))
        assert hs <= hot_slots and hs % 256 == 0
        if hot_slots and len(pats) > 1000:
            assert used > 2  # more than ROOT and DEAD live in the region


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_charwise_device_image_transition_function_is_the_crates(kind):
    """The same exhaustive check for the charwise compact image (CwMachine): every reachable state x every mapped
    code, plus output positions and the 16-bit child signature (src/charwise.rs:1008-1060)."""
    rng = np.random.default_rng(4300 + kind)
    cps = [0x61, 0x62, 0x63, 0xE6, 0x3042, 0x3044, 0x4E00, 0x1F600] + list(range(0x4E10, 0x4E60))
    for npat, alpha in ((3000, 8), (3000, len(cps)), (1, 3)):
        pats = sorted(set("".join(chr(cps[int(i)]) for i in rng.integers(0, alpha, size=int(rng.integers(1, 8))))
                          for _ in range(npat)))
        pma = O.OraclePma.build([p.encode("utf-8") for p in pats], charwise=True, match_kind=kind)
        assert E.check_image_transitions_charwise(pma.serialize()) == 0, (kind, npat, alpha)


def test_hand_made_automata_keep_the_crates_numbering_and_are_scanned_whole():
    """What the builders make is Aho-Corasick's automaton: it may be relaid out and long haystacks may be cut into
    segments.  A hand-made blob that passes the crate's validation need not be one -- a state only failure links
    lead to that shares a BASE with a trie state, a failure link to some other state: the image must still scan
    like the crate's loops (exhaustive check), keeps the crate's layout, and is not segmentable."""
    rng = np.random.default_rng(8)
    pats = sorted(set(rand_patterns(rng, 400, 4, 7)))
    pma = O.OraclePma.build(pats)
    wire = pma.serialize()
    assert E.image_segmentable(wire) == 1 and E.check_image_transitions(wire, 256)[:2] == (0, 256)
    _, a = damage_standard_wire(wire, [])
    n = len(a)
    inner = [s for s in range(2, n) if a[s, 0] != 0 and a[s, 1] != 0]          # trie states with children, depth > 1
    vacant = [s for s in range(2, n) if a[s, 0] == 0 and a[s, 1] == 0 and (a[s, 2] >> 8) == 0]
    leafish = [s for s in range(2, n) if a[s, 0] == 0 and a[s, 1] != 0]
    assert inner and leafish
    # (a) a failure link that is not the longest proper suffix: still a valid automaton for the crate
    w1, _ = damage_standard_wire(wire, [(inner[0], 1, 0)])
    bad, hs, _ = E.check_image_transitions(w1, 256)
    assert bad == 0 and E.image_segmentable(w1) == 0
    # (b) a childless state gets the BASE of a trie state and a trie state fails to it: reachable by failure links only
    w2, _ = damage_standard_wire(wire, [(leafish[0], 0, int(a[inner[0], 0])), (inner[-1], 1, leafish[0])])
    bad, hs, _ = E.check_image_transitions(w2, 256)
    assert bad == 0 and hs == 0 and E.image_segmentable(w2) == 0
    # (c) damage in slots no scan can reach changes nothing
    if vacant:
        w3, _ = damage_standard_wire(wire, [(vacant[0], 1, 5)])
        assert E.check_image_transitions(w3, 256)[0] == 0 and E.image_segmentable(w3) == 1


def test_hand_made_case_of_the_gpu_suite_on_the_emulated_lanes():
    """tests/test_gpu_parity.py::test_hand_made_automaton_with_long_haystacks_is_scanned_whole on the CPU: whole
    haystacks equal the oracle; forcing segments anyway is wrong, which is why the product does not."""
    wire, text, offs = hand_made_case()
    pma, _ = O.OraclePma.deserialize(wire)
    assert E.image_segmentable(wire) == 0
    differs = False
    for mode in (1, 2):
        ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
        assert ref["total"] > 100
        for kernel in (3, 1):
            rc, m, oo, need = E.scan(wire, False, mode, text, offs, kernel=kernel)
            assert rc == 0 and m.tobytes() == ref["matches"].tobytes()
        rc, m, oo, need = E.scan(wire, False, mode, text, offs, kernel=3, seg_len=256)
        differs |= rc != 0 or need != ref["total"] or m.tobytes() != ref["matches"].tobytes()
    assert differs


def test_device_image_transitions_with_the_default_region_on_a_large_automaton():
    """The default region (65536 slots) in front of an automaton several times that size: most states stay in
    the shifted part and many families leave holes behind there."""
    rng = np.random.default_rng(99)
    pats = sorted(set(bytes(rng.integers(0, 256, size=int(rng.integers(2, 7))).tolist()) for _ in range(60000)))
    for kind in (0, 1):
        pma = O.OraclePma.build(pats, match_kind=kind)
        bad, hs, used = E.check_image_transitions(pma.serialize(), 65536)
        assert bad == 0 and hs == 65536 and used > 30000


@pytest.mark.parametrize("kind", [0, 1])
def test_nul_heavy_case_of_the_gpu_suite_on_the_emulated_lanes(kind):
    """tests/test_gpu_parity.py::test_binary_text_full_of_nul_bytes_on_a_relaid_out_automaton, same data, default
    region, on the CPU (fewer haystacks: the emulation is slow)."""
    pats, text, offs = nul_heavy_case(kind, n_hay=150)
    pma = O.OraclePma.build(pats, match_kind=kind)
    wire = pma.serialize()
    assert E.check_image_transitions(wire, 65536)[:2] == (0, 65536)
    for mode in ([3] if kind else [0, 1, 2]):
        ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
        assert ref["total"] > 1000
        for kernel in (3, 1):
            rc, m, oo, need = E.scan(wire, False, mode, text, offs, kernel=kernel)
            assert rc == 0 and need == ref["total"]
            assert m.tobytes() == ref["matches"].tobytes(), (kind, mode, kernel)


@pytest.mark.parametrize("hot_slots", [256, 1024])
def test_nul_bytes_do_not_land_in_slots_a_moved_family_left_behind(hot_slots):
    """Regression: a state whose BASE equals a slot that a re-placed family vacated must not take label 0x00 as
    a child there.  Binary patterns and texts full of NUL bytes, a region much smaller than the automaton."""
    rng = np.random.default_rng(31337)
    pats = sorted(set(bytes(rng.choice([0, 0, 1, 2, 3, 255], size=int(rng.integers(1, 9))).tolist()) for _ in range(4000)))
    E.lib().emu_set_hot_slots(hot_slots)
    try:
        for kind in (0, 1):
            pma = O.OraclePma.build(pats, match_kind=kind)
            wire = pma.serialize()
            n = 16
            hays = [bytes(rng.choice([0, 0, 0, 1, 2, 3, 255], size=int(rng.integers(0, 900))).tolist()) for _ in range(n)]
            offs = np.zeros(n + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(h) for h in hays])
            text = np.frombuffer(b"".join(hays), dtype=np.uint8)
            for mode in ([3] if kind else [0, 1, 2]):
                ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
                for kernel in (1, 2, 3, 4):
                    rc, m, oo, need = E.scan(wire, False, mode, text, offs, kernel=kernel)
                    assert rc == 0 and need == ref["total"], (kind, mode, kernel)
                    assert m.tobytes() == ref["matches"].tobytes(), (kind, mode, kernel)
    finally:
        E.lib().emu_set_hot_slots(65536)


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("hot_slots", [0, 256, 1024, 65536])
def test_hot_first_relayout_at_dictionary_shape(seed, hot_slots):
    """The compact image moves the children of the hottest states into a region at the front and shifts
    every other slot up (dev_image.cpp); results must not depend on the region's size -- none at all,
    one or a few 256-slot blocks (most states stay in the shifted part), everything -- nor on how much
    of it a kernel serves from shared memory.  Standard and leftmost automata, stream state ids included."""
    rng = np.random.default_rng(7000 + seed)
    pats = sorted(set(rand_patterns(rng, 1500, 6, 9)))
    E.lib().emu_set_hot_slots(hot_slots)
    try:
        for kind in (0, 1):
            pma = O.OraclePma.build(pats, match_kind=kind)
            wire = pma.serialize()
            n = 12
            hays = [bytes(rng.integers(97, 104, size=int(rng.integers(0, 700))).tolist()) for _ in range(n)]
            offs = np.zeros(n + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(h) for h in hays])
            text = np.frombuffer(b"".join(hays), dtype=np.uint8)
            for mode in ([3] if kind else [0, 1, 2]):
                ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
                for hot, kernel in ((0, 1), (0, 2), (0, 3), (256, 3), (768, 3), (1 << 16, 3), (0, 4)):
                    rc, m, oo, need = E.scan(wire, False, mode, text, offs, hot_n=hot, kernel=kernel)
                    assert rc == 0 and need == ref["total"]
                    assert m.tobytes() == ref["matches"].tobytes(), (mode, hot, kernel)
            if kind == 0:  # state ids crossing the boundary are the crate's
                E.lib().emu_stream_config(3, 256)
                state = np.zeros(n, dtype=np.uint32)
                rc, m, oo, need = E.scan_stream(wire, 1, text, offs, state)
                assert rc == 0
                for i, h in enumerate(hays):
                    assert int(state[i]) == pma.state_after(h, find_mode=False)
    finally:
        E.lib().emu_set_hot_slots(65536)


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("mode", [0, 1])
def test_charwise_stream_chunks_equal_the_stepper_over_the_whole_stream(seed, mode):
    """The charwise steppers (src/charwise/iter.rs:403-534) as chunks of streams on CwMachine: streams of chars of
    mixed byte widths cut at char boundaries into ragged chunks, state carried over."""
    rng = np.random.default_rng(52000 + seed)
    table = ["a", "b", "é", "あ", "𝄞", "z"]
    alpha = int(rng.integers(2, 5))
    pats = sorted({"".join(table[int(i)] for i in rng.integers(0, alpha, size=int(rng.integers(1, 6)))) for _ in range(int(rng.integers(1, 40)))})
    streams = [[table[int(i)] for i in rng.integers(0, alpha + 1, size=int(rng.integers(0, 200)))] for _ in range(7)]
    pma = O.OraclePma.build(pats, charwise=True)
    wire = pma.serialize()
    orc_mode = O.FIND_STEPPER if mode == 0 else O.FIND_OVERLAPPING_STEPPER
    want = []
    for chars in streams:
        sb = "".join(chars).encode()
        ref = pma.scan_batch(orc_mode, np.frombuffer(sb, dtype=np.uint8), np.array([0, len(sb)], dtype=np.uint64), want_matches=True)
        want.append([x for x in triples(ref["matches"]) if x[1] != 0])
    E.lib().emu_stream_charwise(1)
    try:
        state = np.zeros(len(streams), dtype=np.uint32)
        pos = np.zeros(len(streams), dtype=np.uint32)
        cur = [0] * len(streams)
        got = [[] for _ in streams]
        while any(cur[i] < len(s) for i, s in enumerate(streams)):
            chunks = []
            for i, s in enumerate(streams):
                k = int(rng.integers(0, 25))
                chunks.append("".join(s[cur[i]: cur[i] + k]).encode())
                cur[i] += k
            offs = np.zeros(len(streams) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(c) for c in chunks])
            text = np.frombuffer(b"".join(chunks), dtype=np.uint8)
            rc, m, oo, need = E.scan_stream(wire, mode, text, offs, state, pos)
            assert rc == 0
            tr = triples(m)
            for i in range(len(streams)):
                got[i] += tr[int(oo[i]): int(oo[i + 1])]
                pos[i] += len(chunks[i])
        assert got == want
    finally:
        E.lib().emu_stream_charwise(0)
