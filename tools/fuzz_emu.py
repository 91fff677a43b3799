"""Randomised differential run of the emulated lane logic (tests/emu: scan_lane.cuh + dev_image.cpp compiled for
the CPU) against the oracle: binary and text alphabets, every match kind and iterator, every lane-machine kernel,
hot regions smaller than the automaton, segmented scans.  With --mutate the serialized automaton is damaged first
(random BASE / CHECK / failure fields, the way a hand-made blob handed to deserialize might look): whatever still
passes validation must scan like the crate's loops do.  Not part of the test suite; run it for as long as you like:

    python tools/fuzz_emu.py --seconds 600 --seed 1 [--mutate]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import emu_api as E  # noqa: E402
import oracle_api as O  # noqa: E402

ORC_MODE = {0: O.FIND, 1: O.FIND_OVERLAPPING, 2: O.FIND_OVERLAPPING_NO_SUFFIX, 3: O.LEFTMOST_FIND}


def one_case(rng):
    cw = bool(rng.integers(0, 4) == 0)
    kind = int(rng.integers(0, 3))
    style = int(rng.integers(0, 4))
    if cw:
        cps = [0x61, 0x62, 0x63, 0x64, 0xE6, 0x3042, 0x3044, 0x4E00, 0x1F600, 0x7F, 0x80, 0x7FF, 0x800, 0xFFFF, 0x10000]
        alpha = int(rng.integers(2, len(cps) + 1))
        sym = lambda k: "".join(chr(cps[int(i)]) for i in rng.integers(0, alpha, size=k)).encode("utf-8")  # noqa: E731
    else:
        if style == 0:
            pool = np.array([0, 0, 1, 2, 255], dtype=np.uint8)
        elif style == 1:
            pool = np.arange(256, dtype=np.uint8)
        else:
            pool = np.arange(97, 97 + int(rng.integers(2, 9)), dtype=np.uint8)
        sym = lambda k: bytes(rng.choice(pool, size=k).tolist())  # noqa: E731
    npat = int(rng.choice([1, 5, 60, 800, 6000]))
    maxlen = int(rng.integers(1, 14))
    allow_empty = kind != 0 and rng.integers(0, 3) == 0
    pats = sorted(set(sym(int(rng.integers(0 if allow_empty else 1, maxlen + 1))) for _ in range(npat)))
    pats = [p for p in pats if p or allow_empty] or [sym(1)]
    n = int(rng.integers(1, 40))
    hays = [sym(int(rng.integers(0, int(rng.choice([4, 60, 700, 5000]))))) for _ in range(n)]
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(h) for h in hays])
    text = np.frombuffer(b"".join(hays), dtype=np.uint8) if offs[-1] else np.zeros(0, dtype=np.uint8)
    hot_slots = int(rng.choice([0, 256, 512, 4096, 65536]))
    return cw, kind, pats, text, offs, hot_slots


def mutate(rng, wire, cw, kind):
    """Random field damage in the crate's wire format (src/bytewise.rs:801-820, src/charwise.rs:831-848): BASE,
    CHECK and failure fields.  Output lists are left alone: a length larger than the depth of its state makes the
    crate compute `end - length` below zero (a panic in debug builds), there is no behaviour to match."""
    w = bytearray(wire)
    u32 = lambda off: int(np.frombuffer(w, dtype="<u4", count=1, offset=off)[0])  # noqa: E731
    if cw:
        n = u32(0)
        a = np.frombuffer(w, dtype="<u4", count=n * 4, offset=4).reshape(n, 4)  # base, check, fail, output_pos
        cols = {"base": a[:, 0], "check": a[:, 1], "fail": a[:, 2]}
    elif kind == 0:
        n = u32(0)
        a = np.frombuffer(w, dtype="<u4", count=n * 3, offset=4).reshape(n, 3)  # base, fail, output_pos << 8 | check
        cols = {"base": a[:, 0], "fail": a[:, 1], "opos_ch": a[:, 2]}
    else:
        assert u32(0) == 0
        n = u32(4)
        a = np.frombuffer(w, dtype="<u4", count=n * 2, offset=8).reshape(n, 2)  # base, output_pos << 8 | check
        assert u32(8 + 8 * n) == n
        f = np.frombuffer(w, dtype="<u4", count=n, offset=12 + 8 * n)
        cols = {"base": a[:, 0], "opos_ch": a[:, 1], "fail": f}
    for _ in range(int(rng.integers(1, 6))):
        s, what = int(rng.integers(0, n)), int(rng.integers(0, 4))
        if what == 0:
            cols["base"][s] = int(rng.integers(0, n))
        elif what == 1:
            cols["fail"][s] = int(rng.integers(0, n))
        elif what == 2 and cw:
            cols["check"][s] = int(rng.integers(0, n))
        elif what == 2:
            cols["opos_ch"][s] = (int(cols["opos_ch"][s]) & ~0xFF) | int(rng.integers(0, 256))
        else:
            cols["base"][s] = 0
    return bytes(w)


def stream_rounds(rng, pma, wire, cw, text, offs, ctx, mutated):
    """dach_dev_scan_stream: the haystacks as streams, fed in ragged chunks (cut at char boundaries for charwise)
    with the state carried over, against the crate's steppers over the whole stream."""
    n = len(offs) - 1
    streams = [bytes(text[int(offs[i]): int(offs[i + 1])]) for i in range(n)]
    rounds = 0
    E.lib().emu_stream_charwise(1 if cw else 0)
    try:
        for mode in (0, 1):
            if not cw:
                E.lib().emu_stream_config(int(rng.choice([2, 3, 4])), int(rng.choice([0, 256])))
            orc_mode = O.FIND_STEPPER if mode == 0 else O.FIND_OVERLAPPING_STEPPER
            want = []
            for sb in streams:
                ref = pma.scan_batch(orc_mode, np.frombuffer(sb, dtype=np.uint8), np.array([0, len(sb)], dtype=np.uint64),
                                     want_matches=True)
                mm = ref["matches"]
                want.append([(int(x["start"]), int(x["end"]), int(x["value"])) for x in mm if int(x["end"]) != 0])
            state = np.zeros(n, dtype=np.uint32)
            pos = np.zeros(n, dtype=np.uint32)
            got = [[] for _ in streams]
            step = int(rng.choice([3, 40, 700]))
            while any(int(pos[i]) < len(sb) for i, sb in enumerate(streams)):
                chunks = []
                for i, sb in enumerate(streams):
                    e = min(len(sb), int(pos[i]) + int(rng.integers(0, step)))
                    while cw and e < len(sb) and (sb[e] & 0xC0) == 0x80:
                        e += 1
                    chunks.append(sb[int(pos[i]): e])
                co = np.zeros(n + 1, dtype=np.uint64)
                co[1:] = np.cumsum([len(c) for c in chunks])
                ct = np.frombuffer(b"".join(chunks), dtype=np.uint8) if co[-1] else np.zeros(0, dtype=np.uint8)
                rc, m, oo, need = E.scan_stream(wire, mode, ct, co, state, pos)
                if rc == 1 and mutated and rounds == 0:
                    return 0  # BASE(ROOT) damaged to 0: stream chunks are refused (documented in the header)
                assert rc == 0, ("stream rc", ctx, mode, rc)
                for i in range(n):
                    got[i] += [(int(x["start"]), int(x["end"]), int(x["value"])) for x in m[int(oo[i]): int(oo[i + 1])]]
                    pos[i] += len(chunks[i])
                rounds += 1
            assert got == want, ("stream", ctx, mode)
            for i, sb in enumerate(streams[:3]):  # the carried state is the crate's state id (pure-Python walk: a few only)
                if not cw:
                    assert int(state[i]) == pma.state_after(sb, find_mode=(mode == 0)), ("stream state", ctx, mode, i)
    finally:
        E.lib().emu_stream_charwise(0)
        E.lib().emu_stream_config(3, 0)
    return rounds


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--mutate", action="store_true")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    t0, cases, scans, rejected = time.time(), 0, 0, 0
    while time.time() - t0 < a.seconds:
        cw, kind, pats, text, offs, hot_slots = one_case(rng)
        try:
            pma = O.OraclePma.build(pats, charwise=cw, match_kind=kind)
        except O.OracleError:
            continue
        wire = pma.serialize()
        if a.mutate:
            wire = mutate(rng, wire, cw, kind)
        E.lib().emu_set_hot_slots(hot_slots)
        if not cw:
            bad = E.check_image_transitions(wire, hot_slots)[0]
        else:
            bad = E.check_image_transitions_charwise(wire)
        if a.mutate and bad < 0:  # refused (validation, or failure links that never reach the root)
            rejected += 1
            continue
        assert bad == 0, ("image", a.seed, cases, cw, kind, hot_slots)
        if a.mutate:
            try:
                pma, _ = O.OraclePma.deserialize(wire, charwise=cw)
            except O.OracleError:
                rejected += 1
                continue
        for mode in ([3] if kind else [0, 1, 2]):
            ref = pma.scan_batch(ORC_MODE[mode], text, offs, want_matches=True)
            kernels = (1,) if cw else ((1, 2) if kind else (1, 2, 3, 4))
            for kernel in kernels:
                segs = [(0, 0)]
                if not cw and mode in (1, 2) and kernel != 4 and rng.integers(0, 2) and E.image_segmentable(wire) == 1:
                    segs.append((int(rng.choice([64, 256, 1024])), int(rng.integers(0, len(offs)))))
                for seg_len, seg_from in segs:
                    hot_n = int(rng.choice([0, 256])) if kernel == 3 else 0
                    rc, m, oo, need = E.scan(wire, cw, mode, text, offs, kernel=kernel, seg_len=seg_len, seg_from=seg_from, hot_n=hot_n)
                    ctx = (a.seed, cases, cw, kind, mode, kernel, seg_len, seg_from, hot_slots, hot_n)
                    assert rc == 0 and need == ref["total"], ctx
                    assert m.tobytes() == ref["matches"].tobytes(), ctx
                    assert np.array_equal(np.diff(oo.astype(np.int64)), ref["counts"].astype(np.int64)), ctx
                    scans += 1
                if ref["total"] > 1 and rng.integers(0, 4) == 0:  # the overflow protocol: the count is still exact
                    cap = int(rng.integers(0, ref["total"]))
                    rc, m, oo, need = E.scan(wire, cw, mode, text, offs, kernel=kernel, out_cap=cap)
                    assert rc == 6 and need == ref["total"], ("cap", a.seed, cases, cw, kind, mode, kernel, cap)
                    rc, m, oo, need = E.scan(wire, cw, mode, text, offs, kernel=kernel, out_cap=ref["total"],
                                             pool_blocks=int(rng.integers(1, 4)))
                    assert (rc == 6 and need == ref["total"]) or (rc == 0 and m.tobytes() == ref["matches"].tobytes()), \
                        ("pool", a.seed, cases, cw, kind, mode, kernel)
                    scans += 2
        if kind == 0 and not any(len(p) == 0 for p in pats) and rng.integers(0, 3) == 0:
            scans += stream_rounds(rng, pma, wire, cw, text, offs, (a.seed, cases, cw, hot_slots), a.mutate)
        cases += 1
    E.lib().emu_set_hot_slots(65536)
    print(f"fuzz_emu: seed {a.seed}{' mutated' if a.mutate else ''}: {cases} cases ({rejected} refused), {scans} scans, all equal to the oracle")


if __name__ == "__main__":
    main()
