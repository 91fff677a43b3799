"""Tiny GPU probe for compute-sanitizer: a few automata x modes on small inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import daachorse_b200 as D

def run(pats, hay, modes, kind=0, cw=False):
    B = D.CharwiseDoubleArrayAhoCorasickBuilder if cw else D.DoubleArrayAhoCorasickBuilder
    p = B.new().match_kind(kind).build(pats)
    p.device_handle(0)
    for m in modes:
        r = p.scan_batch_host(m, np.frombuffer(hay, dtype=np.uint8), np.array([0, len(hay)], dtype=np.uint64), device=0)
        print(pats, hay, m, r.triples(0), flush=True)

run([], b"", [0, 1, 2])
run([], b"abc", [0, 1, 2])
run(["a"], b"bababbbba", [0, 1, 2])
run([""], b"a", [0, 1, 2])
run(["abcd", "bcd", "cd", "b"], b"abcd", [0, 1, 2])
run(["ab", "a"], b"xayabbbz", [3], kind=1)
