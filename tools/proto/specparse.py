"""tools/proto/specparse.py -- CPU prototype of the next-round decoder idea:
parse ONE block with many lanes instead of one thread per block.

Lane t owns compressed-stream segment [t*S, (t+1)*S).  It starts walking the token chain W bytes
BEFORE its segment (speculatively: it does not know where a token starts), records
    X_t = first position >= t*S on its walk      (candidate entry into its segment)
    E_t = first position >= (t+1)*S on its walk  (its exit = candidate entry of a later segment)
LZ4 token chains are confluent: two walks that ever land on the same position stay together, and a
walk started at an arbitrary byte usually falls onto the true chain within a few dozen bytes.  Lane 0
starts at the true position 0.  Validation: follow link(t) = segment holding E_t from lane 0; a
visited lane u reached from lane t is consistent iff X_u == E_t.  Inconsistent lanes re-walk their
segment from the true entry; repeat until all visited lanes are consistent (round count reported).
This script measures, on the bench workload and on awkward inputs, how often repair is needed and how
many rounds it takes, for a few (S, W) choices -- all against the exact token chain of the oracle.

    python tools/proto/specparse.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import oracle
from tests import inputs
from k4os.compression.lz4_b200.batch import synth_host


def next_pos(c, p, n):
    """position of the token after the one at p (n = len(c)); returns n when the stream ends/breaks"""
    t = c[p]; q = p + 1; l = t >> 4
    if l == 15:
        while q < n:
            s = c[q]; q += 1; l += s
            if s != 255:
                break
    q += l
    if q + 2 > n:
        return n
    q += 2
    if (t & 15) == 15:
        while q < n:
            s = c[q]; q += 1
            if s != 255:
                break
    return min(q, n)


def true_chain(c):
    n = len(c); p = 0; out = []
    while p < n:
        out.append(p); p = next_pos(c, p, n)
    return out


def simulate(c, S, W):
    n = len(c)
    nl = (n + S - 1) // S
    X = [None] * nl; E = [None] * nl
    hops = 0

    def walk(start, seg):
        nonlocal hops
        p = start; x = None
        lo, hi = seg * S, min((seg + 1) * S, n)
        while p < hi:
            if x is None and p >= lo:
                x = p
            p = next_pos(c, p, n); hops += 1
        if x is None and p >= lo and lo < n:
            x = None            # the walk jumped over the whole segment
        return x, p

    for t in range(nl):
        X[t], E[t] = walk(max(0, t * S - W) if t else 0, t)
    X[0] = 0
    spec_hops = hops
    # validation / repair rounds
    rounds = 0; rewalks = 0
    while True:
        rounds += 1
        bad = []
        t = 0
        while True:
            e = E[t]
            if e >= n:
                break
            u = e // S
            if X[u] != e:
                bad.append((u, e))
                break               # everything downstream is unknown until u is repaired
            t = u
        if not bad:
            break
        for u, e in bad:
            X[u] = e
            _, E[u] = walk(e, u); rewalks += 1
    # check against the exact chain
    tc = true_chain(c)
    got = []
    t = 0
    while True:
        p = X[t]
        hi = min((t + 1) * S, n)
        while p < hi:
            got.append(p); p = next_pos(c, p, n)
        if p >= n:
            break
        t = p // S
    assert got == tc, "prototype disagrees with the true chain"
    return dict(lanes=nl, rounds=rounds, rewalks=rewalks, spec_hops=spec_hops, seqs=len(tc))


def main():
    port = oracle.Port()
    cases = []
    raw = synth_host(16, 65536, 525)
    for i in range(16):
        cases.append(("synth525", port.encode(raw[i * 65536:(i + 1) * 65536])[1]))
    for kind in ("text2", "lowent", "runs", "lorem", "random"):
        for seed in range(3):
            cases.append((kind, port.encode(inputs.gen(kind, 65536, seed))[1]))
    for S, W in ((128, 128), (128, 384), (256, 256), (256, 512), (1024, 0), (1024, 512)):
        agg = {}
        for name, c in cases:
            r = simulate(c, S, W)
            a = agg.setdefault(name, dict(blocks=0, rounds=0, rewalks=0, lanes=0, hops=0, seqs=0, worst=0))
            a["blocks"] += 1; a["rounds"] += r["rounds"]; a["rewalks"] += r["rewalks"]; a["lanes"] += r["lanes"]
            a["hops"] += r["spec_hops"]; a["seqs"] += r["seqs"]; a["worst"] = max(a["worst"], r["rounds"])
        print(f"S={S} W={W}")
        for name, a in agg.items():
            print(f"   {name:9s} blocks {a['blocks']:2d}  lanes/block {a['lanes']/a['blocks']:6.1f}  "
                  f"repair rounds avg {a['rounds']/a['blocks']:5.2f} worst {a['worst']:3d}  "
                  f"re-walked lanes/block {a['rewalks']/a['blocks']:5.2f}  "
                  f"speculative hops / sequence {a['hops']/max(a['seqs'],1):4.2f}")


if __name__ == "__main__":
    main()
