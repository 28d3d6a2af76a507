"""tools/proto/depsim.py -- CPU design study for the round-2 tile decoder.

For LZ4 blocks of several workloads it parses the exact sequence list and reports
  * sequence statistics (count, literal / match length distributions, offsets),
  * for a step size of K sequences (one sequence per thread, steps processed in order with a
    CTA barrier in between): the share of "near" matches -- source range reaching into the step's
    own output -- and how many exact-readiness sub-rounds those need,
  * the whole-block dependency level distribution (level-synchronous alternative).

    python tools/proto/depsim.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np

import oracle
from tests import inputs


def sequences(c):
    """[(tokenPos, litLen, matchLen, offset, outPos)] ; last one has matchLen 0"""
    c = bytes(c); n = len(c); p = 0; op = 0; out = []
    while p < n:
        tp = p
        t = c[p]; p += 1
        l = t >> 4
        if l == 15:
            while True:
                s = c[p]; p += 1; l += s
                if s != 255: break
        p += l
        if p >= n:
            out.append((tp, l, 0, 0, op)); op += l; break
        off = c[p] | (c[p + 1] << 8); p += 2
        m = t & 15
        if m == 15:
            while True:
                s = c[p]; p += 1; m += s
                if s != 255: break
        m += 4
        out.append((tp, l, m, off, op)); op += l + m
    return out, op


def study(name, blocks):
    allseq = [sequences(c) for c in blocks]
    ns = np.array([len(s) for s, _ in allseq])
    lit = np.array([q[1] for s, _ in allseq for q in s])
    ml = np.array([q[2] for s, _ in allseq for q in s if q[2]])
    off = np.array([q[3] for s, _ in allseq for q in s if q[2]])
    print(f"== {name}: {len(blocks)} blocks, seq/block {ns.mean():.0f} (max {ns.max()}), "
          f"lit mean {lit.mean():.2f} p50 {np.percentile(lit,50):.0f} p90 {np.percentile(lit,90):.0f} p99 {np.percentile(lit,99):.0f} max {lit.max()}; "
          f"match mean {ml.mean():.2f} p50 {np.percentile(ml,50):.0f} p90 {np.percentile(ml,90):.0f} p99 {np.percentile(ml,99):.0f} max {ml.max()}; "
          f"off<8 {np.mean(off<8)*100:.2f}% off<len {np.mean(off<ml)*100:.2f}%  lit>32 {np.mean(lit>32)*100:.2f}% ml>32 {np.mean(ml>32)*100:.2f}% ml>64 {np.mean(ml>64)*100:.2f}%")
    # bytes in long runs
    print(f"   bytes in lit runs >32: {lit[lit>32].sum()/max(lit.sum(),1)*100:.1f}%  bytes in matches >32: {ml[ml>32].sum()/max(ml.sum(),1)*100:.1f}%  >64: {ml[ml>64].sum()/max(ml.sum(),1)*100:.1f}%")
    # batch-of-32 divergence: sum over batches of max len vs mean
    def batch_max(x_per_block):
        tot_max = 0; tot = 0; nb = 0
        for x in x_per_block:
            for i in range(0, len(x), 32):
                b = x[i:i + 32]
                tot_max += b.max(); tot += b.sum(); nb += 1
        return tot_max / nb, tot / nb / 32
    lits = [np.minimum(np.array([q[1] for q in s]), 32) for s, _ in allseq]
    mls = [np.minimum(np.array([q[2] for q in s]), 32) for s, _ in allseq]
    print("   per 32-seq batch (runs capped at 32): lit max %.1f mean %.1f | match max %.1f mean %.1f" % (*batch_max(lits), *batch_max(mls)))
    for K in (256, 512, 1024):
        near_tot = 0; m_tot = 0; rounds_tot = 0; steps_tot = 0; worst = 0; r_hist = {}
        for s, total in allseq:
            ready = np.zeros(total + 1, dtype=bool)
            for st in range(0, len(s), K):
                step = s[st:st + K]
                S0 = step[0][4]
                # literals + far matches
                for (tp, l, m, o, op) in step:
                    ready[op:op + l] = True
                pend = []
                for (tp, l, m, o, op) in step:
                    if not m: continue
                    m_tot += 1
                    d = op + l; a = d - o
                    srcend = min(a + m, d)
                    if srcend <= S0:
                        ready[d:d + m] = True      # far: in phase A
                    else:
                        pend.append((d, m, a, srcend))
                near_tot += len(pend)
                # careful: far matches marked ready above BEFORE near check == after the barrier
                r = 0
                while pend:
                    r += 1
                    nxt = []; done = []
                    for (d, m, a, srcend) in pend:
                        if a >= 0 and ready[max(a,0):srcend].all():
                            done.append((d, m))
                        else:
                            nxt.append((d, m, a, srcend))
                    if not done:   # offset beyond start etc.
                        break
                    for d, m in done: ready[d:d + m] = True
                    pend = nxt
                rounds_tot += r; steps_tot += 1; worst = max(worst, r)
                r_hist[r] = r_hist.get(r, 0) + 1
        print(f"   K={K}: near {near_tot/m_tot*100:.1f}% of matches, sub-rounds/step avg {rounds_tot/steps_tot:.2f} worst {worst}, steps/block {steps_tot/len(blocks):.1f}")
    # whole-block levels
    lv_hist = {}
    maxlv = []
    for s, total in allseq:
        level = np.zeros(total + 1, dtype=np.int32)
        mx = 0
        for (tp, l, m, o, op) in s:
            if not m: continue
            d = op + l; a = d - o
            if a < 0: continue
            srcend = min(a + m, d)
            lv = int(level[a:srcend].max()) + 1 if srcend > a else 1
            level[d:d + m] = lv
            lv_hist[lv] = lv_hist.get(lv, 0) + 1
            mx = max(mx, lv)
        maxlv.append(mx)
    tot = sum(lv_hist.values())
    cum = 0; parts = []
    for k in sorted(lv_hist):
        cum += lv_hist[k]
        if k <= 8 or k % 8 == 0: parts.append(f"{k}:{cum/tot*100:.0f}%")
    print(f"   levels: max/block avg {np.mean(maxlv):.1f} worst {max(maxlv)}; cumulative {' '.join(parts)}")


def main():
    ref = oracle.best()
    port = oracle.Port()
    bs = 65536
    raw = port.datagen(48 * bs, 0.63, 0.0, 1234)
    study("datagen0.63", [np.frombuffer(ref.encode(raw[i * bs:(i + 1) * bs])[1], dtype=np.uint8) for i in range(16, 28)])
    from k4os.compression.lz4_b200.batch import synth_host
    raw = synth_host(8, bs, 525)
    study("synth525", [np.frombuffer(ref.encode(raw[i * bs:(i + 1) * bs])[1], dtype=np.uint8) for i in range(8)])
    for kind in ("text2", "lowent", "runs", "lorem"):
        study(kind, [np.frombuffer(ref.encode(inputs.gen(kind, bs, seed))[1], dtype=np.uint8) for seed in range(3)])


if __name__ == "__main__":
    main()
