"""tools/dbench.py -- kernel-only micro-benchmark used while tuning (not the headline bench).
Times k4lz4_decode_batch / k4lz4_encode_batch on device-resident blocks with CUDA events and
checks the result (decode: against the raw input; encode: round trip).
    python tools/dbench.py [--blocks N] [--data datagen|synth] [--mp 630] [--reps 5] [--what decode|encode|both]
--data datagen: the reference's generator (oracle/: RDG_genBuffer, matchProba = mp/1000, seed 1234 + chunk)
--data synth  : the library's own device generator
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from k4os.compression.lz4_b200 import batch as B, _native as N

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=32768)
ap.add_argument("--bs", type=int, default=65536)
ap.add_argument("--mp", type=int, default=630)
ap.add_argument("--data", default="datagen")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--what", default="decode")
ap.add_argument("--check", type=int, default=0, help="encode: compare every CHECK-th block with the oracle engine (bit-exact)")
ap.add_argument("--lib", default=None, help="alternative build of libk4lz4.so (e.g. a -DK4_DT_PROFILE build under scratch/)")
a = ap.parse_args()
if a.lib:
    N.SO_PATH = os.path.abspath(a.lib)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
nb, bs = a.blocks, a.bs
bound = bs + bs // 255 + 16
if a.data == "synth":
    raw = torch.empty(nb * bs, dtype=torch.uint8, device=dev)
    B.synth_device(raw.data_ptr(), nb, bs, a.mp, 1234, 0, st)
else:
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    eng = oracle.best()
    host = np.empty(nb * bs, dtype=np.uint8)
    chunk = 1024 * bs
    nch = (nb * bs + chunk - 1) // chunk
    def gen(c):
        lo = c * chunk; hi = min(lo + chunk, nb * bs)
        eng.datagen(hi - lo, a.mp / 1000.0, 0.0, 1234 + c, out=host[lo:hi])
    with ThreadPoolExecutor(max_workers=min(64, os.cpu_count() or 1)) as ex:
        list(ex.map(gen, range(nch)))
    raw = torch.from_numpy(host).to(dev)
idx = torch.arange(nb, dtype=torch.int64, device=dev)
roff, coff = idx * bs, idx * bound
rlen = torch.full((nb,), bs, dtype=torch.int32, device=dev)
ccap = torch.full((nb,), bound, dtype=torch.int32, device=dev)
slots = torch.empty(nb * bound, dtype=torch.uint8, device=dev)
clen = torch.zeros(nb, dtype=torch.int32, device=dev)
def enc():
    B.encode_batch_device(raw.data_ptr(), roff.data_ptr(), rlen.data_ptr(), slots.data_ptr(), coff.data_ptr(),
                          ccap.data_ptr(), clen.data_ptr(), nb, 0, st)
def timeit(fn, reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))
enc(); torch.cuda.synchronize()
ratio = float(clen.sum()) / (nb * bs)
if a.what in ("encode", "both"):
    ms, med = timeit(enc, a.reps)
    if a.check:
        import oracle
        eng = oracle.best()
        hraw = raw.cpu().numpy(); hs = slots.cpu().numpy(); hl = clen.cpu().numpy()
        nbad = 0
        for i in range(0, nb, a.check):
            r, ref = eng.encode(hraw[i * bs:(i + 1) * bs])
            if r != int(hl[i]) or hs[i * bound:i * bound + r].tobytes() != ref: nbad += 1
        print(f"   encode check: {len(range(0, nb, a.check))} blocks vs oracle, {nbad} differ", flush=True)
    print(f"encode[{a.data}{a.mp}]: {ms:.3f} ms (median {med:.3f})  {nb*bs/ms/1e6:.1f} GB/s  ratio {ratio:.3f}", flush=True)
if a.what in ("decode", "both"):
    poff = torch.cumsum(clen.to(torch.int64), 0) - clen.to(torch.int64)
    packed = torch.empty(int(clen.sum()) + 64, dtype=torch.uint8, device=dev)
    B.copy_blocks_device(slots.data_ptr(), coff.data_ptr(), packed.data_ptr(), poff.data_ptr(), clen.data_ptr(), nb, st)
    out = torch.zeros(nb * bs, dtype=torch.uint8, device=dev)
    olen = torch.zeros(nb, dtype=torch.int32, device=dev)
    def dec():
        B.decode_batch_device(packed.data_ptr(), poff.data_ptr(), clen.data_ptr(), out.data_ptr(), roff.data_ptr(),
                              rlen.data_ptr(), olen.data_ptr(), nb, st)
    B.decode_stats(0, reset=True)
    dec(); torch.cuda.synchronize()
    stats = B.decode_stats(0, reset=True)
    ok = bool(torch.equal(out, raw)) and bool((olen == bs).all())
    if not ok:
        bad = (olen != bs).nonzero().flatten()[:8].tolist()
        o2 = out.view(nb, bs); r2 = raw.view(nb, bs)
        neq = (o2 != r2).any(dim=1).nonzero().flatten()
        first = neq[:8].tolist()
        detail = ""
        if len(first):
            blk = first[0]
            pos = (o2[blk] != r2[blk]).nonzero().flatten()
            detail = f" first bad block {blk}: {len(pos)} bytes differ, first at {int(pos[0])}, last at {int(pos[-1])}"
        print(f"MISMATCH: bad lengths at {bad} ({int((olen != bs).sum())} blocks), differing blocks {len(neq)} {first}{detail}", flush=True)
    prof = None
    if hasattr(N.lib(), "k4lz4_debug_prof"):
        import ctypes as C
        v = (C.c_uint64 * 32)()
        fn = N.lib().k4lz4_debug_prof
        fn.argtypes = [C.c_void_p, C.c_int32]
        fn(C.addressof(v), 1)
        dec(); torch.cuda.synchronize()
        fn(C.addressof(v), 1)
        names = ["load", "pass1", "validate", "scan", "pass2", "hdr", "lit", "far", "bar1", "nearlist", "nearwork", "nearbar", "endbar", "store", "table"]
        nblk = max(int(v[20]), 1)
        prof = {nm: round(int(v[i]) / nblk) for i, nm in enumerate(names)}
        prof["total"] = sum(prof.values())
        prof.update(valrounds=round(int(v[16]) / nblk, 2), steps=round(int(v[17]) / nblk, 2), subrounds=round(int(v[18]) / nblk, 2), inner=round(int(v[19]) / nblk, 2))
    ms, med = timeit(dec, a.reps)
    if prof: print("   cycles/block (thread 0):", prof, flush=True)
    algo = (int(clen.sum()) + nb * bs)
    print(f"decode[{a.data}{a.mp}]: {ms:.3f} ms (median {med:.3f})  {nb*bs/ms/1e6:.1f} GB/s out  {algo/ms/1e6:.1f} GB/s algorithmic  "
          f"ok={ok} ratio {ratio:.3f} stats {stats}", flush=True)
