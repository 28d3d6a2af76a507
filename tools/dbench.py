"""tools/dbench.py -- kernel-only micro-benchmark used while tuning (not the headline bench).
Times k4lz4_decode_batch / k4lz4_encode_batch on device-resident synthetic blocks with CUDA events.
    python tools/dbench.py [--blocks N] [--mp 525] [--reps 3] [--what decode|encode|both]
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from k4os.compression.lz4_b200 import batch as B, _native as N

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, default=32768)
ap.add_argument("--bs", type=int, default=65536)
ap.add_argument("--mp", type=int, default=525)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--what", default="decode")
a = ap.parse_args()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
nb, bs = a.blocks, a.bs
bound = bs + bs // 255 + 16
raw = torch.empty(nb * bs, dtype=torch.uint8, device=dev)
B.synth_device(raw.data_ptr(), nb, bs, a.mp, 1234, 0, st)
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
    return min(ts)
enc(); torch.cuda.synchronize()
ratio = float(clen.sum()) / (nb * bs)
if a.what in ("encode", "both"):
    ms = timeit(enc, a.reps)
    print(f"encode: {ms:.3f} ms  {nb*bs/ms/1e6:.1f} GB/s  ratio {ratio:.3f}")
if a.what in ("decode", "both"):
    poff = torch.cumsum(clen.to(torch.int64), 0) - clen.to(torch.int64)
    packed = torch.empty(int(clen.sum()) + 64, dtype=torch.uint8, device=dev)
    B.copy_blocks_device(slots.data_ptr(), coff.data_ptr(), packed.data_ptr(), poff.data_ptr(), clen.data_ptr(), nb, st)
    out = torch.zeros(nb * bs, dtype=torch.uint8, device=dev)
    olen = torch.zeros(nb, dtype=torch.int32, device=dev)
    def dec():
        B.decode_batch_device(packed.data_ptr(), poff.data_ptr(), clen.data_ptr(), out.data_ptr(), roff.data_ptr(),
                              rlen.data_ptr(), olen.data_ptr(), nb, st)
    ms = timeit(dec, a.reps)
    ok = bool(torch.equal(out, raw)) and bool((olen == bs).all())
    algo = (int(clen.sum()) + nb * bs)
    print(f"decode[{os.environ.get('K4LZ4_COPY_VARIANT','-')}]: {ms:.3f} ms  {nb*bs/ms/1e6:.1f} GB/s out  {algo/ms/1e6:.1f} GB/s algorithmic  ok={ok} ratio {ratio:.3f}")
