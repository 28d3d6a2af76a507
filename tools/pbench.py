"""tools/pbench.py -- BASELINE configs[3]: LZ4Pickler.Pickle/Unpickle over 1 M small messages
(256 B - 4 KiB mixed, seed 42), device-resident, kernel-only timing + parity spot check."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from k4os.compression.lz4_b200 import batch as B

ap = argparse.ArgumentParser(); ap.add_argument("--n", type=int, default=1 << 20); a = ap.parse_args()
n = a.n
rng = np.random.default_rng(42)
sizes = np.where(rng.random(n) < 0.5, rng.choice([256, 512, 1024, 2048, 4096], n), rng.integers(256, 4097, n)).astype(np.int32)
off = np.zeros(n, dtype=np.int64); off[1:] = np.cumsum(sizes[:-1], dtype=np.int64)
total = int(sizes.sum())
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
raw = torch.empty(((total + 65535) // 65536) * 65536, dtype=torch.uint8, device=dev)
B.synth_device(raw.data_ptr(), raw.numel() // 65536, 65536, 435, 42, 0, st)      # messages = slices of the 0.57 stream
d_off = torch.from_numpy(off).to(dev); d_len = torch.from_numpy(sizes).to(dev)
bound = sizes.astype(np.int64) + 1
poff = np.zeros(n, dtype=np.int64); poff[1:] = np.cumsum(bound[:-1])
d_poff = torch.from_numpy(poff).to(dev)
pk = torch.empty(int(bound.sum()) + 16, dtype=torch.uint8, device=dev)
plen = torch.zeros(n, dtype=torch.int32, device=dev)
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms_p = t(lambda: B.pickle_batch_device(raw.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), pk.data_ptr(), d_poff.data_ptr(), plen.data_ptr(), n, 0, st))
out = torch.zeros(total + 16, dtype=torch.uint8, device=dev); olen = torch.zeros(n, dtype=torch.int32, device=dev)
usz = torch.zeros(n, dtype=torch.int32, device=dev)
B.unpickled_size_batch_device(pk.data_ptr(), d_poff.data_ptr(), plen.data_ptr(), usz.data_ptr(), n, st)
ms_u = t(lambda: B.unpickle_batch_device(pk.data_ptr(), d_poff.data_ptr(), plen.data_ptr(), out.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), olen.data_ptr(), n, st))
ok = bool(torch.equal(out[:total], raw[:total])) and bool((olen == d_len).all()) and bool((usz == d_len).all())
# parity spot check against the oracle (test infrastructure)
import oracle
P = oracle.Port(); h_raw = raw[:sizes[:64].sum()].cpu().numpy(); h_pk = pk.cpu().numpy()[:int(poff[64])]; h_pl = plen[:64].cpu().numpy()
for i in range(64):
    m = h_raw[off[i]:off[i] + sizes[i]].tobytes()
    assert h_pk[poff[i]:poff[i] + h_pl[i]].tobytes() == P.pickle(m), i
ratio = float(plen.sum()) / total
print(f"pickle: {ms_p:.2f} ms  {n/ms_p/1e3:.2f} M msg/s  {total/ms_p/1e6:.2f} GB/s | unpickle: {ms_u:.2f} ms  {n/ms_u/1e3:.2f} M msg/s  {total/ms_u/1e6:.2f} GB/s | ratio {ratio:.3f} ok={ok} n={n} bytes={total}")
