"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle.

Bit-exact for everything (integer/byte work): identical compressed bytes and return codes
for Encode at L00_FAST, identical bytes and return codes for Decode (well-formed and
malformed), identical pickles.  Mirrors the reference's BlockRoundtripTests / SpanTests /
PicklingTests on the accelerated path."""
import json
import os
import zlib

import numpy as np
import pytest

from tests import inputs

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def k4(native):
    import k4os.compression.lz4_b200 as k
    if native.k4lz4_device_count() <= 0:
        pytest.fail("no CUDA device: GPU tests must run on the B200 box")
    return k


@pytest.fixture(scope="module")
def chk():
    import oracle
    return oracle.best()


def test_loaded_library_is_in_tree(k4):
    from k4os.compression.lz4_b200 import _native
    assert os.path.dirname(_native.SO_PATH).endswith(os.path.join("k4os", "compression", "lz4_b200"))


def test_issue64_golden_decode_gpu(k4):
    comp = open(os.path.join(G, "issue64_block0.lz4"), "rb").read()
    expect = open(os.path.join(G, "issue64_block0.bin"), "rb").read()
    out = bytearray(65536)
    assert k4.LZ4Codec.Decode(comp, out) == 65536 and bytes(out) == expect
    big = bytearray(b"\xCD" * 131072)
    assert k4.LZ4Codec.Decode(comp, big) == 65536
    assert bytes(big[:65536]) == expect and bytes(big[65536:]) == b"\xCD" * 65536   # untouched tail
    assert k4.LZ4Codec.Decode(comp, bytearray(65535)) == -1


def test_golden_encode_rows_gpu(k4):
    """Every committed known-answer row, as one batch per capacity class."""
    rows = json.load(open(os.path.join(G, "encode_rows.json")))["rows"]
    datas = [inputs.gen(r["kind"], r["size"], r["seed"]) for r in rows]
    enc, lens = k4.batch.encode_batch_host(datas)
    for r, c, n in zip(rows, enc, lens):
        assert n == r["len"], (r["kind"], r["size"], n, r["len"])
        assert zlib.adler32(c) & 0xFFFFFFFF == r["adler32"], (r["kind"], r["size"])
    # limited-output capacities: same return codes as the reference engine
    lim_data, lim_caps, lim_expect = [], [], []
    for r, d in zip(rows, datas):
        for cap, e in r["limited"]:
            lim_data.append(d); lim_caps.append(cap); lim_expect.append(e)
    _, got = k4.batch.encode_batch_host(lim_data, lim_caps)
    assert got.tolist() == lim_expect
    # decode them back, exact and oversized capacity
    dec, dl = k4.batch.decode_batch_host(enc, [len(d) for d in datas])
    assert dl.tolist() == [len(d) for d in datas] and dec == datas
    dec2, dl2 = k4.batch.decode_batch_host(enc, [len(d) + 77 for d in datas])
    assert dl2.tolist() == [len(d) for d in datas] and dec2 == datas


def test_encode_matches_oracle_on_corpus(k4, chk):
    items = list(inputs.corpus(sizes=inputs.THRESHOLD_SIZES + inputs.BIG_SIZES))
    datas = [d for _, d in items]
    enc, lens = k4.batch.encode_batch_host(datas)
    for (name, d), c, n in zip(items, enc, lens):
        r, ref = chk.encode(d)
        assert (int(n), c) == (r, ref), name


def test_encode_never_writes_past_returned_length(k4):
    """PartialDecompressionTests.cs:33-35 / SpanTests.cs:36-37: 0xCD sentinels stay intact."""
    d = inputs.gen("text2", 20000, 5)
    tgt = bytearray(b"\xCD" * 30000)
    n = k4.LZ4Codec.Encode(d, 0, len(d), tgt, 100, 25000)
    assert n > 0 and bytes(tgt[:100]) == b"\xCD" * 100 and bytes(tgt[100 + n:]) == b"\xCD" * (30000 - 100 - n)


def test_decode_malformed_matches_oracle(k4):
    import oracle
    port = oracle.Port()     # the restatement of LL64.dec.cs is the authority for malformed input
    rng = np.random.default_rng(11)
    streams, caps = [], []
    for it in range(4000):
        n = int(rng.choice([20, 50, 100, 300, 1000, 5000, 70000]))
        kind = ["text2", "lowent", "runs", "random", "lorem"][it % 5]
        c = inputs.mutate(port.encode(inputs.gen(kind, n, it))[1], rng)
        streams.append(c)
        caps.append(int(rng.choice([n, n, n + 5, n - 1, 2 * n, n + 64, 1])))
    dec, got = k4.batch.decode_batch_host(streams, caps)
    for i, (c, cap) in enumerate(zip(streams, caps)):
        r, ref = port.decode(c, cap)
        assert int(got[i]) == r, (i, cap, int(got[i]), r)
        if r > 0 and not inputs.uses_zero_offset(c):      # offset-0 content is unspecified
            assert dec[i] == ref, i


def test_codec_edge_semantics_gpu(k4):
    C = k4.LZ4Codec
    assert C.Encode(b"", bytearray(10)) == 0 and C.Decode(b"", bytearray(10)) == 0
    assert C.Decode(b"\x00", bytearray(10)) == -1
    assert C.Encode(b"abc", bytearray(1)) == -1
    with pytest.raises(k4.DelegateToManagedEngine):
        C.Encode(b"a" * 100, bytearray(200), k4.LZ4Level.L09_HC)
    # QuickFox (BlockRoundtripTests.cs:44-61): decode into a 2x buffer returns the true size
    t = b"The quick brown fox jumps over the lazy dog"
    e = bytearray(C.MaximumOutputSize(len(t)))
    n = C.Encode(t, 0, len(t), e, 0, len(e))
    d = bytearray(2 * len(t))
    assert C.Decode(e, 0, n, d, 0, len(d)) == len(t) and bytes(d[:len(t)]) == t


def test_border_line_compression_gpu(k4):
    for kind in ("random", "text2", "synth525"):
        d = inputs.gen(kind, 65536, 11)
        tgt = bytearray(k4.LZ4Codec.MaximumOutputSize(len(d)))
        req = k4.LZ4Codec.Encode(d, tgt)
        tgt2 = bytearray(req)
        assert k4.LZ4Codec.Encode(d, tgt2) == req and tgt2 == tgt[:req]


def test_pickler_matches_oracle(k4):
    import oracle
    port = oracle.Port()
    rng = np.random.default_rng(3)
    msgs = [b"x", inputs.gen("random", 300, 1), b"a" * 200, b"a" * 5000, b"a" * 100000]
    for i in range(300):
        n = int(rng.integers(1, 4097)) if i % 3 else int(rng.choice([255, 256, 257, 1023, 1024, 1025, 4096]))
        msgs.append(inputs.gen(["text2", "synth435", "lorem", "random", "lowent"][i % 5], n, i))
    pk, lens = k4.batch.pickle_batch_host(msgs)
    for m, p in zip(msgs, pk):
        assert p == port.pickle(m), len(m)
    sizes = k4.batch.unpickled_size_batch_host(pk)
    assert sizes.tolist() == [len(m) for m in msgs]
    un, ul = k4.batch.unpickle_batch_host(pk)
    assert un == msgs and ul.tolist() == [len(m) for m in msgs]
    # single-message mirror + corruption -> InvalidDataException (PicklingTests.cs:149-172)
    P = k4.LZ4Pickler
    assert P.Pickle(b"") == b"" and P.Unpickle(b"") == b""
    good = P.Pickle(b"a" * 200)
    assert P.Unpickle(good) == b"a" * 200 and P.UnpickledSize(good) == 200
    for bad in (bytes([good[0] | 1]) + good[1:], good[:-1], b"\xC0\x01"):
        with pytest.raises(k4.InvalidDataException):
            P.Unpickle(bad)


def test_device_path_batch_roundtrip_and_host_generator(k4):
    """cfg-2/3 shape at reduced count: synth on device == synth on host; encode on device ==
    oracle; decode(encode(x)) == x; property: checksum over all blocks."""
    import torch
    import oracle
    chk = oracle.best()
    B = k4.batch
    nb, bs = 512, 65536
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    raw = torch.empty(nb * bs, dtype=torch.uint8, device=dev)
    B.synth_device(raw.data_ptr(), nb, bs, 435, 77, 5, st)
    host = B.synth_host(nb, bs, 435, 77, 5)
    assert np.array_equal(raw.cpu().numpy(), host)
    bound = k4.LZ4Codec.MaximumOutputSize(bs)
    idx = torch.arange(nb, dtype=torch.int64, device=dev)
    roff, coff = idx * bs, idx * bound
    rlen = torch.full((nb,), bs, dtype=torch.int32, device=dev)
    ccap = torch.full((nb,), bound, dtype=torch.int32, device=dev)
    comp = torch.full((nb * bound,), 0xCD, dtype=torch.uint8, device=dev)
    clen = torch.zeros(nb, dtype=torch.int32, device=dev)
    B.encode_batch_device(raw.data_ptr(), roff.data_ptr(), rlen.data_ptr(), comp.data_ptr(),
                          coff.data_ptr(), ccap.data_ptr(), clen.data_ptr(), nb, 0, st)
    out = torch.zeros(nb * bs, dtype=torch.uint8, device=dev)
    olen = torch.zeros(nb, dtype=torch.int32, device=dev)
    B.decode_batch_device(comp.data_ptr(), coff.data_ptr(), clen.data_ptr(), out.data_ptr(),
                          roff.data_ptr(), rlen.data_ptr(), olen.data_ptr(), nb, st)
    torch.cuda.synchronize()
    assert torch.equal(out, raw) and bool((olen == bs).all())
    c, l = comp.cpu().numpy(), clen.cpu().numpy()
    for i in range(0, nb, 7):
        r, ref = chk.encode(host[i * bs:(i + 1) * bs])
        assert r == int(l[i]) and c[i * bound:i * bound + r].tobytes() == ref
        assert (c[i * bound + r:(i + 1) * bound] == 0xCD).all()      # slot tail untouched
    # tight packing through copy_blocks, then decode from unaligned starts
    poff = torch.cumsum(clen.to(torch.int64), 0) - clen.to(torch.int64)
    packed = torch.empty(int(clen.sum()) + 64, dtype=torch.uint8, device=dev)
    B.copy_blocks_device(comp.data_ptr(), coff.data_ptr(), packed.data_ptr(), poff.data_ptr(), clen.data_ptr(), nb, st)
    out.zero_()
    B.decode_batch_device(packed.data_ptr(), poff.data_ptr(), clen.data_ptr(), out.data_ptr(),
                          roff.data_ptr(), rlen.data_ptr(), olen.data_ptr(), nb, st)
    torch.cuda.synchronize()
    assert torch.equal(out, raw) and bool((olen == bs).all())


def test_full_size_roundtrip_property(k4):
    """BASELINE configs[1]/[2] at FULL size (65 536 x 64 KiB = 4 GiB): size-independent
    properties -- decode(encode(x)) == x for every block, every length == 64 KiB, and the
    compressed size total lands on the expected ratio."""
    import torch
    B = k4.batch
    nb, bs = 65536, 65536
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    raw = torch.empty(nb * bs, dtype=torch.uint8, device=dev)
    B.synth_device(raw.data_ptr(), nb, bs, 525, 1234, 0, st)
    bound = k4.LZ4Codec.MaximumOutputSize(bs)
    idx = torch.arange(nb, dtype=torch.int64, device=dev)
    roff, coff = idx * bs, idx * bound
    rlen = torch.full((nb,), bs, dtype=torch.int32, device=dev)
    ccap = torch.full((nb,), bound, dtype=torch.int32, device=dev)
    comp = torch.empty(nb * bound, dtype=torch.uint8, device=dev)
    clen = torch.zeros(nb, dtype=torch.int32, device=dev)
    B.encode_batch_device(raw.data_ptr(), roff.data_ptr(), rlen.data_ptr(), comp.data_ptr(),
                          coff.data_ptr(), ccap.data_ptr(), clen.data_ptr(), nb, 0, st)
    out = torch.zeros(nb * bs, dtype=torch.uint8, device=dev)
    olen = torch.zeros(nb, dtype=torch.int32, device=dev)
    B.decode_batch_device(comp.data_ptr(), coff.data_ptr(), clen.data_ptr(), out.data_ptr(),
                          roff.data_ptr(), rlen.data_ptr(), olen.data_ptr(), nb, st)
    torch.cuda.synchronize()
    assert bool((olen == bs).all())
    assert torch.equal(out, raw)
    ratio = float(clen.sum()) / (nb * bs)
    assert 0.47 < ratio < 0.53, ratio


def test_host_path_all_devices_split_and_scattered_layout(k4, chk):
    """K4LZ4_ALL_DEVICES (NCCL-free contiguous split over every visible GPU; one device here is the
    degenerate case) and a scattered, gap-filled layout that forces the packed staging path: results
    identical to the single-device contiguous call, sentinels between slots untouched."""
    from k4os.compression.lz4_b200 import _native as N
    blocks = [inputs.gen(kind, n, i) for i, (kind, n) in enumerate(
        [("text2", 70000), ("random", 300), ("lorem", 65536), ("synth525", 65536), ("lowent", 4097),
         ("runs", 20000), ("repeat", 65536), ("text2", 13), ("text2", 12), ("synth435", 131072)] * 3)]
    ref_enc, ref_len = k4.batch.encode_batch_host(blocks)
    for i in (0, 3, 9):
        assert (int(ref_len[i]), ref_enc[i]) == chk.encode(blocks[i])
    # scattered layout: 4 KiB gaps between source blocks and between destination slots
    gap = 4096
    src_len = np.array([len(b) for b in blocks], dtype=np.int32)
    src_off = np.cumsum(np.concatenate([[gap], src_len[:-1].astype(np.int64) + gap])).astype(np.int64)
    src = np.full(int(src_off[-1] + src_len[-1] + gap), 0xEE, dtype=np.uint8)
    for o, b in zip(src_off, blocks):
        src[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    cap = np.array([k4.LZ4Codec.MaximumOutputSize(len(b)) for b in blocks], dtype=np.int32)
    dst_off = np.cumsum(np.concatenate([[gap], cap[:-1].astype(np.int64) + 64 * gap])).astype(np.int64)
    dst = np.full(int(dst_off[-1] + cap[-1] + gap), 0xCD, dtype=np.uint8)
    out = k4.batch.encode_batch_flat_host(src, src_off, src_len, dst, dst_off, cap, device=N.ALL_DEVICES)
    assert out.tolist() == ref_len.tolist()
    for i, (o, n) in enumerate(zip(dst_off, out)):
        assert dst[o:o + n].tobytes() == ref_enc[i]
        assert (dst[o + n:o + cap[i]] == 0xCD).all()
    # decode the scattered compressed slots back into exact-size slots (direct D2H path) and into
    # oversized slots (staging + scatter path)
    for extra in (0, 100):
        dcap = (src_len + extra).astype(np.int32)
        doff = np.cumsum(np.concatenate([[0], dcap[:-1].astype(np.int64)])).astype(np.int64)
        back = np.full(int(dcap.sum()) + 1, 0xCD, dtype=np.uint8)
        got = k4.batch.decode_batch_flat_host(dst, dst_off, out, back, doff, dcap, device=N.ALL_DEVICES)
        assert got.tolist() == src_len.tolist()
        for i, b in enumerate(blocks):
            assert back[doff[i]:doff[i] + len(b)].tobytes() == b
            assert (back[doff[i] + len(b):doff[i] + dcap[i]] == 0xCD).all()


def test_pickler_batch_at_scale_property(k4):
    """BASELINE configs[3] shape at 131 072 messages (256 B - 4 KiB mixed): unpickle(pickle(x)) == x,
    sizes agree, header bytes well-formed; 64 messages compared byte for byte with the oracle."""
    import torch
    import oracle
    port = oracle.Port()
    B = k4.batch
    n = 1 << 17
    rng = np.random.default_rng(42)
    sizes = np.where(rng.random(n) < 0.5, rng.choice([256, 512, 1024, 2048, 4096], n),
                     rng.integers(256, 4097, n)).astype(np.int32)
    off = np.zeros(n, dtype=np.int64); off[1:] = np.cumsum(sizes[:-1], dtype=np.int64)
    total = int(sizes.sum())
    dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
    raw = torch.empty(((total + 65535) // 65536) * 65536, dtype=torch.uint8, device=dev)
    B.synth_device(raw.data_ptr(), raw.numel() // 65536, 65536, 435, 42, 0, st)
    d_off, d_len = torch.from_numpy(off).to(dev), torch.from_numpy(sizes).to(dev)
    poff = np.zeros(n, dtype=np.int64); poff[1:] = np.cumsum(sizes[:-1].astype(np.int64) + 1)
    d_poff = torch.from_numpy(poff).to(dev)
    pk = torch.full((int(sizes.sum()) + n + 16,), 0xCD, dtype=torch.uint8, device=dev)
    plen = torch.zeros(n, dtype=torch.int32, device=dev)
    B.pickle_batch_device(raw.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), pk.data_ptr(), d_poff.data_ptr(),
                          plen.data_ptr(), n, 0, st)
    usz = torch.zeros(n, dtype=torch.int32, device=dev)
    B.unpickled_size_batch_device(pk.data_ptr(), d_poff.data_ptr(), plen.data_ptr(), usz.data_ptr(), n, st)
    out = torch.zeros(total + 16, dtype=torch.uint8, device=dev)
    olen = torch.zeros(n, dtype=torch.int32, device=dev)
    B.unpickle_batch_device(pk.data_ptr(), d_poff.data_ptr(), plen.data_ptr(), out.data_ptr(), d_off.data_ptr(),
                            d_len.data_ptr(), olen.data_ptr(), n, st)
    torch.cuda.synchronize()
    assert bool((usz == d_len).all()) and bool((olen == d_len).all())
    assert torch.equal(out[:total], raw[:total])
    assert bool((plen <= d_len + 1).all()) and bool((plen > 0).all())
    h_raw, h_pk, h_pl = raw[:int(off[64])].cpu().numpy(), pk[:int(poff[64])].cpu().numpy(), plen[:64].cpu().numpy()
    for i in range(64):
        assert h_pk[poff[i]:poff[i] + h_pl[i]].tobytes() == port.pickle(h_raw[off[i]:off[i] + sizes[i]].tobytes()), i
