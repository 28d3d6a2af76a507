"""Independent-block stream pair with a BATCHED top-up.

Mirrors of the reference's only in-product callers of the block codec,
    Encoders/LZ4EncoderBase.cs:27-97, Encoders/LZ4BlockEncoder.cs:7-23, Encoders/LZ4BlockDecoder.cs:11-102,
with the same member names and error behaviour, plus the one thing a GPU needs: instead of one
`LZ4Codec.Encode` per 64 KiB block (one PCIe round trip each) the encoder queues up to
`batch_blocks` full blocks and encodes them with ONE `k4lz4_encode_batch` call; the decoder takes a
list of compressed blocks and decodes them with ONE `k4lz4_decode_batch` call.  Every block's bytes
and return value equal what the reference's per-block call produces (independent blocks: no
dictionary, fresh table per block -- LZ4BlockEncoder.cs:18-23).  Chained encoders (dependent
blocks) stay with the managed engine: they are not data parallel.
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .batch import decode_batch_flat_host, encode_batch_flat_host
from .codec import LZ4Codec, LZ4Level

K1 = 1024


def _round_up(v: int, step: int) -> int:      # Mem.RoundUp
    return (v + step - 1) // step * step


class LZ4BlockEncoder:
    """LZ4BlockEncoder(level, blockSize) -- LZ4BlockEncoder.cs:11-15; `batch_blocks` blocks are held
    back and encoded by one GPU call."""

    def __init__(self, level: LZ4Level = LZ4Level.L00_FAST, blockSize: int = 65536, batch_blocks: int = 256):
        self._level = LZ4Level(level)
        self._block = _round_up(max(int(blockSize), K1), K1)            # LZ4EncoderBase.cs:29
        self._depth = max(int(batch_blocks), 1)
        self._buf = np.zeros(self._depth * self._block, dtype=np.uint8)   # queue of block slots
        self._fill = np.zeros(self._depth, dtype=np.int32)               # bytes in each slot
        self._cur = 0                                                      # slot being topped up
        self._disposed = False

    # -- ILZ4Encoder ----------------------------------------------------------------------------
    @property
    def BlockSize(self) -> int:
        return self._block

    @property
    def BytesReady(self) -> int:
        """Bytes waiting in the block that is being filled (LZ4EncoderBase.cs:44)."""
        return int(self._fill[self._cur]) if self._cur < self._depth else 0

    @property
    def BlocksQueued(self) -> int:
        return int((self._fill > 0).sum())

    def Topup(self, source) -> int:
        """Adds bytes to the current block; returns how many were taken (0 when the block is full,
        LZ4EncoderBase.cs:47-62)."""
        self._check()
        src = np.frombuffer(source, dtype=np.uint8) if not isinstance(source, np.ndarray) else source
        if src.size == 0 or self._cur >= self._depth:
            return 0
        left = self._block - int(self._fill[self._cur])
        if left <= 0:
            return 0
        chunk = min(left, int(src.size))
        at = self._cur * self._block + int(self._fill[self._cur])
        self._buf[at:at + chunk] = src[:chunk]
        self._fill[self._cur] += chunk
        return chunk

    def TopupMany(self, source) -> int:
        """Batched top-up: fills block after block until the queue or the source is exhausted."""
        self._check()
        src = np.frombuffer(source, dtype=np.uint8) if not isinstance(source, np.ndarray) else source
        taken = 0
        while taken < src.size and self._cur < self._depth:
            got = self.Topup(src[taken:])
            taken += got
            if int(self._fill[self._cur]) == self._block:
                self._cur += 1
            elif got == 0:
                break
        return taken

    def Encode(self, target, allowCopy: bool = True) -> int:
        """Encodes the (single) pending block into `target` -- LZ4EncoderBase.cs:65-87: returns the
        encoded length, or -length when allowCopy stored the block raw; 0 when nothing is pending."""
        out = self.EncodeMany(allowCopy, _targets=[target])
        return out[0][0] if out else 0

    def EncodeMany(self, allowCopy: bool = True, _targets=None):
        """Encodes every queued block with one GPU call.  Returns [(encoded, bytes)] in block order
        with the reference's per-block convention (encoded < 0: stored raw, |encoded| bytes)."""
        self._check()
        nb = int((self._fill > 0).sum())
        if nb == 0:
            return []
        lens = self._fill[:nb].copy()
        if self._level >= LZ4Level.L03_HC:
            from .codec import DelegateToManagedEngine
            raise DelegateToManagedEngine("HC/OPT levels stay with the managed engine")
        bound = LZ4Codec.MaximumOutputSize(self._block)
        src_off = np.arange(nb, dtype=np.int64) * self._block
        if _targets is None:
            caps = np.full(nb, bound, dtype=np.int32)
        else:
            caps = np.array([len(t) for t in _targets], dtype=np.int32)
            assert len(_targets) == nb, "Encode() handles exactly one pending block"
        dst_off = np.zeros(nb, dtype=np.int64)
        dst_off[1:] = np.cumsum(caps[:-1].astype(np.int64))
        dst = np.zeros(int(caps.astype(np.int64).sum()) + 16, dtype=np.uint8)
        out_len = encode_batch_flat_host(self._buf, src_off, lens, dst, dst_off, caps, int(self._level))
        res = []
        for i in range(nb):
            enc = int(out_len[i])
            if enc <= 0:                                                  # LZ4EncoderBase.cs:75-77
                raise RuntimeError("Failed to encode chunk. Target buffer too small.")   # InvalidOperationException
            n = int(lens[i])
            if allowCopy and enc >= n:                                    # :79-83
                data = self._buf[i * self._block:i * self._block + n].tobytes()
                enc = -n
            else:
                data = dst[dst_off[i]:dst_off[i] + enc].tobytes()
            if _targets is not None:
                t = np.frombuffer(_targets[i], dtype=np.uint8) if not isinstance(_targets[i], np.ndarray) else _targets[i]
                t[:len(data)] = np.frombuffer(data, dtype=np.uint8)
            res.append((enc, data))
        self._fill[:] = 0                                                 # Commit(), :89-96 (no dictionary)
        self._cur = 0
        return res

    def Dispose(self) -> None:
        self._disposed = True

    def _check(self) -> None:
        if self._disposed:
            raise RuntimeError("ObjectDisposedException")


class LZ4BlockDecoder:
    """LZ4BlockDecoder(blockSize) -- LZ4BlockDecoder.cs:22-30, with DecodeMany for whole batches."""

    def __init__(self, blockSize: int = 65536):
        self._block = _round_up(max(int(blockSize), K1), K1)
        self._out_len = self._block + 8                                   # LZ4BlockDecoder.cs:26
        self._out = np.zeros(self._out_len + 8, dtype=np.uint8)
        self._index = 0
        self._disposed = False

    @property
    def BlockSize(self) -> int:
        return self._block

    @property
    def BytesReady(self) -> int:
        return self._index

    def Decode(self, source, blockSize: int = 0) -> int:
        """LZ4BlockDecoder.cs:39-55."""
        self._check()
        if blockSize <= 0:
            blockSize = self._block
        if blockSize > self._block:
            raise RuntimeError("InvalidOperationException")
        decoded = LZ4Codec.Decode(bytes(source), self._out[:self._out_len])
        if decoded < 0:
            raise RuntimeError("InvalidOperationException")
        self._index = decoded
        return decoded

    def DecodeMany(self, blocks):
        """Decodes a list of blocks with one GPU call.  An item is compressed bytes, or a tuple
        (bytes, True) for a block that was stored raw (the encoder's negative length).  Returns the
        list of decoded blocks; raises like Decode() if any block is malformed or larger than the
        block size.  The last block stays available through Drain/Peek."""
        self._check()
        comp, raw_at = [], {}
        for i, b in enumerate(blocks):
            if isinstance(b, tuple) and b[1]:
                raw_at[i] = bytes(b[0])
                comp.append(b"")
            else:
                comp.append(bytes(b[0] if isinstance(b, tuple) else b))
        n = len(comp)
        if n == 0:
            return []
        src = np.frombuffer(b"".join(comp) or b"\x00", dtype=np.uint8)
        lens = np.array([len(c) for c in comp], dtype=np.int32)
        off = np.zeros(n, dtype=np.int64)
        off[1:] = np.cumsum(lens[:-1].astype(np.int64))
        caps = np.full(n, self._out_len, dtype=np.int32)
        doff = np.arange(n, dtype=np.int64) * self._out_len
        dst = np.zeros(n * self._out_len + 16, dtype=np.uint8)
        out_len = decode_batch_flat_host(src, off, lens, dst, doff, caps)
        res = []
        for i in range(n):
            if i in raw_at:
                if len(raw_at[i]) > self._out_len:
                    raise RuntimeError("InvalidOperationException")
                res.append(raw_at[i])
                continue
            r = int(out_len[i])
            if r < 0 or (r == 0 and lens[i] > 0):
                raise RuntimeError("InvalidOperationException")
            res.append(dst[doff[i]:doff[i] + r].tobytes())
        last = res[-1]
        self._out[:len(last)] = np.frombuffer(last, dtype=np.uint8)
        self._index = len(last)
        return res

    def Inject(self, source) -> int:
        """LZ4BlockDecoder.cs:58-71."""
        self._check()
        n = len(source)
        if n <= 0:
            self._index = 0
            return 0
        if n > self._out_len:
            raise RuntimeError("InvalidOperationException")
        self._out[:n] = np.frombuffer(bytes(source), dtype=np.uint8)
        self._index = n
        return n

    def Drain(self, target, offset: int, length: int) -> None:
        """LZ4BlockDecoder.cs:74-83 (offset is negative: counted from the end of the block)."""
        self._check()
        offset = self._index + offset
        if offset < 0 or length < 0 or offset + length > self._index:
            raise RuntimeError("InvalidOperationException")
        t = np.frombuffer(target, dtype=np.uint8) if not isinstance(target, np.ndarray) else target
        t[:length] = self._out[offset:offset + length]

    def Peek(self, offset: int) -> np.ndarray:
        self._check()
        offset = self._index + offset
        if offset < 0 or offset > self._index:
            raise RuntimeError("InvalidOperationException")
        return self._out[offset:self._index]

    def Dispose(self) -> None:
        self._disposed = True

    def _check(self) -> None:
        if self._disposed:
            raise RuntimeError("ObjectDisposedException")
