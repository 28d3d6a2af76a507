"""oracle -- CPU checker for the K4os LZ4 block hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(``k4os.compression.lz4_b200``) never does.

Two engines are exposed through the same thin ctypes surface:

* ``port``  -- ``oracle/_build/libk4oracle.so``: the C restatement in
  ``k4lz4_oracle.c`` (each function cites the reference file:line it follows);
* ``ref``   -- ``oracle/_ref/libk4ref.so``: the reference's own upstream C engine
  (``/root/reference/orig/lib/lz4.c``) compiled as-is by ``oracle/Makefile``.  It exists
  only where it was built (this container) or shipped prebuilt (the GPU box).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "_build", "libk4oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libk4ref.so")
PICKLE_CORRUPT = -1000

_u8p = C.POINTER(C.c_uint8)


def build(quiet: bool = True) -> None:
    """Compile the restatement and (when /root/reference is present) the reference engine."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_u8p)


def _as_u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8)


class _Harness:
    def __init__(self, lib):
        self._lib = lib
        lib.k4h_run_batch.restype = C.c_double
        lib.k4h_run_batch.argtypes = [
            C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_void_p, C.c_void_p, C.c_int64, C.c_int]

    _datagen_sym = "k4o_datagen"

    def datagen(self, size: int, match_proba: float, lit_proba: float = 0.0, seed: int = 1234,
                out: np.ndarray | None = None) -> np.ndarray:
        """RDG_genBuffer(buf, size, matchProba, litProba, seed) of orig/programs/datagen.c:156-162
        (0.63 -> configs[1], 0.55 -> configs[2]).  One serial stream."""
        fn = getattr(self._lib, self._datagen_sym)
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_uint32]
        buf = out if out is not None else np.empty(size, dtype=np.uint8)
        assert buf.dtype == np.uint8 and buf.size >= size and buf.flags.c_contiguous
        if fn(buf.ctypes.data, size, float(match_proba), float(lit_proba), int(seed)) != 0:
            raise ValueError("datagen: bad arguments")
        return buf

    def run_batch(self, mode, src, src_off, src_len, dst, dst_off, dst_cap, out_len, threads):
        """mode 0 = encode, 1 = decode.  All arrays numpy (u8 / i64 / i32).  Returns seconds."""
        n = int(src_len.shape[0])
        return float(self._lib.k4h_run_batch(
            mode, src.ctypes.data, src_off.ctypes.data, src_len.ctypes.data,
            dst.ctypes.data, dst_off.ctypes.data, dst_cap.ctypes.data,
            out_len.ctypes.data, n, int(threads)))


class Port(_Harness):
    """The C restatement (k4lz4_oracle.c)."""
    kind = "port"

    def __init__(self):
        if not os.path.exists(PORT_SO):
            build()
        lib = C.CDLL(PORT_SO)
        super().__init__(lib)
        lib.k4o_compress_fast.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int]
        lib.k4o_decompress_safe.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        lib.k4o_codec_encode.argtypes = [_u8p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int]
        lib.k4o_codec_decode.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        lib.k4o_codec_decode_dict.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int]
        lib.k4o_codec_partial_decode.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        lib.k4o_pickle.argtypes = [_u8p, C.c_int, _u8p, _u8p, C.c_int]
        lib.k4o_unpickled_size.argtypes = [_u8p, C.c_int]
        lib.k4o_unpickle.argtypes = [_u8p, C.c_int, _u8p, C.c_int]

    def max_output_size(self, n: int) -> int:
        return int(self._lib.k4o_max_output_size(n))

    def encode(self, src, cap: int | None = None, level: int = 0, enforce32: bool = False):
        """LZ4Codec.Encode semantics -> (ret, bytes)."""
        s = _as_u8(src)
        n = int(s.shape[0])
        if cap is None:
            cap = self.max_output_size(n)
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        r = int(self._lib.k4o_codec_encode(_ptr(s), n, _ptr(d), cap, level, int(enforce32)))
        return r, (d[:r].tobytes() if r > 0 else b"")

    def decode(self, src, cap: int):
        """LZ4Codec.Decode semantics -> (ret, bytes)."""
        s = _as_u8(src)
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        r = int(self._lib.k4o_codec_decode(_ptr(s), int(s.shape[0]), _ptr(d), cap))
        return r, (d[:r].tobytes() if r > 0 else b"")

    def decode_dict(self, src, cap: int, dictionary):
        """LZ4Codec.Decode(src, dst, dict) semantics (LZ4Codec.cs:144-157) -> (ret, bytes)."""
        s = _as_u8(src)
        dd = _as_u8(dictionary)
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        r = int(self._lib.k4o_codec_decode_dict(_ptr(s), int(s.shape[0]), _ptr(d), cap, _ptr(dd), int(dd.shape[0])))
        return r, (d[:r].tobytes() if r > 0 else b"")

    def partial_decode(self, src, target: int):
        """LZ4Codec.PartialDecode semantics (LZ4Codec.cs:123-134) -> (ret, bytes)."""
        s = _as_u8(src)
        d = np.zeros(max(target, 1), dtype=np.uint8)
        r = int(self._lib.k4o_codec_partial_decode(_ptr(s), int(s.shape[0]), _ptr(d), target))
        return r, (d[:r].tobytes() if r > 0 else b"")

    def decompress_safe_raw(self, src, cap: int) -> int:
        """Engine-level return value (negative error position), LL64.dec.cs:465."""
        s = _as_u8(src)
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        return int(self._lib.k4o_decompress_safe(_ptr(s), int(s.shape[0]), _ptr(d), cap))

    def pickle(self, src, level: int = 0) -> bytes:
        s = _as_u8(src)
        n = int(s.shape[0])
        if n == 0:
            return b""
        d = np.zeros(n + 1, dtype=np.uint8)
        scratch = np.zeros(max(n, 1024), dtype=np.uint8)
        r = int(self._lib.k4o_pickle(_ptr(s), n, _ptr(d), _ptr(scratch), level))
        return d[:r].tobytes()

    def pickle_writer(self, src, level: int = 0) -> bytes:
        """LZ4Pickler.Pickle<TBufferWriter> (LZ4Pickler.pickle.cs:113-148): bytes advanced in the writer."""
        s = _as_u8(src)
        n = int(s.shape[0])
        if n == 0:
            return b""
        self._lib.k4o_pickle_writer.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        d = np.zeros(int(self._lib.k4o_pickle_writer_bound(n)), dtype=np.uint8)
        r = int(self._lib.k4o_pickle_writer(_ptr(s), n, _ptr(d), level))
        return d[:r].tobytes()

    def unpickled_size(self, src) -> int:
        s = _as_u8(src)
        return int(self._lib.k4o_unpickled_size(_ptr(s), int(s.shape[0])))

    def unpickle(self, src):
        """-> (ret, bytes); ret == PICKLE_CORRUPT where the reference throws."""
        s = _as_u8(src)
        n = int(s.shape[0])
        if n == 0:
            return 0, b""
        size = self.unpickled_size(s)
        if size < 0:
            return PICKLE_CORRUPT, b""
        d = np.zeros(max(size, 1), dtype=np.uint8)
        r = int(self._lib.k4o_unpickle(_ptr(s), n, _ptr(d), size))
        return r, (d[:r].tobytes() if r > 0 else b"")


class Ref(_Harness):
    """The reference's own upstream C engine (orig/lib/lz4.c), LZ4Codec post-processing applied."""
    kind = "reference"
    _datagen_sym = "k4ref_datagen"

    def __init__(self):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO)
        lib = C.CDLL(REF_SO)
        super().__init__(lib)
        lib.k4ref_compress_fast.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        lib.k4ref_decompress_safe.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        lib.k4ref_decompress_safe_usingDict.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _u8p, C.c_int]
        lib.k4ref_decompress_safe_partial.argtypes = [_u8p, C.c_int, _u8p, C.c_int]

    def version(self) -> int:
        return int(self._lib.k4ref_version())

    def encode(self, src, cap: int | None = None):
        s = _as_u8(src)
        n = int(s.shape[0])
        if n <= 0:
            return 0, b""
        if cap is None:
            cap = n + n // 255 + 16
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        r = int(self._lib.k4ref_compress_fast(_ptr(s), n, _ptr(d), cap))
        r = -1 if r <= 0 else r
        return r, (d[:r].tobytes() if r > 0 else b"")

    def decode(self, src, cap: int):
        s = _as_u8(src)
        n = int(s.shape[0])
        if n <= 0:
            return 0, b""
        d = np.zeros(max(cap, 1), dtype=np.uint8)
        r = int(self._lib.k4ref_decompress_safe(_ptr(s), n, _ptr(d), cap))
        r = -1 if r <= 0 else r
        return r, (d[:r].tobytes() if r > 0 else b"")


def _ref_decode_dict(self, src, cap: int, dictionary):
    s = _as_u8(src)
    n = int(s.shape[0])
    if n <= 0:
        return 0, b""
    dd = _as_u8(dictionary)
    d = np.zeros(max(cap, 1), dtype=np.uint8)
    r = int(self._lib.k4ref_decompress_safe_usingDict(_ptr(s), n, _ptr(d), cap, _ptr(dd), int(dd.shape[0])))
    r = -1 if r <= 0 else r
    return r, (d[:r].tobytes() if r > 0 else b"")


def _ref_partial_decode(self, src, target: int):
    s = _as_u8(src)
    n = int(s.shape[0])
    if n <= 0:
        return 0, b""
    d = np.zeros(max(target, 1), dtype=np.uint8)
    r = int(self._lib.k4ref_decompress_safe_partial(_ptr(s), n, _ptr(d), target))
    r = -1 if r <= 0 else r
    return r, (d[:r].tobytes() if r > 0 else b"")


Ref.decode_dict = _ref_decode_dict
Ref.partial_decode = _ref_partial_decode


def _ref_xxh32(self, data, seed: int = 0) -> int:
    s = _as_u8(data)
    self._lib.k4ref_xxh32.restype = C.c_uint32
    self._lib.k4ref_xxh32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    return int(self._lib.k4ref_xxh32(s.ctypes.data, int(s.shape[0]), seed))


def _ref_frame_compress(self, data, block_checksum: bool = False, content_checksum: bool = False) -> bytes:
    """upstream LZ4F_compressFrame: independent 64 KiB blocks, acceleration 1 (orig/lib/lz4frame.c)."""
    s = _as_u8(data)
    L = self._lib
    L.k4ref_frame_bound.restype = C.c_size_t; L.k4ref_frame_bound.argtypes = [C.c_size_t]
    L.k4ref_frame_compress.restype = C.c_size_t
    L.k4ref_frame_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    cap = int(L.k4ref_frame_bound(int(s.shape[0])))
    d = np.zeros(cap, dtype=np.uint8)
    r = int(L.k4ref_frame_compress(s.ctypes.data, int(s.shape[0]), d.ctypes.data, cap, int(block_checksum), int(content_checksum)))
    if r == 0:
        raise RuntimeError("LZ4F_compressFrame failed")
    return d[:r].tobytes()


def _ref_frame_decompress(self, frame, cap: int):
    """upstream LZ4F_decompress over a whole frame -> bytes, or None on any error."""
    s = _as_u8(frame)
    L = self._lib
    L.k4ref_frame_decompress.restype = C.c_size_t
    L.k4ref_frame_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    d = np.zeros(max(cap, 1), dtype=np.uint8)
    r = int(L.k4ref_frame_decompress(s.ctypes.data, int(s.shape[0]), d.ctypes.data, cap))
    return None if r == (1 << 64) - 1 else d[:r].tobytes()


Ref.xxh32 = _ref_xxh32
Ref.frame_compress = _ref_frame_compress
Ref.frame_decompress = _ref_frame_decompress


def _port_xxh32(self, data, seed: int = 0) -> int:
    s = _as_u8(data)
    self._lib.k4o_xxh32.restype = C.c_uint32
    self._lib.k4o_xxh32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    return int(self._lib.k4o_xxh32(s.ctypes.data, int(s.shape[0]), seed))


Port.xxh32 = _port_xxh32


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def best():
    """The strongest checker available here: the compiled reference if present, else the port."""
    return Ref() if have_ref() else Port()
