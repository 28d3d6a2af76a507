"""The C-ABI shared library: loads, exports every symbol include/k4lz4.h declares, and its
non-compute entry points behave.  CPU only -- no compute call is made without a GPU."""
import ctypes
import os
import re

import numpy as np

from tests.conftest import ROOT, has_gpu


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "k4lz4.h")).read()
    return sorted(set(re.findall(r"K4LZ4_API\s+[\w\s\*]+?\b(k4lz4_\w+)\s*\(", text)))


def test_header_symbols_all_exported(native):
    from k4os.compression.lz4_b200 import _native
    decl = _declared_symbols()
    assert len(decl) >= 15
    assert sorted(_native.SYMBOLS) == decl
    raw = ctypes.CDLL(_native.SO_PATH)
    for s in decl:
        assert hasattr(raw, s), s


def test_information_entry_points(native):
    assert native.k4lz4_codec_version() == 192          # LZ4Codec.cs:13
    assert native.k4lz4_device_count() >= 0
    assert native.k4lz4_max_output_size(65536) == 65809
    assert native.k4lz4_max_output_size(0) == 16
    assert native.k4lz4_max_output_size(0x7E000001) == 0
    assert native.k4lz4_pickle_bound(0) == 0 and native.k4lz4_pickle_bound(100) == 101


def test_no_device_fails_loudly(native):
    """Without a GPU the product refuses to run: no silent CPU fallback."""
    if has_gpu():
        return
    from k4os.compression.lz4_b200 import LZ4Codec, _native
    import pytest
    with pytest.raises(_native.K4Error) as e:
        LZ4Codec.Encode(b"some bytes some bytes some bytes", bytearray(100))
    assert e.value.code == _native.E_NODEVICE
    # reference semantics that need no device still hold
    assert LZ4Codec.Encode(b"", bytearray(10)) == 0
    assert LZ4Codec.Decode(b"", bytearray(10)) == 0


def test_synth_host_is_deterministic(native):
    from k4os.compression.lz4_b200.batch import synth_host
    a = synth_host(3, 4096, 525, seed=9)
    b = synth_host(3, 4096, 525, seed=9)
    c = synth_host(1, 4096, 525, seed=9, first_block=2)
    assert np.array_equal(a, b) and np.array_equal(a[8192:], c)
    assert not np.array_equal(a[:4096], a[4096:8192])


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (checked textually over its sources)."""
    pkg = os.path.join(ROOT, "k4os")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "k4lz4_oracle" not in text and "libk4ref" not in text, f
