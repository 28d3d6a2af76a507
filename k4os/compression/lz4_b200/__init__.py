"""k4os.compression.lz4_b200 -- B200-native (sm_100a CUDA) drop-in for one hot path of
K4os.Compression.LZ4: ``LZ4Codec.Encode`` at ``L00_FAST``, ``LZ4Codec.Decode`` and
``LZ4Pickler.Pickle/Unpickle`` over batches of independent blocks (plus ``Decode`` with a
dictionary, ``PartialDecode`` and the independent-block ``LZ4BlockEncoder`` / ``LZ4BlockDecoder``
pair with a batched top-up).

The product is ``libk4lz4.so`` (C ABI in ``include/k4lz4.h``, kernels in ``csrc/``); this
package is the host-side mirror of the reference's public interface for that path plus the
batched entry points.  Importing it never imports ``oracle/``.
"""
from . import _native
from .codec import LZ4Codec, LZ4Level, DelegateToManagedEngine
from .pickler import LZ4Pickler, InvalidDataException
from . import batch
from .encoders import LZ4BlockEncoder, LZ4BlockDecoder

__all__ = ["LZ4Codec", "LZ4Level", "LZ4Pickler", "InvalidDataException",
           "DelegateToManagedEngine", "batch", "_native", "LZ4BlockEncoder", "LZ4BlockDecoder"]
