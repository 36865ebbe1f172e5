"""zlib-streams-ts_b200 -- B200-native (sm_100a) deflate / inflate / checksum engine behind the
zlib-streams-ts API.

Layers (mirroring the reference's):
  capi      ctypes binding of the C ABI (include/zsgpu.h, libzsgpu.so) -- no CPU fallback
  batch     deflateBatch / inflateBatch / checksum entry points (new, feed the GPU)
  zlib_api  createDeflateStream / deflateInit / deflate / deflateEnd / inflateInit2_ / inflate ...
  streams   CompressionStream / DecompressionStream
  sharded   contiguous chunk ranges over the ranks of one box (torch.distributed)

The directory name contains a hyphen (it is the reference's name); import it with
``importlib.import_module("zlib-streams-ts_b200")``.
"""
from . import capi  # noqa: F401
from .capi import (Z_BUF_ERROR, Z_DATA_ERROR, Z_FINISH, Z_FULL_FLUSH, Z_MEM_ERROR, Z_NEED_DICT, Z_NO_FLUSH,  # noqa: F401
                   Z_OK, Z_PARTIAL_FLUSH, Z_STREAM_END, Z_STREAM_ERROR, Z_SYNC_FLUSH, Z_BLOCK)

__all__ = ["capi"]
