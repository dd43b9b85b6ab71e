"""ctypes loaders for the checker libraries (oracle/ and oracle/_ref) and the product
C-ABI (brotli_b200/libbrotlienc_b200.so).  Test infrastructure only."""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libbrotli_ref.so")
PROBE_SO = os.path.join(ROOT, "oracle", "_ref", "libbrotli_probe.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
TABLES = os.path.join(ROOT, "brotli_b200", "data", "brotli_tables.bin")
PRODUCT_SO = os.path.join(ROOT, "brotli_b200", "libbrotlienc_b200.so")

_u8p = C.POINTER(C.c_uint8)


def _buf(b):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(b if len(b) else b"\0")


class Ref:
    """The unmodified reference compiled from /root/reference (oracle/Makefile)."""

    def __init__(self, path=REF_SO):
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL)
        L = self.lib
        L.BrotliEncoderCompress.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                            C.POINTER(C.c_size_t), C.c_void_p]
        L.BrotliEncoderCompress.restype = C.c_int
        L.BrotliEncoderMaxCompressedSize.argtypes = [C.c_size_t]
        L.BrotliEncoderMaxCompressedSize.restype = C.c_size_t
        L.BrotliDecoderDecompress.argtypes = [C.c_size_t, C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p]
        L.BrotliDecoderDecompress.restype = C.c_int

    def compress(self, data, quality, lgwin, mode=0):
        n = len(data)
        cap = self.lib.BrotliEncoderMaxCompressedSize(n) + 16
        out = C.create_string_buffer(cap)
        out_n = C.c_size_t(cap)
        ok = self.lib.BrotliEncoderCompress(quality, lgwin, mode, n, data, C.byref(out_n), out)
        assert ok == 1
        return out.raw[:out_n.value]

    def decompress(self, comp, max_out):
        out = C.create_string_buffer(max(1, max_out))
        out_n = C.c_size_t(max_out)
        ok = self.lib.BrotliDecoderDecompress(len(comp), comp, C.byref(out_n), out)
        assert ok == 1, "decoder rejected the stream"
        return out.raw[:out_n.value]


class Oracle:
    """Our CPU restatement (oracle/brotli_oracle.c)."""

    def __init__(self, path=ORACLE_SO):
        self.lib = C.CDLL(path, mode=os.RTLD_LOCAL)
        L = self.lib
        L.oracle_init.argtypes = [C.c_void_p, C.c_size_t]
        L.oracle_init.restype = C.c_int
        L.oracle_brotli_compress.argtypes = [C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                             C.POINTER(C.c_size_t), C.c_void_p]
        L.oracle_brotli_compress.restype = C.c_int
        L.oracle_fast_log2.argtypes = [C.c_size_t]
        L.oracle_fast_log2.restype = C.c_double
        blob = open(TABLES, "rb").read()
        assert L.oracle_init(blob, len(blob)) == 1

    def compress(self, data, quality, lgwin):
        n = len(data)
        cap = n + (n >> 1) + 4096
        out = C.create_string_buffer(cap)
        out_n = C.c_size_t(cap)
        ok = self.lib.oracle_brotli_compress(quality, lgwin, n, data, C.byref(out_n), out)
        assert ok == 1
        return out.raw[:out_n.value]
