"""Reader / writer of the .cdbf container written by tools/ref_fixtures/b200_fixtures.rs (the reference-side fixture
generator).  Record: u32 name_len, name, u32 dtype (0 u8, 1 i8, 2 u32, 3 f32), u32 ndim, u64 dims[ndim], payload (LE)."""
import struct

import numpy as np

DTYPES = {0: np.uint8, 1: np.int8, 2: np.uint32, 3: np.float32}
CODES = {np.dtype(v): k for k, v in DTYPES.items()}


def load_cdbf(path):
    out = {}
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        (nl,) = struct.unpack_from("<I", data, pos); pos += 4
        name = data[pos:pos + nl].decode("utf-8"); pos += nl
        dt, nd = struct.unpack_from("<II", data, pos); pos += 8
        dims = struct.unpack_from(f"<{nd}Q", data, pos); pos += 8 * nd
        dtype = np.dtype(DTYPES[dt])
        count = int(np.prod(dims)) if nd else 1
        out[name] = np.frombuffer(data, dtype=dtype, count=count, offset=pos).reshape(dims).copy()
        pos += count * dtype.itemsize
    return out


def write_cdbf(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            nb = name.encode("utf-8")
            f.write(struct.pack("<I", len(nb)) + nb + struct.pack("<II", CODES[a.dtype], a.ndim))
            f.write(struct.pack(f"<{a.ndim}Q", *a.shape))
            f.write(a.tobytes())
