"""Exact-integer emulation of IEEE binary32 ops (round-to-nearest-even).

Independent of the C oracle and of numpy's float arithmetic: every finite f32
is an integer multiple of 2**-149, so sums/products are exact Python ints and
rounding is done by hand.  Used to pin the oracle's FMA / reduction order.
"""
import struct

SCALE = 149  # f32 value = n * 2**-149


def f32_to_int(x):
    """finite float32 (python float holding an f32 value) -> integer multiple of 2**-149"""
    bits = struct.unpack("<I", struct.pack("<f", x))[0]
    sign = -1 if bits >> 31 else 1
    e = (bits >> 23) & 0xFF
    m = bits & 0x7FFFFF
    assert e != 0xFF, "non-finite"
    if e == 0:
        return sign * m
    return sign * ((m | 0x800000) << (e - 1))


def round_int_to_f32(n, scale):
    """round n * 2**-scale to nearest-even binary32, returned as python float"""
    if n == 0:
        return 0.0
    sign = -1.0 if n < 0 else 1.0
    a = abs(n)
    # express as a * 2**-scale ; target quantum is 2**q with q >= -149
    bl = a.bit_length()
    # exponent of leading bit: bl - 1 - scale ; mantissa has 24 bits
    q = max(bl - 24 - scale, -149)  # value quantum exponent
    shift = q + scale               # bits to drop from a
    if shift <= 0:
        m = a << (-shift)
    else:
        m = a >> shift
        rem = a & ((1 << shift) - 1)
        half = 1 << (shift - 1)
        if rem > half or (rem == half and (m & 1)):
            m += 1
    val = m * (2.0 ** q)  # exact in double (m < 2**25, q >= -149)
    if val >= 2.0 ** 128:
        return sign * float("inf")
    return sign * val


def fma32(a, b, c):
    n = f32_to_int(a) * f32_to_int(b) + (f32_to_int(c) << SCALE)
    return round_int_to_f32(n, 2 * SCALE)


def mul32(a, b):
    return round_int_to_f32(f32_to_int(a) * f32_to_int(b), 2 * SCALE)


def add32(a, b):
    return round_int_to_f32(f32_to_int(a) + f32_to_int(b), SCALE)


def dot_f32_simd_order(a, b):
    """dot_product_f32_simd (x86_64.rs:418-444): 8 FMA lanes, hadd tree, scalar tail"""
    n = len(a)
    lanes = [0.0] * 8
    chunks = n // 8
    for i in range(chunks):
        for j in range(8):
            lanes[j] = fma32(float(a[8 * i + j]), float(b[8 * i + j]), lanes[j])
    lo = add32(add32(lanes[0], lanes[1]), add32(lanes[2], lanes[3]))
    hi = add32(add32(lanes[4], lanes[5]), add32(lanes[6], lanes[7]))
    r = add32(lo, hi)
    for i in range(chunks * 8, n):
        r = add32(r, mul32(float(a[i]), float(b[i])))
    return r


def sumsq_sequential(v):
    s = 0.0
    for x in v:
        s = add32(s, mul32(float(x), float(x)))
    return s


def sqrt32(x):
    """correctly rounded binary32 sqrt: the binary64 result rounded once more is exact because 53 >= 2*24+2"""
    import math
    return float(struct.unpack("<f", struct.pack("<f", math.sqrt(x)))[0])


def div32(a, b):
    """correctly rounded binary32 quotient (same double-rounding argument as sqrt32)"""
    return float(struct.unpack("<f", struct.pack("<f", a / b))[0])
