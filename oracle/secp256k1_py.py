"""TEST INFRASTRUCTURE — independent pure-Python restatement of the arithmetic the
reference delegates to OpenSSL libcrypto (README.md:13 pins "3.2.0", the build
container has 3.0.13; not vendored in /root/reference):

  * the secp256k1 group (OpenSSL NID 714, reference SEAL/params.h:4) as published
    in SEC 2 v2 section 2.4.1: y^2 = x^3 + 7 over F_p, affine chord-and-tangent law;
  * EC_POINT_point2oct uncompressed / compressed encodings (SEC 1 section 2.3.3)
    as used at reference SEAL/hash.cpp:27-29;
  * the Fiat-Shamir challenge H(points..., id) of reference SEAL/hash.cpp:8-53:
    SHA-256 over the encodings followed by the 8-byte little-endian id, read as
    a big-endian integer and reduced modulo the group order;
  * the PA deterministic draw stream of oracle/shim/pa_seed_shim.h.

It exists so that the libcrypto-based checker (oracle/pa_oracle.c) and the CUDA
engine are both compared against something that shares no code with either.
Only tests/ , __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
it.  Pure Python integers: use it for small cases only.
"""
import hashlib

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
GX = 0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798
GY = 0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8
G = (GX, GY)
INF = None  # point at infinity


def on_curve(pt):
    if pt is INF:
        return True
    x, y = pt
    return (y * y - x * x * x - 7) % P == 0


def neg(pt):
    if pt is INF:
        return INF
    return (pt[0], (-pt[1]) % P)


def add(a, b):
    """EC_POINT_add (reference SEAL/bidder.cpp:130)."""
    if a is INF:
        return b
    if b is INF:
        return a
    x1, y1 = a
    x2, y2 = b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def mul(k, pt):
    """EC_POINT_mul(group, r, NULL, P, k, ctx) (reference SEAL/bidder.cpp:129); any k >= 0."""
    k %= N
    r = INF
    q = pt
    while k:
        if k & 1:
            r = add(r, q)
        q = add(q, q)
        k >>= 1
    return r


def lincomb(a, p, b, q):
    return add(mul(a, p), mul(b, q))


def enc64(pt):
    """Engine wire form: X||Y big-endian, infinity = 64 zero bytes."""
    if pt is INF:
        return bytes(64)
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def dec64(b):
    b = bytes(b)
    if b == bytes(64):
        return INF
    return (int.from_bytes(b[:32], "big"), int.from_bytes(b[32:], "big"))


def point2oct(pt, compressed=False):
    """EC_POINT_point2oct (reference SEAL/hash.cpp:27-29): infinity is the single byte 00."""
    if pt is INF:
        return b"\x00"
    if compressed:
        return bytes([2 + (pt[1] & 1)]) + pt[0].to_bytes(32, "big")
    return b"\x04" + enc64(pt)


def challenge(points, ident):
    """SHA256inNIZK* (reference SEAL/hash.cpp:8-53, 55-104, 106-162, 164-228).

    `points` does NOT include the generator; it is prepended here as the
    reference does (points[] = {generator, ...})."""
    h = hashlib.sha256()
    for pt in [G] + list(points):
        h.update(point2oct(pt))
    h.update(int(ident).to_bytes(8, "little"))  # raw size_t, SURVEY.md Q7
    return int.from_bytes(h.digest(), "big") % N


def pa_draw(seed, stream, ctr):
    msg = b"PAv1" + seed.to_bytes(8, "little") + stream.to_bytes(8, "little") + ctr.to_bytes(8, "little")
    return int.from_bytes(hashlib.sha256(msg).digest(), "big")


REJECT_BITS = 0


class PaStream:
    """BN_rand_range(., order) replacement of oracle/shim/pa_seed_shim.cpp."""

    def __init__(self, seed, stream, ctr=0):
        self.seed, self.stream, self.ctr = seed, stream, ctr

    def rand_range(self, rng=N):
        while True:
            v = pa_draw(self.seed, self.stream, self.ctr)
            self.ctr += 1
            # REJECT_BITS mirrors the engine's test hook PA_DBG_REJECT_BITS: a draw whose top k bits are all
            # ones counts as out of range, so that the redraw paths are exercised (normally 2^-128 per draw)
            if REJECT_BITS and (v >> (256 - REJECT_BITS)) == (1 << REJECT_BITS) - 1:
                continue
            if v < rng:
                return v

    def rand256(self):
        v = pa_draw(self.seed, self.stream, self.ctr)
        self.ctr += 1
        return v


def derive_bid(seed, j, c):
    d = hashlib.sha256(b"PAbid" + seed.to_bytes(8, "little") + j.to_bytes(8, "little")).digest()
    return int.from_bytes(d[:8], "little") & ((1 << min(c, 31)) - 1)
