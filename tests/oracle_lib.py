"""TEST INFRASTRUCTURE — ctypes loader of the CPU checker oracle/libpa_oracle.so
(Tier-B restatement on libcrypto, oracle/pa_oracle.c).  Builds it on first use
(gcc + system libcrypto are present on the build container and the GPU box)."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libpa_oracle.so")


def _ensure_built():
    src = os.path.join(ORACLE_DIR, "pa_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def _p(b):
    if isinstance(b, bytearray):
        return ctypes.cast((ctypes.c_uint8 * len(b)).from_buffer(b), ctypes.c_void_p)
    return ctypes.cast(ctypes.c_char_p(bytes(b)), ctypes.c_void_p)


class Oracle:
    def __init__(self):
        _ensure_built()
        self.lib = ctypes.CDLL(LIB)

    def _call(self, name, *args):
        fn = getattr(self.lib, name)
        fn.restype = ctypes.c_int
        rc = fn(*args)
        if rc != 0:
            raise RuntimeError(f"{name} failed: {rc}")

    def curve_constants(self):
        bufs = [bytearray(32) for _ in range(4)]
        self._call("po_curve_constants", *[_p(b) for b in bufs])
        return tuple(int.from_bytes(b, "big") for b in bufs)

    def fixed_base_mul(self, scalars):
        n = len(scalars) // 32
        out = bytearray(64 * n)
        self._call("po_fixed_base_mul", _p(scalars), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def var_base_mul(self, points, scalars):
        n = len(scalars) // 32
        out = bytearray(64 * n)
        self._call("po_var_base_mul", _p(points), _p(scalars), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def double_mul(self, a, points, b):
        n = len(a) // 32
        out = bytearray(64 * n)
        self._call("po_double_mul", _p(a), _p(points), _p(b), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def lincomb2(self, p, a, q, b):
        n = len(a) // 32
        out = bytearray(64 * n)
        self._call("po_lincomb2", _p(p), _p(a), _p(q), _p(b), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def point_add(self, p, q, sub=False):
        n = len(p) // 64
        out = bytearray(64 * n)
        self._call("po_point_add", _p(p), _p(q), _p(out), ctypes.c_size_t(n), ctypes.c_int(1 if sub else 0))
        return bytes(out)

    def point_encode(self, points, compressed=False, stride=None):
        n = len(points) // 64
        stride = stride or (33 if compressed else 65)
        out = bytearray(stride * n)
        lens = (ctypes.c_uint32 * max(n, 1))()
        self._call("po_point_encode", _p(points), ctypes.c_size_t(n), ctypes.c_int(1 if compressed else 0), _p(out),
                   ctypes.c_size_t(stride), ctypes.cast(lens, ctypes.c_void_p))
        return [bytes(out[i * stride:i * stride + lens[i]]) for i in range(n)]


def _ids(ids):
    return (ctypes.c_uint64 * max(len(ids), 1))(*ids)


def _add_protocol_methods(cls):
    """Protocol-level restatement functions (oracle/pa_oracle.c) — same argument
    order and layouts as the engine's ABI."""

    def challenge(self, points, k, ids):
        n = len(ids)
        out = bytearray(32 * n)
        self._call("po_challenge", _p(points), ctypes.c_size_t(k), _ids(ids), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def commit_points(self, alpha, beta, bits):
        n = len(bits)
        out = bytearray(192 * n)
        self._call("po_commit_points", _p(alpha), _p(beta), _p(bits), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def pokdlog_prove(self, X, x, ids, rnd):
        n = len(ids)
        out = bytearray(96 * n)
        self._call("po_pokdlog_prove", _p(X), _p(x), _ids(ids), _p(rnd), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def pokdlog_verify(self, proofs, X, ids):
        n = len(ids)
        out = bytearray(n)
        self._call("po_pokdlog_verify", _p(proofs), _p(X), _ids(ids), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def powfcom_prove(self, stmt, alpha, bits, ids, rnd):
        n = len(ids)
        out = bytearray(352 * n)
        self._call("po_powfcom_prove", _p(stmt), _p(alpha), _p(bits), _ids(ids), _p(rnd), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def powfcom_verify(self, proofs, stmt, ids):
        n = len(ids)
        out = bytearray(n)
        self._call("po_powfcom_verify", _p(proofs), _p(stmt), _ids(ids), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def stage1_prove(self, stmt, secrets, bits, ids, rnd):
        n = len(ids)
        out = bytearray(672 * n)
        self._call("po_stage1_prove", _p(stmt), _p(secrets), _p(bits), _ids(ids), _p(rnd), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def stage1_verify(self, proofs, stmt, ids):
        n = len(ids)
        out = bytearray(n)
        self._call("po_stage1_verify", _p(proofs), _p(stmt), _ids(ids), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def stage2_prove(self, stmt, secrets, bi, bj, ids, rnd):
        n = len(ids)
        out = bytearray(1344 * n)
        self._call("po_stage2_prove", _p(stmt), _p(secrets), _p(bi), _p(bj), _ids(ids), _p(rnd), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def stage2_verify(self, proofs, stmt, ids):
        n = len(ids)
        out = bytearray(n)
        self._call("po_stage2_verify", _p(proofs), _p(stmt), _ids(ids), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def ccs22_setup_hash(self, scalars, k):
        n = len(scalars) // (32 * k)
        out = bytearray(32 * n)
        self._call("po_ccs22_setup_hash", _p(scalars), ctypes.c_size_t(k), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def y_scan(self, X):
        n = len(X) // 64
        out = bytearray(64 * n)
        self._call("po_y_scan", _p(X), _p(out), ctypes.c_size_t(n))
        return bytes(out)

    def point_sum_is_inf(self, b):
        flag = ctypes.c_int(0)
        self._call("po_point_sum_is_inf", _p(b), ctypes.c_size_t(len(b) // 64), ctypes.byref(flag))
        return bool(flag.value)

    for name, fn in list(locals().items()):
        if callable(fn) and name != "cls":
            setattr(cls, name, fn)
    return cls


_add_protocol_methods(Oracle)
