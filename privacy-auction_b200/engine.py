"""ctypes binding of the engine's C ABI (include/pa_engine.h).

This is plumbing for the tests and bench.py: every method forwards to the
symbol of the same name in libpa_engine.so.  There is no Python or CPU
implementation behind it — if the library or a B200 is missing, construction
fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PA_ENGINE_LIB overrides the library path (kernel-variant experiments); the default is the in-tree build
LIB_PATH = os.environ.get("PA_ENGINE_LIB") or os.path.join(_HERE, "libpa_engine.so")

PA_OK, PA_EINVAL, PA_ENODEV, PA_ECUDA, PA_ENOMEM = 0, -1, -2, -3, -4
POINT_BYTES, SCALAR_BYTES = 64, 32

_u8p = ctypes.POINTER(ctypes.c_uint8)
_sz = ctypes.c_size_t
_ctx = ctypes.c_void_p
_vp = ctypes.c_void_p

# symbol -> (restype, argtypes); kept in step with include/pa_engine.h (tests/test_abi.py checks)
SIGNATURES = {
    "pa_ctx_create": (ctypes.c_int, [ctypes.POINTER(_ctx), ctypes.c_int]),
    "pa_ctx_destroy": (ctypes.c_int, [_ctx]),
    "pa_sync": (ctypes.c_int, [_ctx]),
    "pa_last_error": (ctypes.c_char_p, [_ctx]),
    "pa_abi_version": (ctypes.c_int, []),
    "pa_abi_sizeof": (ctypes.c_size_t, [ctypes.c_int]),
    "pa_ctx_stream": (ctypes.c_void_p, [_ctx]),
    "pa_ctx_launches": (ctypes.c_uint64, [_ctx]),
    "pa_dev_alloc": (ctypes.c_int, [_ctx, ctypes.POINTER(ctypes.c_void_p), _sz]),
    "pa_dev_free": (ctypes.c_int, [_ctx, ctypes.c_void_p]),
    "pa_dev_upload": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_dev_download": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_fixed_base_mul": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_fixed_base_mul_dev": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_var_base_mul": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_var_base_mul_dev": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_double_mul": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_double_mul_dev": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _sz]),
    "pa_lincomb2": (ctypes.c_int, [_ctx] + [ctypes.c_void_p] * 5 + [_sz]),
    "pa_lincomb2_dev": (ctypes.c_int, [_ctx] + [ctypes.c_void_p] * 5 + [_sz]),
    "pa_point_add": (ctypes.c_int, [_ctx, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _sz, ctypes.c_int]),
    "pa_point_on_curve": (ctypes.c_int, [_ctx, ctypes.c_void_p, _sz, ctypes.c_void_p]),
    "pa_point_encode": (ctypes.c_int, [_ctx, ctypes.c_void_p, _sz, ctypes.c_int, ctypes.c_void_p, _sz, ctypes.c_void_p]),
    "pa_measure_int_peak": (ctypes.c_int, [_ctx, ctypes.POINTER(ctypes.c_double)]),
    "pa_challenge": (ctypes.c_int, [_ctx, _vp, _sz, _vp, _vp, _sz]),
    "pa_challenge_dev": (ctypes.c_int, [_ctx, _vp, _sz, _vp, _vp, _sz]),
    "pa_pokdlog_prove": (ctypes.c_int, [_ctx] + [_vp] * 5 + [_sz]),
    "pa_pokdlog_prove_dev": (ctypes.c_int, [_ctx] + [_vp] * 5 + [_sz]),
    "pa_pokdlog_verify": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_pokdlog_verify_dev": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_powfcom_prove": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_powfcom_prove_dev": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_powfcom_verify": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_powfcom_verify_dev": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_stage1_prove": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_stage1_prove_dev": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_stage1_verify": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_stage1_verify_dev": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_stage2_prove": (ctypes.c_int, [_ctx] + [_vp] * 7 + [_sz]),
    "pa_stage2_prove_dev": (ctypes.c_int, [_ctx] + [_vp] * 7 + [_sz]),
    "pa_stage2_verify": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_powfcom_prove_w": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_powfcom_prove_w_dev": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_stage1_prove_w": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_stage1_prove_w_dev": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_stage2_prove_w": (ctypes.c_int, [_ctx] + [_vp] * 8 + [_sz]),
    "pa_stage2_prove_w_dev": (ctypes.c_int, [_ctx] + [_vp] * 8 + [_sz]),
    "pa_stage2_verify_dev": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_commit_points": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_commit_points_dev": (ctypes.c_int, [_ctx] + [_vp] * 4 + [_sz]),
    "pa_y_scan": (ctypes.c_int, [_ctx, _vp, _vp, _sz]),
    "pa_y_scan_batch": (ctypes.c_int, [_ctx, _vp, _vp, _vp, _sz]),
    "pa_y_scan_dev": (ctypes.c_int, [_ctx, _vp, _vp, _vp, _sz, _sz]),
    "pa_point_sum_is_inf": (ctypes.c_int, [_ctx, _vp, _sz, ctypes.POINTER(ctypes.c_int)]),
    "pa_point_sum_is_inf_batch": (ctypes.c_int, [_ctx, _vp, _vp, _sz, _vp]),
    "pa_point_sum_is_inf_dev": (ctypes.c_int, [_ctx, _vp, _vp, _sz, _sz, _vp]),
    "pa_rng_fill": (ctypes.c_int, [_ctx, ctypes.c_uint64, _vp, _vp, _sz, _vp, _sz]),
    "pa_rng_fill_dev": (ctypes.c_int, [_ctx, ctypes.c_uint64, _vp, _vp, _sz, _vp, _sz]),
    "pa_seal_run": (ctypes.c_int, [_ctx, _vp]),
    "pa_ccs22_run": (ctypes.c_int, [_ctx, _vp]),
    "pa_ccs22_commit": (ctypes.c_int, [_ctx, _vp, _sz] + [_vp] * 5 + [_sz]),
    "pa_ccs22_commit_dev": (ctypes.c_int, [_ctx, _vp, _sz] + [_vp] * 5 + [_sz]),
    "pa_ccs22_bes_encode": (ctypes.c_int, [_ctx, _vp, _sz] + [_vp] * 5 + [_sz]),
    "pa_ccs22_bes_encode_dev": (ctypes.c_int, [_ctx, _vp, _sz] + [_vp] * 5 + [_sz]),
    "pa_ccs22_ot_recv2": (ctypes.c_int, [_ctx, _vp, _vp, _vp, _sz, ctypes.POINTER(ctypes.c_int)]),
    "pa_ccs22_ot_recv2_dev": (ctypes.c_int, [_ctx, _vp, _vp, _vp, _sz, _vp]),
    "pa_ccs22_ot_recv1": (ctypes.c_int, [_ctx] + [_vp] * 5 + [_sz]),
    "pa_ccs22_ot_recv1_dev": (ctypes.c_int, [_ctx] + [_vp] * 5 + [_sz]),
    "pa_ccs22_ot_send": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_ccs22_ot_send_dev": (ctypes.c_int, [_ctx] + [_vp] * 6 + [_sz]),
    "pa_rng_fill256": (ctypes.c_int, [_ctx, ctypes.c_uint64, _vp, _vp, _sz, _vp, _sz]),
    "pa_rng_fill256_dev": (ctypes.c_int, [_ctx, ctypes.c_uint64, _vp, _vp, _sz, _vp, _sz]),
    "pa_ccs22_setup_hash": (ctypes.c_int, [_ctx, _vp, _sz, _vp, _sz]),
    "pa_ccs22_setup_hash_dev": (ctypes.c_int, [_ctx, _vp, _sz, _vp, _sz]),
    "pa_ctx_set_entropy": (ctypes.c_int, [_ctx, _vp]),
    "pa_debug_set": (ctypes.c_int, [_ctx, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]),
    "pa_ctx_reruns": (ctypes.c_uint64, [_ctx]),
    "pa_xchg_create": (ctypes.c_int, [_ctx, _vp]),
    "pa_xchg_connect": (ctypes.c_int, [_ctx, _vp, ctypes.c_int, ctypes.c_int]),
    "pa_xchg_skip": (ctypes.c_int, [_ctx]),
    "pa_xchg_close": (ctypes.c_int, [_ctx]),
    "pa_scalar_mul_jobs": (ctypes.c_int, [_ctx, _vp, _sz]),
    "pa_profile_begin": (ctypes.c_int, [_ctx]),
    "pa_profile_end": (ctypes.c_int, [_ctx, ctypes.c_void_p, _sz, ctypes.POINTER(_sz)]),
}


class KernelStat(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("launches", ctypes.c_uint64), ("total_ms", ctypes.c_double)]


ALLGATHER_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int)


class MulJob(ctypes.Structure):
    """pa_mul_job (include/pa_engine.h)"""
    _fields_ = [("kind", ctypes.c_int), ("a", ctypes.c_void_p), ("p", ctypes.c_void_p), ("b", ctypes.c_void_p), ("q", ctypes.c_void_p),
                ("out", ctypes.c_void_p), ("n", ctypes.c_size_t)]


MUL_FIXED, MUL_VAR, MUL_DOUBLE, MUL_LINCOMB2 = 0, 1, 2, 3


class SealJob(ctypes.Structure):
    """pa_seal_job of include/pa_engine.h"""
    _fields_ = [
        ("seed", ctypes.c_uint64), ("n_auctions", ctypes.c_size_t), ("n", ctypes.c_void_p), ("c", ctypes.c_void_p),
        ("auction_ids", ctypes.c_void_p), ("bids", ctypes.c_void_p), ("verify", ctypes.c_int),
        ("lo", ctypes.c_uint32), ("hi", ctypes.c_uint32), ("slice", ctypes.c_uint32),
        ("allgather", ALLGATHER_FN), ("user", ctypes.c_void_p), ("d_send", ctypes.c_void_p), ("d_recv", ctypes.c_void_p),
        ("max_bid", ctypes.c_void_p), ("ok", ctypes.c_void_p),
        ("out_commit", ctypes.c_void_p), ("out_commit_ok", ctypes.c_void_p), ("out_r1", ctypes.c_void_p),
        ("out_r1_ok", ctypes.c_void_p), ("out_r2_tag", ctypes.c_void_p), ("out_r2_b", ctypes.c_void_p),
        ("out_r2_proof", ctypes.c_void_p), ("out_r2_ok", ctypes.c_void_p), ("out_r3", ctypes.c_void_p),
        ("schedule", ctypes.c_int), ("xchg_bytes", ctypes.c_size_t),
        ("use_xchg", ctypes.c_int), ("ok_all", ctypes.c_void_p),
    ]


class Ccs22Job(ctypes.Structure):
    """pa_ccs22_job of include/pa_engine.h"""
    _fields_ = [
        ("seed", ctypes.c_uint64), ("n_auctions", ctypes.c_size_t), ("n", ctypes.c_void_p), ("c", ctypes.c_void_p),
        ("evaluator", ctypes.c_void_p), ("auction_ids", ctypes.c_void_p), ("bids", ctypes.c_void_p), ("max_bid", ctypes.c_void_p),
        ("out_params", ctypes.c_void_p), ("out_com", ctypes.c_void_p), ("out_pub", ctypes.c_void_p), ("out_r1", ctypes.c_void_p),
        ("out_ots", ctypes.c_void_p), ("out_d", ctypes.c_void_p), ("schedule", ctypes.c_int),
    ]


class EngineError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    """dlopen libpa_engine.so and attach the prototypes.  Works without a GPU
    (the CUDA runtime is only touched by pa_ctx_create)."""
    if not os.path.exists(path):
        raise EngineError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no fallback implementation"
        )
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    return lib


def _buf(b):
    """bytes-like -> (ctypes pointer, keepalive)"""
    if isinstance(b, (bytes, bytearray, memoryview)):
        arr = (ctypes.c_uint8 * len(b)).from_buffer_copy(bytes(b)) if not isinstance(b, bytearray) else (ctypes.c_uint8 * len(b)).from_buffer(b)
        return ctypes.cast(arr, ctypes.c_void_p), arr
    if hasattr(b, "ctypes"):  # numpy array
        return ctypes.c_void_p(b.ctypes.data), b
    raise TypeError(type(b))


class Engine:
    """One engine context on one GPU (pa_ctx)."""

    def __init__(self, device=0, lib=None, entropy=None):
        """entropy: None = the seeded draw stream (tests, benchmarks: reproducible); "os" = a fresh 32-byte key
        from the operating system mixed into every draw (what a deployment uses); or 32 bytes."""
        self.lib = lib or load_library()
        self.ctx = _ctx()
        self.xchg_world = 0
        rc = self.lib.pa_ctx_create(ctypes.byref(self.ctx), device)
        if rc != PA_OK:
            msg = self.lib.pa_last_error(None)
            self.ctx = None
            raise EngineError(f"pa_ctx_create failed ({rc}): {msg.decode() if msg else ''}")
        if entropy is not None:
            self.set_entropy(os.urandom(32) if entropy == "os" else entropy)

    def set_entropy(self, key32):
        """pa_ctx_set_entropy: 32 secret bytes mixed into every draw; None = back to the seeded test stream."""
        if key32 is None:
            self._check(self.lib.pa_ctx_set_entropy(self.ctx, None))
            return
        assert len(key32) == 32
        p, keep = _buf(bytes(key32))
        self._check(self.lib.pa_ctx_set_entropy(self.ctx, p))

    def debug_set(self, what, a=0, b=0, c=0):
        """pa_debug_set (test hooks): what = 1 reject bits (a), 2 corrupt (a = section, b = step << 32 | bidder, c = offset)."""
        self._check(self.lib.pa_debug_set(self.ctx, what, a, b, c))

    @property
    def reruns(self):
        return self.lib.pa_ctx_reruns(self.ctx)

    # -- peer exchange window (one process per GPU) -----------------------------------
    def xchg_create(self):
        """pa_xchg_create: allocate this rank's window, return its 64-byte IPC handle."""
        h = (ctypes.c_uint8 * 64)()
        self._check(self.lib.pa_xchg_create(self.ctx, ctypes.addressof(h)))
        return bytes(h)

    def xchg_connect(self, handles, rank):
        """pa_xchg_connect: handles = every rank's handle in rank order."""
        blob = b"".join(handles)
        p, keep = _buf(blob)
        self._check(self.lib.pa_xchg_connect(self.ctx, p, len(handles), rank))
        self.xchg_world, self.xchg_rank = len(handles), rank

    def xchg_skip(self):
        self._check(self.lib.pa_xchg_skip(self.ctx))

    def xchg_close(self):
        """pa_xchg_close: unmap the peers' windows and free this rank's."""
        self._check(self.lib.pa_xchg_close(self.ctx))
        self.xchg_world = 0

    def close(self):
        if self.ctx:
            self.lib.pa_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != PA_OK:
            msg = self.lib.pa_last_error(self.ctx)
            raise EngineError(f"engine call failed ({rc}): {msg.decode() if msg else ''}")

    # -- raw helpers -------------------------------------------------------------
    @property
    def stream(self):
        return self.lib.pa_ctx_stream(self.ctx)

    @property
    def launches(self):
        return self.lib.pa_ctx_launches(self.ctx)

    def sync(self):
        self._check(self.lib.pa_sync(self.ctx))

    def dev_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._check(self.lib.pa_dev_alloc(self.ctx, ctypes.byref(p), nbytes))
        return p.value

    def dev_free(self, dptr):
        self._check(self.lib.pa_dev_free(self.ctx, dptr))

    def upload(self, dptr, data):
        p, keep = _buf(data)
        self._check(self.lib.pa_dev_upload(self.ctx, dptr, p, len(data) if not hasattr(data, "nbytes") else data.nbytes))
        self.sync()
        return keep

    def download(self, dptr, nbytes):
        out = bytearray(nbytes)
        p, keep = _buf(out)
        self._check(self.lib.pa_dev_download(self.ctx, p, dptr, nbytes))
        return bytes(out)

    # -- host-buffer entry points -----------------------------------------------------
    def fixed_base_mul(self, scalars):
        n = len(scalars) // 32
        out = bytearray(64 * n)
        ps, k1 = _buf(scalars)
        po, k2 = _buf(out)
        self._check(self.lib.pa_fixed_base_mul(self.ctx, ps, po, n))
        return bytes(out)

    def var_base_mul(self, points, scalars):
        n = len(scalars) // 32
        out = bytearray(64 * n)
        pp, k0 = _buf(points)
        ps, k1 = _buf(scalars)
        po, k2 = _buf(out)
        self._check(self.lib.pa_var_base_mul(self.ctx, pp, ps, po, n))
        return bytes(out)

    def double_mul(self, a, points, b):
        n = len(a) // 32
        out = bytearray(64 * n)
        pa_, k0 = _buf(a)
        pp, k1 = _buf(points)
        pb, k2 = _buf(b)
        po, k3 = _buf(out)
        self._check(self.lib.pa_double_mul(self.ctx, pa_, pp, pb, po, n))
        return bytes(out)

    def lincomb2(self, p, a, q, b):
        n = len(a) // 32
        out = bytearray(64 * n)
        bufs = [_buf(x) for x in (p, a, q, b, out)]
        self._check(self.lib.pa_lincomb2(self.ctx, *[x[0] for x in bufs], n))
        return bytes(out)

    def point_add(self, p, q, sub=False):
        n = len(p) // 64
        out = bytearray(64 * n)
        bufs = [_buf(x) for x in (p, q, out)]
        self._check(self.lib.pa_point_add(self.ctx, *[x[0] for x in bufs], n, 1 if sub else 0))
        return bytes(out)

    def point_on_curve(self, points):
        n = len(points) // 64
        out = bytearray(n)
        pp, k0 = _buf(points)
        po, k1 = _buf(out)
        self._check(self.lib.pa_point_on_curve(self.ctx, pp, n, po))
        return bytes(out)

    def point_encode(self, points, compressed=False, stride=None):
        n = len(points) // 64
        stride = stride or (33 if compressed else 65)
        out = bytearray(stride * n)
        lens = (ctypes.c_uint32 * max(n, 1))()
        pp, k0 = _buf(points)
        po, k1 = _buf(out)
        self._check(self.lib.pa_point_encode(self.ctx, pp, n, 1 if compressed else 0, po, stride, ctypes.cast(lens, ctypes.c_void_p)))
        return [bytes(out[i * stride:i * stride + lens[i]]) for i in range(n)]

    def profile_begin(self):
        self._check(self.lib.pa_profile_begin(self.ctx))

    def profile_end(self):
        arr = (KernelStat * 64)()
        cnt = _sz(0)
        self._check(self.lib.pa_profile_end(self.ctx, ctypes.cast(arr, ctypes.c_void_p), 64, ctypes.byref(cnt)))
        return {arr[i].name.decode(): {"launches": int(arr[i].launches), "total_ms": float(arr[i].total_ms)}
                for i in range(min(cnt.value, 64))}

    # device-pointer entry points (asynchronous on the context's stream)
    def fixed_base_mul_dev(self, d_scalars, d_out, n):
        self._check(self.lib.pa_fixed_base_mul_dev(self.ctx, d_scalars, d_out, n))

    def var_base_mul_dev(self, d_points, d_scalars, d_out, n):
        self._check(self.lib.pa_var_base_mul_dev(self.ctx, d_points, d_scalars, d_out, n))

    def double_mul_dev(self, d_a, d_points, d_b, d_out, n):
        self._check(self.lib.pa_double_mul_dev(self.ctx, d_a, d_points, d_b, d_out, n))

    def lincomb2_dev(self, d_p, d_a, d_q, d_b, d_out, n):
        self._check(self.lib.pa_lincomb2_dev(self.ctx, d_p, d_a, d_q, d_b, d_out, n))

    # -- proofs and round logic (host buffers) -----------------------------------------------
    def _run(self, fn, ins, out_bytes, n, extra=()):
        out = bytearray(out_bytes)
        bufs = [_buf(x) for x in ins] + [_buf(out)]
        self._check(getattr(self.lib, fn)(self.ctx, *[b[0] for b in bufs], *extra, n))
        return bytes(out)

    @staticmethod
    def _ids(ids):
        return bytes((ctypes.c_uint64 * max(len(ids), 1))(*ids))[:8 * len(ids)]

    def challenge(self, points, k, ids):
        n = len(ids)
        out = bytearray(32 * n)
        b = [_buf(points), _buf(self._ids(ids)), _buf(out)]
        self._check(self.lib.pa_challenge(self.ctx, b[0][0], k, b[1][0], b[2][0], n))
        return bytes(out)

    def pokdlog_prove(self, X, x, ids, rnd):
        return self._run("pa_pokdlog_prove", (X, x, self._ids(ids), rnd), 96 * len(ids), len(ids))

    def pokdlog_verify(self, proofs, X, ids):
        return self._run("pa_pokdlog_verify", (proofs, X, self._ids(ids)), len(ids), len(ids))

    def powfcom_prove(self, stmt, alpha, bits, ids, rnd):
        return self._run("pa_powfcom_prove", (stmt, alpha, bytes(bits), self._ids(ids), rnd), 352 * len(ids), len(ids))

    def powfcom_verify(self, proofs, stmt, ids):
        return self._run("pa_powfcom_verify", (proofs, stmt, self._ids(ids)), len(ids), len(ids))

    def stage1_prove(self, stmt, secrets, bits, ids, rnd):
        return self._run("pa_stage1_prove", (stmt, secrets, bytes(bits), self._ids(ids), rnd), 672 * len(ids), len(ids))

    def stage1_verify(self, proofs, stmt, ids):
        return self._run("pa_stage1_verify", (proofs, stmt, self._ids(ids)), len(ids), len(ids))

    def stage2_prove(self, stmt, secrets, bi, bj, ids, rnd):
        return self._run("pa_stage2_prove", (stmt, secrets, bytes(bi), bytes(bj), self._ids(ids), rnd), 1344 * len(ids), len(ids))

    def stage2_verify(self, proofs, stmt, ids):
        return self._run("pa_stage2_verify", (proofs, stmt, self._ids(ids)), len(ids), len(ids))

    # provers with witnesses (include/pa_engine.h): extended secrets, byte-identical proofs, 2-3 times less work
    def powfcom_prove_w(self, stmt, secrets, bits, ids, rnd):
        return self._run("pa_powfcom_prove_w", (stmt, secrets, bytes(bits), self._ids(ids), rnd), 352 * len(ids), len(ids))

    def stage1_prove_w(self, stmt, secrets, bits, ids, rnd):
        return self._run("pa_stage1_prove_w", (stmt, secrets, bytes(bits), self._ids(ids), rnd), 672 * len(ids), len(ids))

    def stage2_prove_w(self, stmt, secrets, bi, bj, cbit, ids, rnd):
        return self._run("pa_stage2_prove_w", (stmt, secrets, bytes(bi), bytes(bj), bytes(cbit), self._ids(ids), rnd), 1344 * len(ids), len(ids))

    def scalar_mul_jobs(self, jobs):
        """pa_scalar_mul_jobs: jobs = [(kind, dict(a=, p=, b=, q=), n)] with host byte strings; returns the outputs in order.
        Same results as the single calls, one interleaved copy/compute pipeline."""
        arr = (MulJob * len(jobs))()
        keep, outs = [], []
        for k, (kind, ops, n) in enumerate(jobs):
            arr[k].kind, arr[k].n = kind, n
            for name in ("a", "p", "b", "q"):
                if ops.get(name) is not None:
                    ptr, ref = _buf(ops[name])
                    keep.append(ref)
                    setattr(arr[k], name, ptr.value)
            out = bytearray(64 * n)
            ptr, ref = _buf(out)
            keep.append(ref)
            arr[k].out = ptr.value
            outs.append(out)
        self._check(self.lib.pa_scalar_mul_jobs(self.ctx, ctypes.addressof(arr), len(jobs)))
        return [bytes(o) for o in outs]

    def commit_points(self, alpha, beta, bits):
        return self._run("pa_commit_points", (alpha, beta, bytes(bits)), 192 * len(bits), len(bits))

    def y_scan(self, X):
        n = len(X) // 64
        if n == 0:
            return b""
        return self._run("pa_y_scan", (X,), 64 * n, n)

    def y_scan_batch(self, X, offsets):
        import struct
        nseg = len(offsets) - 1
        out = bytearray(len(X))
        b = [_buf(X), _buf(out), _buf(struct.pack(f"<{len(offsets)}I", *offsets))]
        self._check(self.lib.pa_y_scan_batch(self.ctx, b[0][0], b[1][0], b[2][0], nseg))
        return bytes(out)

    def point_sum_is_inf(self, B):
        flag = ctypes.c_int(0)
        p, keep = _buf(B) if len(B) else (None, None)
        self._check(self.lib.pa_point_sum_is_inf(self.ctx, p, len(B) // 64, ctypes.byref(flag)))
        return bool(flag.value)

    def point_sum_is_inf_batch(self, B, offsets):
        import struct
        nseg = len(offsets) - 1
        flags = bytearray(4 * nseg)
        b = [_buf(B), _buf(struct.pack(f"<{len(offsets)}I", *offsets)), _buf(flags)]
        self._check(self.lib.pa_point_sum_is_inf_batch(self.ctx, b[0][0], b[1][0], nseg, b[2][0]))
        return [bool(v) for v in struct.unpack(f"<{nseg}i", bytes(flags))]

    def rng_fill(self, seed, streams, counters, per_item):
        """returns (draws bytes, advanced counters list)"""
        import struct
        n = len(streams)
        out = bytearray(32 * per_item * n)
        ctr = bytearray(struct.pack(f"<{n}Q", *counters))
        b = [_buf(struct.pack(f"<{n}Q", *streams)), _buf(ctr), _buf(out)]
        self._check(self.lib.pa_rng_fill(self.ctx, seed, b[0][0], b[1][0], per_item, b[2][0], n))
        return bytes(out), list(struct.unpack(f"<{n}Q", bytes(ctr)))

    def seal_run(self, seed, n, c, bids, verify=True, sections=False, auction_ids=None, shard=None, schedule=0):
        """pa_seal_run: whole SEAL auctions, device resident.
        n, c: per-auction lists; bids: flat list of the LOCAL bidders' bids (auction-major).
        shard = dict(lo, hi, slice, d_send, d_recv, allgather=callable(which) -> 0) for one
        auction sharded by bidder slice.  Returns dict(max_bid, ok[, sections...])."""
        A = len(n)
        m = len(bids)
        n_arr = (ctypes.c_uint32 * A)(*n)
        c_arr = (ctypes.c_uint32 * A)(*c)
        bid_arr = (ctypes.c_uint64 * max(m, 1))(*bids)
        max_bid = (ctypes.c_uint64 * A)()
        ok = (ctypes.c_uint8 * A)()
        job = SealJob()
        job.seed, job.n_auctions = seed, A
        job.n, job.c, job.bids = ctypes.addressof(n_arr), ctypes.addressof(c_arr), ctypes.addressof(bid_arr)
        keep = [n_arr, c_arr, bid_arr, max_bid, ok]
        if auction_ids is not None:
            aid = (ctypes.c_uint64 * A)(*auction_ids)
            job.auction_ids = ctypes.addressof(aid)
            keep.append(aid)
        job.verify = int(verify)   # False/True, or k > 1: every proof verified k times (k = n - 1: the reference's all-pairs work)
        job.schedule = schedule   # 0 auto, 1 step-major, 2 phase-major (one unsharded auction)
        job.max_bid, job.ok = ctypes.addressof(max_bid), ctypes.addressof(ok)
        cb = None
        ok_all = ctypes.c_uint8(1)
        if shard is not None:
            job.lo, job.hi, job.slice = shard["lo"], shard["hi"], shard["slice"]
            if shard.get("use_xchg"):   # kernels exchange through the peer window: no callback, no buffers
                job.use_xchg = 1
                job.ok_all = ctypes.addressof(ok_all)
            else:
                fn = shard["allgather"]
                cb = ALLGATHER_FN(lambda user, which: int(fn(which) or 0))
                job.allgather = cb
                job.d_send, job.d_recv = shard["d_send"], shard["d_recv"]
                job.xchg_bytes = shard.get("xchg_bytes", 0)
        out = {}
        if sections:
            cmax = max(c)
            if shard is None:
                per_bidder_c = [c[a] for a in range(A) for _ in range(n[a])]
            else:
                per_bidder_c = [c[0]] * m
            Mb = sum(per_bidder_c)
            sizes = {"out_commit": Mb * 736, "out_commit_ok": Mb, "out_r1": cmax * m * 320, "out_r1_ok": cmax * m,
                     "out_r2_tag": cmax * m, "out_r2_b": cmax * m * 64, "out_r2_proof": cmax * m * 1344,
                     "out_r2_ok": cmax * m, "out_r3": cmax * A}
            for name, sz in sizes.items():
                buf = (ctypes.c_uint8 * max(sz, 1))()
                setattr(job, name, ctypes.addressof(buf))
                out[name] = buf
        self._check(self.lib.pa_seal_run(self.ctx, ctypes.byref(job)))
        res = {"max_bid": list(max_bid), "ok": [bool(v) for v in ok]}
        if shard is not None and shard.get("use_xchg"):
            res["ok_all"] = bool(ok_all.value)
        for name, buf in out.items():
            res[name[4:]] = bytes(buf)
        return res

    def ccs22_run(self, seed, n, c, evaluator, bids, sections=False, auction_ids=None, schedule=0):
        """pa_ccs22_run: whole CCS22 auctions, device resident.  Returns dict(max_bid[, sections...])."""
        A, m = len(n), len(bids)
        arr = lambda T, v: (T * max(len(v), 1))(*v)
        n_a, c_a, e_a, b_a = arr(ctypes.c_uint32, n), arr(ctypes.c_uint32, c), arr(ctypes.c_uint32, evaluator), arr(ctypes.c_uint64, bids)
        mb = (ctypes.c_uint64 * max(m, 1))()
        job = Ccs22Job()
        job.seed, job.n_auctions = seed, A
        job.n, job.c, job.evaluator, job.bids = map(ctypes.addressof, (n_a, c_a, e_a, b_a))
        job.max_bid = ctypes.addressof(mb)
        job.schedule = schedule
        keep = [n_a, c_a, e_a, b_a, mb]
        if auction_ids is not None:
            aid = arr(ctypes.c_uint64, auction_ids)
            job.auction_ids = ctypes.addressof(aid)
            keep.append(aid)
        out = {}
        if sections:
            cmax = max(c)
            Mb = sum(n[a] * c[a] for a in range(A))
            ms = sum(n[a] - 1 for a in range(A))
            sizes = {"out_params": A * 128, "out_com": m * 64, "out_pub": Mb * 64, "out_r1": cmax * ms * 192,
                     "out_ots": cmax * ms * 192, "out_d": cmax * A}
            for name, sz in sizes.items():
                buf = (ctypes.c_uint8 * max(sz, 1))()
                setattr(job, name, ctypes.addressof(buf))
                out[name] = buf
        self._check(self.lib.pa_ccs22_run(self.ctx, ctypes.byref(job)))
        res = {"max_bid": list(mb)[:m]}
        for name, buf in out.items():
            res[name[4:]] = bytes(buf)
        return res

    def ccs22_commit(self, scalars, k, bid, R, params):
        n = len(bid) // 32
        H, com = bytearray(32 * n), bytearray(64 * n)
        b = [_buf(x) for x in (scalars, bid, R, params, H, com)]
        self._check(self.lib.pa_ccs22_commit(self.ctx, b[0][0], k, b[1][0], b[2][0], b[3][0], b[4][0], b[5][0], n))
        return bytes(H), bytes(com)

    def ccs22_bes_encode(self, X, ids, d, x, r):
        n, m = len(X) // 64, len(ids)
        ida = (ctypes.c_uint64 * max(m, 1))(*ids)
        out = bytearray(64 * m)
        b = [_buf(v) for v in (X, bytes(d), x, r, out)]
        self._check(self.lib.pa_ccs22_bes_encode(self.ctx, b[0][0], n, ida, b[1][0], b[2][0], b[3][0], b[4][0], m))
        return bytes(out)

    def ccs22_ot_recv2(self, ots, beta, B):
        n = len(beta) // 32
        flag = ctypes.c_int(0)
        b = [_buf(v) for v in (ots, beta, B)]
        self._check(self.lib.pa_ccs22_ot_recv2(self.ctx, b[0][0], b[1][0], b[2][0], n, ctypes.byref(flag)))
        return bool(flag.value)

    def ccs22_ot_recv1(self, k, beta, alpha, params):
        n = len(k) // 32
        return self._run("pa_ccs22_ot_recv1", (k, beta, alpha, params), 192 * n, n)

    def ccs22_ot_send(self, r1, params, B, st, m):
        n = len(m) // 32
        return self._run("pa_ccs22_ot_send", (r1, params, B, st, m), 192 * n, n)

    def ccs22_setup_hash(self, scalars, k):
        n = len(scalars) // (32 * k)
        out = bytearray(32 * n)
        b = [_buf(scalars), _buf(out)]
        self._check(self.lib.pa_ccs22_setup_hash(self.ctx, b[0][0], k, b[1][0], n))
        return bytes(out)

    def measure_int_peak(self):
        out = (ctypes.c_double * 6)()
        self._check(self.lib.pa_measure_int_peak(self.ctx, out))
        return {"imad_per_s": out[0], "imad_wide_per_s": out[1], "fe_mul_per_s": out[2], "fe_sqr_per_s": out[3],
                "imad_wide_unchained_per_s": out[4]}
