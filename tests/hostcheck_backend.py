"""TEST INFRASTRUCTURE — a SealFlow backend whose proofs are produced and checked by
the device proof code (pa_proof.cuh) compiled for the host (tests/hostcheck), one
proof at a time; everything else is delegated to the libcrypto oracle.  Lets the
proof tables be checked against the reference's transcripts without a GPU."""
import ctypes
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HC = os.path.join(ROOT, "tests", "hostcheck")
KIND = {"pokdlog": (0, 96, 1, 1, 1, 1), "powfcom": (1, 352, 3, 1, 3, 4), "stage1": (2, 672, 7, 2, 5, 8), "stage2": (3, 1344, 11, 3, 11, 16)}
# name -> (kind id, record bytes, statement points, secrets, draws, checks)


def build():
    so = os.path.join(HC, "libhostcheck.so")
    src = os.path.join(HC, "hostcheck.cpp")
    hdrs = [os.path.join(ROOT, "privacy-auction_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "privacy-auction_b200", "csrc")) if f.endswith(".cuh")]
    newest = max(os.path.getmtime(p) for p in [src] + hdrs)
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.run(["g++", "-O2", "-x", "c++", "-std=c++17", "-fPIC", "-shared",
                        "-I" + os.path.join(ROOT, "privacy-auction_b200", "csrc"), "-o", so, src], check=True)
    return ctypes.CDLL(so)


class HostcheckBackend:
    def __init__(self, oracle):
        self.o = oracle
        self.hc = build()
        self.hc.hc_proof_verify.restype = ctypes.c_uint
        for name in ("fixed_base_mul", "var_base_mul", "commit_points", "y_scan", "point_sum_is_inf"):
            setattr(self, name, getattr(oracle, name))

    def _prove(self, name, stmt, secrets, branches, ids, rnd):
        kind, rec, nst, nsec, nrnd, _ = KIND[name]
        out = bytearray()
        for i, ident in enumerate(ids):
            proof = ctypes.create_string_buffer(rec)
            self.hc.hc_proof_prove(kind, proof, stmt[64 * nst * i:64 * nst * (i + 1)], ctypes.c_ulonglong(ident),
                                   secrets[32 * nsec * i:32 * nsec * (i + 1)], rnd[32 * nrnd * i:32 * nrnd * (i + 1)],
                                   int(branches[i]))
            out += proof.raw
        return bytes(out)

    def _verify(self, name, proofs, stmt, ids):
        kind, rec, nst, _, _, nchk = KIND[name]
        out = bytearray()
        for i, ident in enumerate(ids):
            m = self.hc.hc_proof_verify(kind, proofs[rec * i:rec * (i + 1)], stmt[64 * nst * i:64 * nst * (i + 1)], ctypes.c_ulonglong(ident))
            out.append(1 if m == (1 << nchk) - 1 else 0)
        return bytes(out)

    def pokdlog_prove(self, X, x, ids, rnd):
        return self._prove("pokdlog", X, x, [0] * len(ids), ids, rnd)

    def pokdlog_verify(self, proofs, X, ids):
        return self._verify("pokdlog", proofs, X, ids)

    def powfcom_prove(self, stmt, alpha, bits, ids, rnd):
        return self._prove("powfcom", stmt, alpha, bits, ids, rnd)

    def powfcom_verify(self, proofs, stmt, ids):
        return self._verify("powfcom", proofs, stmt, ids)

    def stage1_prove(self, stmt, secrets, bits, ids, rnd):
        return self._prove("stage1", stmt, secrets, bits, ids, rnd)

    def stage1_verify(self, proofs, stmt, ids):
        return self._verify("stage1", proofs, stmt, ids)

    def stage2_prove(self, stmt, secrets, bi, bj, ids, rnd):
        br = [0 if a == 1 else (1 if b == 1 else 2) for a, b in zip(bi, bj)]
        return self._prove("stage2", stmt, secrets, br, ids, rnd)

    def stage2_verify(self, proofs, stmt, ids):
        return self._verify("stage2", proofs, stmt, ids)

    # ---- provers with witnesses (pa_engine.h, pa_*_prove_w): extended secrets, same bytes --------------
    def _prove_w(self, name, nsec, stmt, secrets, branches, veto_i, veto_j, cbit, ids, rnd):
        kind, rec, nst, _, nrnd, _ = KIND[name]
        out = bytearray()
        for i, ident in enumerate(ids):
            proof = ctypes.create_string_buffer(rec)
            self.hc.hc_proof_prove_w(kind, proof, stmt[64 * nst * i:64 * nst * (i + 1)], ctypes.c_ulonglong(ident),
                                     secrets[32 * nsec * i:32 * nsec * (i + 1)], rnd[32 * nrnd * i:32 * nrnd * (i + 1)],
                                     int(branches[i]), int(veto_i[i]), int(veto_j[i]), int(cbit[i]))
            out += proof.raw
        return bytes(out)

    def powfcom_prove_w(self, stmt, secrets, bits, ids, rnd):
        return self._prove_w("powfcom", 2, stmt, secrets, bits, bits, [0] * len(ids), bits, ids, rnd)

    def stage1_prove_w(self, stmt, secrets, bits, ids, rnd):
        return self._prove_w("stage1", 4, stmt, secrets, bits, bits, [0] * len(ids), bits, ids, rnd)

    def stage2_prove_w(self, stmt, secrets, bi, bj, cbit, ids, rnd):
        br = [0 if a == 1 else (1 if b == 1 else 2) for a, b in zip(bi, bj)]
        return self._prove_w("stage2", 6, stmt, secrets, br, bi, bj, cbit, ids, rnd)
