"""Host-side orchestration of one CCS22 auction over a batched backend, in the order of the
reference's main (CCS22/main.cpp:16-130) and party logic (CCS22/bidder.cpp:48-212,
CCS22/evaluator.cpp:22-156).  Backend = the CUDA engine or the libcrypto oracle port (same
method names).  Output: the PACCS22T transcript that oracle/ref_ccs22_driver.cpp writes for the
unmodified reference.

Randomness: party i draws from PA stream (seed, i); the bulletin board from (seed, 0xFFFFFFFF).
"""
import struct

import secp256k1_py as E

G64 = E.enc64(E.G)


def b32(x):
    return int(x).to_bytes(32, "big")


class Ccs22Flow:
    def __init__(self, backend, n, c, seed, evaluator_id, bids, auction=0):
        self.be, self.n, self.c, self.seed, self.e, self.bids = backend, n, c, seed, evaluator_id, list(bids)
        self.st = [E.PaStream(seed, (auction << 32) | i) for i in range(n)]
        self.bb = E.PaStream(seed, (auction << 32) | 0xFFFFFFFF)
        self.bits = [[(bid >> (c - 1 - i)) & 1 for i in range(c)] for bid in bids]
        self.in_race = [True] * n
        self.max_bid = [0] * n
        self.others = [i for i in range(n) if i != evaluator_id]     # pos(i) order, CCS22/main.cpp:38-41
        self.out = bytearray()

    def run(self):
        be, n, c, ev = self.be, self.n, self.c, self.e
        out = self.out
        out += b"PACCS22T" + struct.pack("<QQQQ", n, c, self.seed, ev)
        for bid in self.bids:
            out += struct.pack("<Q", bid)
        # public parameters: g1 = g^rand256, h = g^rand256 (unreduced draws)     CCS22/bulletinBoard.cpp:28-51
        k1, k2 = self.bb.rand256(), self.bb.rand256()
        gh = be.fixed_base_mul(b32(k1) + b32(k2))
        g1, h = gh[:64], gh[64:]
        out += g1 + h
        # party constructors draw R                                            CCS22/bidder.cpp:21-27
        R = [self.st[i].rand_range() for i in range(n)]
        # ---- setup (CCS22/bidder.cpp:48-89, evaluator.cpp:22-63) --------------------------------
        x, r, s, t, beta = {}, {}, {}, {}, []
        for i in range(n):
            if i == ev:
                x[i], r[i] = [], []
                for _ in range(c):
                    x[i].append(self.st[i].rand_range()); r[i].append(self.st[i].rand_range())
                    beta.append([self.st[i].rand_range() for _ in range(n - 1)])   # beta[step][j]
            else:
                x[i], r[i], s[i], t[i] = [], [], [], []
                for _ in range(c):
                    x[i].append(self.st[i].rand_range()); r[i].append(self.st[i].rand_range())
                    s[i].append(self.st[i].rand_range()); t[i].append(self.st[i].rand_range())
        X = {i: be.fixed_base_mul(b"".join(map(b32, x[i]))) for i in range(n)}   # X_i = g^x_i  :67
        Com = {}
        for i in range(n):
            if i == ev:   # hash over [x.., r.., beta_{0,0} .. beta_{c-1,n-2}]      evaluator.cpp:43-50
                arr = x[i] + r[i] + [b for row in beta for b in row]
            else:         # hash over [x.., r.., s.., t..]                          bidder.cpp:74-77
                arr = x[i] + r[i] + s[i] + t[i]
            H = be.ccs22_setup_hash(b"".join(map(b32, arr)), len(arr))
            # Com = g^bid * g1^H + h^R                                              bidder.cpp:84-88
            t1 = be.double_mul(b32(self.bids[i]), g1, H)
            t2 = be.var_base_mul(h, b32(R[i]))
            Com[i] = be.point_add(t1, t2)
        for i in range(n):
            out += Com[i] + X[i]
        # ---- computation phase --------------------------------------------------------------
        nb = n - 1
        for step in range(c):
            Xs = b"".join(X[i][64 * step:64 * step + 64] for i in range(n))      # getPublicKeysByStep
            Y = be.y_scan(Xs)                                                    # every party uses its own Y only, bidder.cpp:124-136
            d, B = [], {}
            fb, vb = [], []
            for i in range(n):
                d.append(1 if (self.in_race[i] and self.bits[i][step] == 1) else 0)   # bidder.cpp:122
            # B = Y^x (d = 0) or g^r (d = 1)                                       bidder.cpp:138-142
            for i in range(n):
                if d[i] == 0:
                    B[i] = be.var_base_mul(Y[64 * i:64 * i + 64], b32(x[i][step]))
                else:
                    B[i] = be.fixed_base_mul(b32(r[i][step]))
            # OTReceive1 (evaluator.cpp:78-115): alpha = d_e; per j: k <- rand256, T2 = g^k, G = g^beta * g1^alpha, H = T2^alpha + h^beta
            alpha = b32(d[ev]) * nb
            ks = b"".join(b32(self.st[ev].rand256()) for _ in range(nb))
            betas = b"".join(map(b32, beta[step])) if nb else b""
            T2 = be.fixed_base_mul(ks)
            Gp = be.double_mul(betas, g1 * nb, alpha)
            Hp = be.lincomb2(T2, alpha, h * nb, betas)
            for j in range(nb):
                out += T2[64 * j:64 * j + 64] + Gp[64 * j:64 * j + 64] + Hp[64 * j:64 * j + 64]
            # OTSend (bidder.cpp:155-198): M1 = g^rand256; z = g^s * h^t; C0 = G^s + H^t + B; C1 = (G - g1)^s + (H - T2)^t + M1
            ms = b"".join(b32(self.st[i].rand256()) for i in self.others)
            ss = b"".join(b32(s[i][step]) for i in self.others)
            ts = b"".join(b32(t[i][step]) for i in self.others)
            Bs = b"".join(B[i] for i in self.others)
            M1 = be.fixed_base_mul(ms)
            z = be.double_mul(ss, h * nb, ts)
            C0 = be.point_add(be.lincomb2(Gp, ss, Hp, ts), Bs)
            Gm = be.point_add(Gp, g1 * nb, sub=True)
            Hm = be.point_add(Hp, T2, sub=True)
            C1 = be.point_add(be.lincomb2(Gm, ss, Hm, ts), M1)
            for j in range(nb):
                out += z[64 * j:64 * j + 64] + C0[64 * j:64 * j + 64] + C1[64 * j:64 * j + 64]
            # OTReceive2 (evaluator.cpp:117-156)
            if d[ev] == 1:
                newd = 1
            else:
                zb = be.var_base_mul(z, betas)                        # (z^-1)^beta = -(beta z)
                M0 = be.point_add(C0, zb, sub=True)                   # M0 = C0 * z^-beta
                newd = 0 if be.point_sum_is_inf(M0 + B[ev]) else 1
                if newd:
                    self.in_race[ev] = False
            out.append(newd)
            if newd:
                self.max_bid[ev] |= 1 << (c - step - 1)
                for i in self.others:                                 # checkIfEnterDeciderRound, bidder.cpp:200-212
                    if d[i] == 0:
                        self.in_race[i] = False
                    self.max_bid[i] |= 1 << (c - step - 1)
        for i in range(n):
            out += struct.pack("<Q", self.max_bid[i])
        return bytes(out)


def sections_to_transcripts(seed, n, c, evaluator, bids, res):
    """PACCS22T transcript of every auction from the section arrays pa_ccs22_run returns."""
    A = len(n)
    ms = sum(x - 1 for x in n)
    outs, p0, b0, s0 = [], 0, 0, 0
    for a in range(A):
        na, ca = n[a], c[a]
        out = bytearray(b"PACCS22T" + struct.pack("<QQQQ", na, ca, seed, evaluator[a]))
        for i in range(na):
            out += struct.pack("<Q", bids[p0 + i])
        out += res["params"][128 * a:128 * (a + 1)]
        for i in range(na):
            out += res["com"][64 * (p0 + i):64 * (p0 + i + 1)]
            out += res["pub"][64 * (b0 + i * ca):64 * (b0 + (i + 1) * ca)]
        for step in range(ca):
            base = step * ms + s0
            out += res["r1"][192 * base:192 * (base + na - 1)]
            out += res["ots"][192 * base:192 * (base + na - 1)]
            out.append(res["d"][step * A + a])
        for i in range(na):
            out += struct.pack("<Q", res["max_bid"][p0 + i])
        outs.append(bytes(out))
        p0 += na
        b0 += na * ca
        s0 += na - 1
    return outs
