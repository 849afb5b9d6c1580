"""Host-side orchestration of one SEAL auction over a batched backend, written to
read like the reference's own main (SEAL/main.cpp:32-120) and Bidder round logic
(SEAL/bidder.cpp:1109-1421).  The backend is anything exposing the engine's
batched operator interface (include/pa_engine.h): the CUDA engine, the libcrypto
oracle port, or the host-compiled device headers.  Output is the PASEALT1
transcript the Tier-A driver (oracle/ref_seal_driver.cpp) writes for the
unmodified reference, so parity is a byte comparison.

Randomness: bidder j draws from PA stream (seed, j) in the reference's draw
order.
"""
import struct

import secp256k1_py as E

N_ORD = E.N


def b32(x):
    return int(x).to_bytes(32, "big")


class SealFlow:
    """One SEAL auction.  `mine` (default: everybody) are the bidder ids this process plays;
    with a subset, `exchange(kind, local_bytes)` must return the concatenation over all
    processes in id order (the all-gather of SURVEY.md section 8e) — used for the X of
    round one and the b of round two, the only data a bidder needs from the others."""

    def __init__(self, backend, n, c, seed, bids, verify=True, auction=0, mine=None, exchange=None, witness=False):
        """witness: prove with the backend's pa_*_prove_w entry points (extended secrets: the prover's knowledge of the
        discrete logarithms of its own points turns variable-base into fixed-base work); same transcript."""
        assert len(bids) == n
        self.be, self.n, self.c, self.seed, self.bids, self.verify = backend, n, c, seed, list(bids), verify
        self.witness = witness
        self.mine = list(range(n)) if mine is None else list(mine)
        self.exchange = exchange or (lambda kind, data: data)
        self.streams = {j: E.PaStream(seed, (auction << 32) | j) for j in self.mine}
        # binaryBidStr: MSB first (SEAL/bidder.cpp:31, 1128)
        self.bits = {j: [(bids[j] >> (c - 1 - i)) & 1 for i in range(c)] for j in self.mine}
        self.junction = False            # junctionFlag, identical for all bidders
        self.prev_step = None            # prevDecidingStep
        self.prev_bit = {j: 1 for j in self.mine}   # prevDecidingBit (private, initialised to 1, SEAL/bidder.cpp:23)
        self.max_bid = 0
        # what gets published, keyed by bidder id
        self.sec = {"commit": {}, "commit_ok": {}, "r1": [], "r1_ok": [], "r2": [], "r2_ok": [], "r3": []}
        self.ok = True

    # -- helpers ---------------------------------------------------------------------------
    def _draw(self, j, k):
        return [self.streams[j].rand_range() for _ in range(k)]

    @staticmethod
    def _cat(items):
        return b"".join(items)

    # -- phases ------------------------------------------------------------------------------
    def commit(self):
        """Bidder::commitBid for every bidder and bit (SEAL/bidder.cpp:1109-1162)."""
        c, be, mine = self.c, self.be, self.mine
        al, bt, vA, vB, rc, bits, ids = [], [], [], [], [], [], []
        for j in mine:
            for i in range(c):
                a, b, va, vb, r1, d1, d2 = self._draw(j, 7)
                al.append(b32(a)); bt.append(b32(b)); vA.append(b32(va)); vB.append(b32(vb))
                rc.append(b32(r1) + b32(d1) + b32(d2))
                bits.append(self.bits[j][i]); ids.append(j)
        m = len(ids)
        alpha, beta = self._cat(al), self._cat(bt)
        pts = be.commit_points(alpha, beta, bytes(bits))                      # phi, A, B
        A = self._cat(pts[192 * k + 64:192 * k + 128] for k in range(m))
        B = self._cat(pts[192 * k + 128:192 * k + 192] for k in range(m))
        pokA = be.pokdlog_prove(A, alpha, ids, self._cat(vA))
        pokB = be.pokdlog_prove(B, beta, ids, self._cat(vB))
        if self.witness:
            com = be.powfcom_prove_w(pts, self._cat(a + b for a, b in zip(al, bt)), bytes(bits), ids, self._cat(rc))
        else:
            com = be.powfcom_prove(pts, alpha, bytes(bits), ids, self._cat(rc))
        self.beta = {j: [bt[q * c + i] for i in range(c)] for q, j in enumerate(mine)}
        self.alpha = {j: [al[q * c + i] for i in range(c)] for q, j in enumerate(mine)}
        self.cpts = {j: [pts[192 * (q * c + i):192 * (q * c + i + 1)] for i in range(c)] for q, j in enumerate(mine)}
        # Bidder::verifyCommitment (SEAL/bidder.cpp:1171-1195): every proof once
        if self.verify and self.n > 1:
            vA_ = be.pokdlog_verify(pokA, A, ids)
            vB_ = be.pokdlog_verify(pokB, B, ids)
            vC_ = be.powfcom_verify(com, pts, ids)
        else:
            vA_ = vB_ = vC_ = bytes([1] * m)
        for q, j in enumerate(mine):
            rec = bytearray()
            for k in range(q * c, (q + 1) * c):
                rec += pts[192 * k:192 * k + 192] + pokA[96 * k:96 * k + 96] + pokB[96 * k:96 * k + 96] + com[352 * k:352 * k + 352]
            self.sec["commit"][j] = bytes(rec)
            self.sec["commit_ok"][j] = all(vA_[k] and vB_[k] and vC_[k] for k in range(q * c, (q + 1) * c))

    def round_one(self, step):
        """Bidder::roundOne (SEAL/bidder.cpp:1203-1236) + verifyRoundOne (:1245-1262)."""
        be, mine = self.be, self.mine
        xs, rs, vx, vr = [], [], [], []
        for j in mine:
            x, r, a, b = self._draw(j, 4)
            xs.append(b32(x)); rs.append(b32(r)); vx.append(b32(a)); vr.append(b32(b))
        x_b, r_b = self._cat(xs), self._cat(rs)
        X = be.fixed_base_mul(x_b)
        R = be.fixed_base_mul(r_b)
        pokX = be.pokdlog_prove(X, x_b, mine, self._cat(vx))
        pokR = be.pokdlog_prove(R, r_b, mine, self._cat(vr))
        self.x, self.r, self.X, self.R = xs, rs, X, R
        if self.verify and self.n > 1:
            v1 = be.pokdlog_verify(pokX, X, mine)
            v2 = be.pokdlog_verify(pokR, R, mine)
        else:
            v1 = v2 = bytes([1] * len(mine))
        self.sec["r1"].append({j: X[64 * q:64 * q + 64] + R[64 * q:64 * q + 64] + pokX[96 * q:96 * q + 96] + pokR[96 * q:96 * q + 96]
                               for q, j in enumerate(mine)})
        self.sec["r1_ok"].append({j: bool(v1[q] and v2[q]) for q, j in enumerate(mine)})

    def round_two(self, step):
        """Bidder::roundTwo (SEAL/bidder.cpp:1271-1336) + verifyRoundTwo (:1346-1377)."""
        be, mine = self.be, self.mine
        X_all = self.exchange("X", self.X)
        Y_all = be.y_scan(X_all)                                          # :1286-1299
        Y = self._cat(Y_all[64 * j:64 * j + 64] for j in mine)
        base, ebit = [], []
        for q, j in enumerate(mine):
            bit = self.bits[j][step]
            if (not self.junction and bit == 0) or (self.junction and (bit == 0 or self.prev_bit[j] == 0)):
                base.append(Y[64 * q:64 * q + 64]); ebit.append(0)        # b = Y^x   :1303
            else:
                base.append(self.R[64 * q:64 * q + 64]); ebit.append(1)   # b = R^x   :1307
        b = be.var_base_mul(self._cat(base), self._cat(self.x))
        pt = lambda buf, q: buf[64 * q:64 * q + 64]
        stmt, sec, rnd = [], [], []
        if not self.junction:
            for q, j in enumerate(mine):
                stmt.append(pt(b, q) + pt(self.X, q) + pt(Y, q) + pt(self.R, q) + self.cpts[j][step])
                sec.append(self.x[q] + self.alpha[j][step] + (self.r[q] + self.beta[j][step] if self.witness else b""))
                rnd.append(self._cat(b32(v) for v in self._draw(j, 5)))
            stmt_b = self._cat(stmt)
            prove = be.stage1_prove_w if self.witness else be.stage1_prove
            proofs = prove(stmt_b, self._cat(sec), bytes(ebit), mine, self._cat(rnd))
            rec, tag = 672, 1
        else:
            P = self.prev
            for q, j in enumerate(mine):
                stmt.append(pt(b, q) + pt(self.X, q) + pt(self.R, q) + pt(P["b"], q) + pt(P["X"], q) + pt(P["R"], q) +
                            self.cpts[j][step] + pt(Y, q) + pt(P["Y"], q))
                sec.append(self.x[q] + P["x"][q] + self.alpha[j][step] +
                           (self.r[q] + P["r"][q] + self.beta[j][step] if self.witness else b""))
                rnd.append(self._cat(b32(v) for v in self._draw(j, 11)))
            stmt_b = self._cat(stmt)
            bjs = bytes(self.prev_bit[j] for j in mine)
            if self.witness:
                proofs = be.stage2_prove_w(stmt_b, self._cat(sec), bytes(ebit), bjs, bytes(self.bits[j][step] for j in mine), mine, self._cat(rnd))
            else:
                proofs = be.stage2_prove(stmt_b, self._cat(sec), bytes(ebit), bjs, mine, self._cat(rnd))
            rec, tag = 1344, 2
        if self.verify and self.n > 1:
            v = be.stage1_verify(proofs, stmt_b, mine) if tag == 1 else be.stage2_verify(proofs, stmt_b, mine)
        else:
            v = bytes([1] * len(mine))
        self.sec["r2"].append({j: struct.pack("<I", tag) + pt(b, q) + proofs[rec * q:rec * (q + 1)] for q, j in enumerate(mine)})
        self.sec["r2_ok"].append({j: bool(v[q]) for q, j in enumerate(mine)})
        self.Y, self.b = Y, b

    def round_three(self, step):
        """Bidder::roundThree (SEAL/bidder.cpp:1386-1421)."""
        b_all = self.exchange("b", self.b)
        deciding = not self.be.point_sum_is_inf(b_all)                   # :1393-1397
        if deciding:
            self.junction = True
            self.prev_step = step
            for j in self.mine:
                self.prev_bit[j] &= self.bits[j][step]                    # :1402 (true bit)
            self.max_bid |= 1 << (self.c - step - 1)                      # :1403 (64-bit shift here)
            self.prev = {"X": self.X, "R": self.R, "Y": self.Y, "b": self.b, "x": list(self.x), "r": list(self.r)}   # :1406-1411
        self.sec["r3"].append(1 if deciding else 0)

    def run_sections(self):
        self.commit()
        for step in range(self.c):
            self.round_one(step)
            self.round_two(step)
            self.round_three(step)
        return self.sec

    def run(self):
        out = assemble_transcript(self.n, self.c, self.seed, self.bids, [self.run_sections()])
        self.ok = transcript_ok(out) and self.max_bid == max(self.bids)
        return out


def _per_verifier(okl):
    """verifier j's byte = AND of the once-computed verdicts of every prover i != j"""
    bad = [i for i, v in enumerate(okl) if not v]
    if not bad:
        return bytes([1] * len(okl))
    if len(bad) == 1:   # only the failing prover itself still sees everybody else pass
        return bytes(1 if j == bad[0] else 0 for j in range(len(okl)))
    return bytes(len(okl))


def assemble_transcript(n, c, seed, bids, section_list):
    """PASEALT1 from the sections of one or several processes (each holding a subset of the bidders)."""
    def merged(key, step=None):
        d = {}
        for sec in section_list:
            d.update(sec[key] if step is None else sec[key][step])
        return d

    def per_verifier(ok):
        # reference: verifier j checks every i != j (SEAL/bidder.cpp:1178-1190); verifier j's answer is
        # the AND over i != j of the once-computed verdicts
        return _per_verifier([ok[i] for i in range(n)])

    out = bytearray(b"PASEALT1" + struct.pack("<QQQ", n, c, seed))
    for bid in bids:
        out += struct.pack("<Q", bid)
    cm, cok = merged("commit"), merged("commit_ok")
    for j in range(n):
        out += cm[j]
    out += per_verifier(cok)
    r3 = section_list[0]["r3"]
    maxbid = 0
    for step in range(c):
        r1, r1ok, r2, r2ok = merged("r1", step), merged("r1_ok", step), merged("r2", step), merged("r2_ok", step)
        for j in range(n):
            out += r1[j]
        out += per_verifier(r1ok)
        for j in range(n):
            out += r2[j]
        out += per_verifier(r2ok)
        out += bytes([r3[step]] * n)
        if r3[step]:
            maxbid |= 1 << (c - step - 1)
    for j in range(n):
        out += struct.pack("<Q", maxbid)
    return bytes(out)


def transcript_ok(buf):
    t = parse_transcript(buf)
    v = t["commit_verdict"] + b"".join(s["r1_verdict"] + s["r2_verdict"] for s in t["steps"])
    return all(v)


def parse_transcript(buf):
    """PASEALT1 -> dict of sections (used to feed single proofs to verifiers)."""
    assert buf[:8] == b"PASEALT1"
    n, c, seed = struct.unpack_from("<QQQ", buf, 8)
    off = 32
    bids = list(struct.unpack_from(f"<{n}Q", buf, off)); off += 8 * n
    t = {"n": n, "c": c, "seed": seed, "bids": bids, "commit": [], "steps": []}
    for j in range(n):
        row = []
        for i in range(c):
            row.append(buf[off:off + 736]); off += 736
        t["commit"].append(row)
    t["commit_verdict"] = buf[off:off + n]; off += n
    for step in range(c):
        s = {"r1": [], "r2": []}
        for j in range(n):
            s["r1"].append(buf[off:off + 320]); off += 320
        s["r1_verdict"] = buf[off:off + n]; off += n
        for j in range(n):
            tag = struct.unpack_from("<I", buf, off)[0]; off += 4
            b = buf[off:off + 64]; off += 64
            rec = 672 if tag == 1 else 1344
            s["r2"].append((tag, b, buf[off:off + rec])); off += rec
        s["r2_verdict"] = buf[off:off + n]; off += n
        s["r3"] = buf[off:off + n]; off += n
        t["steps"].append(s)
    t["max_bid"] = list(struct.unpack_from(f"<{n}Q", buf, off)); off += 8 * n
    assert off == len(buf), (off, len(buf))
    return t


def sections_to_transcripts(seed, n, c, bids, res):
    """Assemble the PASEALT1 transcript of every auction from the section arrays
    pa_seal_run returns (unsharded run: every auction fully local)."""
    A = len(n)
    m = sum(n)
    cmax = max(c)
    first = [0]
    for a in range(A):
        first.append(first[-1] + n[a])
    boff = [0]
    for a in range(A):
        for _ in range(n[a]):
            boff.append(boff[-1] + c[a])

    per_verifier = _per_verifier

    outs = []
    for a in range(A):
        na, ca, s0 = n[a], c[a], first[a]
        out = bytearray(b"PASEALT1" + struct.pack("<QQQ", na, ca, seed))
        for j in range(na):
            out += struct.pack("<Q", bids[s0 + j])
        ok = []
        for j in range(na):
            lo, hi = boff[s0 + j], boff[s0 + j + 1]
            out += res["commit"][736 * lo:736 * hi]
            ok.append(all(res["commit_ok"][lo:hi]))
        out += per_verifier(ok) if na > 1 else bytes([1] * na)
        for step in range(ca):
            base = step * m + s0
            for j in range(na):
                out += res["r1"][320 * (base + j):320 * (base + j + 1)]
            out += per_verifier([res["r1_ok"][base + j] for j in range(na)])
            for j in range(na):
                tag = res["r2_tag"][base + j]
                rec = 672 if tag == 1 else 1344
                out += struct.pack("<I", tag) + res["r2_b"][64 * (base + j):64 * (base + j + 1)]
                out += res["r2_proof"][1344 * (base + j):1344 * (base + j) + rec]
            out += per_verifier([res["r2_ok"][base + j] for j in range(na)])
            out += bytes([res["r3"][step * A + a]] * na)
        for j in range(na):
            out += struct.pack("<Q", res["max_bid"][a])
        outs.append(bytes(out))
    return outs
