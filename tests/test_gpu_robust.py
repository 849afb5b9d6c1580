"""What the reference's own tests never look at, through pa_seal_run and the verifier entry points:
a rejected draw (the phase-major schedule must fall back by itself), a record corrupted between proving and
verifying (the verdict bytes must say so, as the reference's verifiers would: SEAL/bidder.cpp:119-136,
1171-1195, 1245-1262, 1346-1377), received points that are off the curve or not canonical (what
EC_POINT_set_affine_coordinates refuses), and the keyed draw stream of a deployment."""
import os
import random

import pytest

import secp256k1_py as E
import seal_flow

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = E.P


@pytest.fixture()
def hooks(engine):
    yield engine
    engine.debug_set(1, 0)          # reject bits off
    engine.debug_set(2, 0)          # corruption off
    engine.set_entropy(None)
    E.REJECT_BITS = 0


@pytest.mark.parametrize("schedule", [0, 1], ids=["auto(phase-major)", "step-major"])
def test_rejected_draws_fall_back_and_match_oracle(hooks, oracle, schedule):
    """With the test hook every 16th draw is rejected (normally 2^-128): the arithmetic draw counters of the
    phase-major schedule are wrong, the runner must notice, run the auction again step-major by itself, and
    publish what the oracle publishes under the same rejection rule."""
    eng = hooks
    n, c, seed = 7, 9, 31
    bids = [random.Random(9).randrange(1 << c) for _ in range(n)]
    eng.debug_set(1, 4)
    E.REJECT_BITS = 4
    before = eng.reruns
    res = eng.seal_run(seed, [n], [c], bids, verify=True, sections=True, schedule=schedule)
    assert res["ok"] == [True] and res["max_bid"] == [max(bids)]
    got = seal_flow.sections_to_transcripts(seed, [n], [c], bids, res)[0]
    want = seal_flow.SealFlow(oracle, n, c, seed, bids).run()
    assert got == want
    assert eng.reruns == before + (1 if schedule == 0 else 0)
    # and the hook really changes the stream
    E.REJECT_BITS = 0
    assert seal_flow.SealFlow(oracle, n, c, seed, bids).run() != want


@pytest.mark.parametrize("schedule", [2, 1], ids=["phase-major", "step-major"])
@pytest.mark.parametrize("section,step,bidder,offset", [
    (1, 2, 1, 192 + 64 + 31),      # commitment record of bit 2 of bidder 1: rho of the Schnorr proof of A
    (1, 0, 3, 384 + 256 + 31),     # ... rho1 of the PoWFCom proof
    (2, 1, 2, 128 + 64 + 31),      # round-one record of step 1: rho of the Schnorr proof of X
    (2, 4, 0, 224 + 64 + 31),      # ... of R
    (3, 0, 1, 512 + 31),           # stage-1 proof of step 0: rho11
    (3, 4, 3, 1024 + 31),          # stage-2 proof of step 4: rho11
])
def test_corrupted_record_is_rejected(hooks, oracle, schedule, section, step, bidder, offset):
    """One bit of one published proof is flipped after proving: `ok` goes 0 and exactly that prover's verdict
    byte of that section goes 0; the oracle's verifier says the same about the corrupted record."""
    eng = hooks
    n, c, seed = 5, 6, 77
    bids = [21, 50, 7, 50, 33]     # junction at step 0 (bit 5 set in 50, 33): steps >= 1 carry stage-2 proofs
    eng.debug_set(2, section, (step << 32) | bidder, offset)
    res = eng.seal_run(seed, [n], [c], bids, verify=True, sections=True, schedule=schedule)
    assert res["ok"] == [False]
    assert res["max_bid"] == [max(bids)]          # the auction itself still runs to its end
    cok = [all(res["commit_ok"][c * j:c * (j + 1)]) for j in range(n)]
    r1ok = [bool(v) for v in res["r1_ok"]]
    r2ok = [bool(v) for v in res["r2_ok"]]
    bad = {"commit": [j for j in range(n) if not cok[j]], "r1": [i for i, v in enumerate(r1ok) if not v],
           "r2": [i for i, v in enumerate(r2ok) if not v]}
    want = {"commit": [], "r1": [], "r2": []}
    want[{1: "commit", 2: "r1", 3: "r2"}[section]] = [bidder if section == 1 else step * n + bidder]
    assert bad == want
    # the oracle on the published (corrupted) record
    ids = [bidder]
    if section == 1:
        rec = res["commit"][736 * (c * bidder + step):736 * (c * bidder + step + 1)]
        if offset < 384:
            which = (offset - 192) // 96
            assert oracle.pokdlog_verify(rec[192 + 96 * which:288 + 96 * which], rec[64 + 64 * which:128 + 64 * which], ids) == b"\x00"
        else:
            assert oracle.powfcom_verify(rec[384:736], rec[0:192], ids) == b"\x00"
    elif section == 2:
        rec = res["r1"][320 * (step * n + bidder):320 * (step * n + bidder + 1)]
        which = (offset - 128) // 96
        assert oracle.pokdlog_verify(rec[128 + 96 * which:224 + 96 * which], rec[64 * which:64 * which + 64], ids) == b"\x00"
    # without the hook the same auction verifies
    eng.debug_set(2, 0)
    assert eng.seal_run(seed, [n], [c], bids, verify=True, schedule=schedule)["ok"] == [True]


def _valid_pok(engine, oracle, n=8):
    rnd = random.Random(5)
    x = b"".join(rnd.randrange(1, E.N).to_bytes(32, "big") for _ in range(n))
    v = b"".join(rnd.randrange(1, E.N).to_bytes(32, "big") for _ in range(n))
    X = oracle.fixed_base_mul(x)
    ids = list(range(n))
    proofs = engine.pokdlog_prove(X, x, ids, v)
    assert engine.pokdlog_verify(proofs, X, ids) == bytes([1] * n)
    return proofs, X, ids


def test_verifiers_refuse_points_off_the_curve_or_not_canonical(engine, oracle):
    """A verifier checks what came off the wire: y^2 = x^3 + 7 and both coordinates < p, or the 64 zero bytes of
    infinity.  The a = 0 formulas never use the 7, and x + p hashes differently from x, so neither may pass."""
    proofs, X, ids = _valid_pok(engine, oracle)
    n = len(ids)
    bad_p, bad_X = bytearray(proofs), bytearray(X)
    # 0: eps off the curve (y + 1)
    y = int.from_bytes(proofs[32:64], "big")
    bad_p[32:64] = ((y + 1) % P).to_bytes(32, "big")
    # 1: statement off the curve
    yx = int.from_bytes(X[64 + 32:64 + 64], "big")
    bad_X[64 + 32:64 + 64] = ((yx + 1) % P).to_bytes(32, "big")
    # 2: a non-canonical encoding: a curve point with x < 2^256 - p sent as (x + p, y)
    x = 1
    while pow((x ** 3 + 7) % P, (P - 1) // 2, P) != 1:
        x += 1
    ysmall = pow((x ** 3 + 7) % P, (P + 1) // 4, P)
    small = x.to_bytes(32, "big") + ysmall.to_bytes(32, "big")
    twin = (x + P).to_bytes(32, "big") + ysmall.to_bytes(32, "big")
    bad_X[64 * 2:64 * 3] = twin
    # 3: (p, p), which reduces to (0, 0) but is not the encoding of infinity
    bad_X[64 * 3:64 * 4] = P.to_bytes(32, "big") * 2
    # 4: eps = (p, p)
    bad_p[96 * 4:96 * 4 + 64] = P.to_bytes(32, "big") * 2
    got = engine.pokdlog_verify(bytes(bad_p), bytes(bad_X), ids)
    assert list(got) == [0, 0, 0, 0, 0] + [1] * (n - 5)
    probe = small + twin + P.to_bytes(32, "big") * 2 + bytes(64) + bytes(bad_X[64:128]) + X[0:64]
    assert list(engine.point_on_curve(probe)) == [1, 0, 0, 1, 0, 1]
    # a statement that IS the small point, canonical, passes the point test and fails only on the equation
    ok_X = bytearray(X)
    ok_X[0:64] = small
    assert list(engine.pokdlog_verify(proofs, bytes(ok_X), ids)) == [0] + [1] * (n - 1)


def test_keyed_stream_hides_the_seed(hooks):
    """pa_ctx_set_entropy: with a key installed the same (seed, bids) publish different records, still valid;
    two keys give two different transcripts; removing the key brings the seeded test stream back."""
    eng = hooks
    n, c, seed, bids = 3, 4, 42, [11, 6, 13]
    plain = eng.seal_run(seed, [n], [c], bids, verify=True, sections=True)
    eng.set_entropy(bytes(range(32)))
    k1 = eng.seal_run(seed, [n], [c], bids, verify=True, sections=True)
    eng.set_entropy(bytes(range(1, 33)))
    k2 = eng.seal_run(seed, [n], [c], bids, verify=True, sections=True)
    eng.set_entropy(None)
    again = eng.seal_run(seed, [n], [c], bids, verify=True, sections=True)
    for r in (plain, k1, k2, again):
        assert r["ok"] == [True] and r["max_bid"] == [13]
    assert plain["commit"] == again["commit"] and plain["r1"] == again["r1"]
    assert len({plain["commit"], k1["commit"], k2["commit"]}) == 3
    assert len({plain["r1"], k1["r1"], k2["r1"]}) == 3
