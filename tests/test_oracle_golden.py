"""Pins the checker to the UNMODIFIED reference: the libcrypto oracle port
(oracle/pa_oracle.c) driven by tests/seal_flow.py must reproduce, byte for byte,
the transcripts that oracle/_ref/seal_ref (the reference's own SEAL classes,
compiled from /root/reference with only the RNG renamed) wrote into
tests/golden/ (generator: tests/golden/make_golden.sh)."""
import glob
import os

import pytest

import seal_flow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "seal_*.bin")))


def test_goldens_present():
    assert len(GOLDEN) >= 7


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_reference_transcript(oracle, path):
    gold = open(path, "rb").read()
    t = seal_flow.parse_transcript(gold)
    fl = seal_flow.SealFlow(oracle, t["n"], t["c"], t["seed"], t["bids"])
    assert fl.run() == gold
    assert fl.ok and [fl.max_bid] * t["n"] == t["max_bid"] == [max(t["bids"])] * t["n"]


def test_goldens_cover_every_branch():
    """both stage-1 bits, all three stage-2 branches, infinity (n = 1), all-zero bids, ties"""
    tags, stage2_shapes, r3 = set(), set(), set()
    for path in GOLDEN:
        t = seal_flow.parse_transcript(open(path, "rb").read())
        for s in t["steps"]:
            r3.update(s["r3"])
            for tag, b, proof in s["r2"]:
                tags.add(tag)
                if tag == 2:
                    rho11 = proof[1024:1056]
                    stage2_shapes.add("rho11=0" if rho11 == bytes(32) else "rho11!=0")
                if b == bytes(64):
                    tags.add("b=inf")
    assert tags >= {1, 2, "b=inf"} and stage2_shapes == {"rho11=0", "rho11!=0"} and r3 == {0, 1}


@pytest.mark.parametrize("witness", [False, True], ids=["plain", "witness"])
@pytest.mark.parametrize("path", GOLDEN[:4], ids=[os.path.basename(p) for p in GOLDEN[:4]])
def test_device_proof_code_host_build_reproduces_reference(oracle, path, witness):
    """pa_proof.cuh / pa_sha256.cuh compiled for the host (PTX carry flag emulated):
    provers and verifiers, one proof at a time, against the reference transcript.  witness: the provers that use
    the prover's knowledge of discrete logarithms (pa_*_prove_w) must publish the very same bytes."""
    import hostcheck_backend
    gold = open(path, "rb").read()
    t = seal_flow.parse_transcript(gold)
    fl = seal_flow.SealFlow(hostcheck_backend.HostcheckBackend(oracle), t["n"], t["c"], t["seed"], t["bids"], witness=witness)
    assert fl.run() == gold and fl.ok
