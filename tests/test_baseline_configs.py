"""BASELINE.json configs 1 and 2 at full size: `./SEAL 10 20` and `CCS22 20 x 32` (and 31, where the
reference's own bids do not degenerate).  The unmodified reference's transcripts are pinned by
SHA-256 (tests/golden/baseline_config_digests.json, made by oracle/_ref/{seal_ref,ccs22_ref});
the oracle port (CPU) and the CUDA engine (GPU) must reproduce them byte for byte."""
import hashlib
import json
import os

import pytest

import ccs22_flow
import seal_flow

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = json.load(open(os.path.join(ROOT, "tests", "golden", "baseline_config_digests.json")))


def _sha(b):
    return hashlib.sha256(b).hexdigest()


def test_oracle_seal_10_20(oracle):
    g = D["seal_10_20"]
    fl = seal_flow.SealFlow(oracle, g["n"], g["c"], g["seed"], g["bids"])
    out = fl.run()
    assert len(out) == g["bytes"] and _sha(out) == g["sha256"] and fl.ok


@pytest.mark.parametrize("name", ["ccs22_20_32", "ccs22_20_31"])
def test_oracle_ccs22_config2(oracle, name):
    g = D[name]
    out = ccs22_flow.Ccs22Flow(oracle, g["n"], g["c"], g["seed"], g["evaluator"], g["bids"]).run()
    assert len(out) == g["bytes"] and _sha(out) == g["sha256"]


@pytest.mark.gpu
def test_engine_seal_10_20_runner_and_batched_abi(engine):
    g = D["seal_10_20"]
    res = engine.seal_run(g["seed"], [g["n"]], [g["c"]], g["bids"], verify=True, sections=True)
    out = seal_flow.sections_to_transcripts(g["seed"], [g["n"]], [g["c"]], g["bids"], res)[0]
    assert _sha(out) == g["sha256"] and res["ok"] == [True] and res["max_bid"] == [max(g["bids"])]
    fl = seal_flow.SealFlow(engine, g["n"], g["c"], g["seed"], g["bids"])
    assert _sha(fl.run()) == g["sha256"] and fl.ok


@pytest.mark.gpu
def test_engine_seal_10_20_step_major_schedule(engine):
    g = D["seal_10_20"]
    res = engine.seal_run(g["seed"], [g["n"]], [g["c"]], g["bids"], verify=True, sections=True, schedule=1)
    out = seal_flow.sections_to_transcripts(g["seed"], [g["n"]], [g["c"]], g["bids"], res)[0]
    assert _sha(out) == g["sha256"] and res["ok"] == [True]


@pytest.mark.gpu
def test_engine_config4_full_size_schedules_agree(engine):
    """BASELINE config 4 at full size (1000 bidders x 32-bit bids, ~4.5 M scalar mults; far beyond the CPU
    oracle): both schedules of the runner must publish the same 47 MB of records, every proof must
    verify, and the maximum must come out.  The step-major bytes are tied to the oracle at n = 64
    (tests/test_gpu_seal.py) and to the reference at n <= 10."""
    import random
    rnd = random.Random(2024)
    n, c = 1000, 32
    bids = [rnd.randrange(1 << 31) for _ in range(n)]
    a = engine.seal_run(4, [n], [c], bids, verify=True, sections=True, schedule=1)
    b = engine.seal_run(4, [n], [c], bids, verify=True, sections=True, schedule=2)
    assert a["ok"] == b["ok"] == [True] and a["max_bid"] == b["max_bid"] == [max(bids)]
    for key in ("commit", "commit_ok", "r1", "r1_ok", "r2_tag", "r2_b", "r2_proof", "r2_ok", "r3"):
        assert _sha(bytes(a[key])) == _sha(bytes(b[key])), key


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ccs22_20_32", "ccs22_20_31"])
def test_engine_ccs22_config2(engine, name):
    g = D[name]
    out = ccs22_flow.Ccs22Flow(engine, g["n"], g["c"], g["seed"], g["evaluator"], g["bids"]).run()
    assert _sha(out) == g["sha256"]


@pytest.mark.gpu
def test_cli_seal_10_20(tmp_path):
    import subprocess
    g = D["seal_10_20"]
    out = tmp_path / "t.bin"
    r = subprocess.run([os.path.join(ROOT, "privacy-auction_b200", "bin", "SEAL"), "10", "20", "--seed", str(g["seed"]), "--bids",
                        ",".join(map(str, g["bids"])), "--transcript", str(out), "--quiet"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-1500:]
    assert _sha(out.read_bytes()) == g["sha256"]
    s = json.loads([l for l in r.stderr.splitlines() if l.startswith("{")][-1])
    assert (s["data_bidder"], s["data_verifier"], s["data_total"]) == (g["data_bidder"], g["data_verifier"], g["data_total"])
