"""The host C++ mirror of the reference's Bidder / BulletinBoard classes and the
`SEAL <n> <c>` command line (privacy-auction_b200/host), run as the reference's
own tests run it (exit code, SEAL/tests/CMakeLists.txt), and compared with the
transcripts and DataTracker byte totals of the unmodified reference."""
import glob
import json
import os
import random
import subprocess

import pytest

import seal_flow

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEAL = os.path.join(ROOT, "privacy-auction_b200", "bin", "SEAL")
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "seal_*.bin")))
SUMMARY = json.load(open(os.path.join(ROOT, "tests", "golden", "seal_reference_summary.json")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cli_reproduces_reference(tmp_path, path):
    gold = open(path, "rb").read()
    t = seal_flow.parse_transcript(gold)
    out = tmp_path / "t.bin"
    r = subprocess.run([SEAL, str(t["n"]), str(t["c"]), "--seed", str(t["seed"]), "--bids", ",".join(map(str, t["bids"])),
                        "--transcript", str(out), "--quiet"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_bytes() == gold
    s = json.loads([l for l in r.stderr.splitlines() if l.startswith("{")][-1])
    ref = SUMMARY[os.path.basename(path)]
    assert s["maxbid"] == ref["maxbid"]
    assert (s["data_bidder"], s["data_verifier"], s["data_total"]) == (ref["data_bidder"], ref["data_verifier"], ref["data_total"])


def test_cli_gentests_style_sweep():
    """the reference's own test: random `<n> <c>` pairs, pass = exit code 0 (tests/genTests.py:13-17)"""
    rnd = random.Random(2024)
    for _ in range(6):
        n, c = rnd.randint(1, 8), rnd.randint(1, 12)
        r = subprocess.run([SEAL, str(n), str(c), "--seed", str(rnd.randrange(1 << 30)), "--quiet"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (n, c, r.stderr[-1500:])


def test_cli_usage_error():
    r = subprocess.run([SEAL, "3"], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr


# ---- CCS22 command line ------------------------------------------------------------------------
CCS22 = os.path.join(ROOT, "privacy-auction_b200", "bin", "CCS22")
GOLDEN_CCS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "ccs22_*.bin")))


@pytest.mark.parametrize("path", GOLDEN_CCS, ids=[os.path.basename(p) for p in GOLDEN_CCS])
def test_ccs22_cli_reproduces_reference(tmp_path, path):
    import struct
    gold = open(path, "rb").read()
    n, c, seed, ev = struct.unpack_from("<QQQQ", gold, 8)
    bids = struct.unpack_from(f"<{n}Q", gold, 40)
    out = tmp_path / "t.bin"
    r = subprocess.run([CCS22, str(n), str(c), "--seed", str(seed), "--evaluator", str(ev), "--bids", ",".join(map(str, bids)),
                        "--transcript", str(out), "--quiet"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert out.read_bytes() == gold


def test_ccs22_cli_config2_and_sweep():
    """BASELINE config 2 (20 bidders, 32-bit bids) and a few genTests-style pairs: exit code 0"""
    r = subprocess.run([CCS22, "20", "32", "--seed", "4", "--quiet"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    rnd = random.Random(77)
    for _ in range(4):
        n, c = rnd.randint(1, 10), rnd.randint(1, 16)
        r = subprocess.run([CCS22, str(n), str(c), "--seed", str(rnd.randrange(1 << 30)), "--quiet"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (n, c, r.stderr[-1500:])


def test_cli_without_seed_uses_os_entropy(tmp_path):
    """No --seed: the draw stream is keyed from getrandom(2) and bids come from std::random_device, like the
    reference (SEAL/bidder.cpp:27, :97): two runs publish different records, both verify, and the transcript
    header carries no seed."""
    outs = []
    for k in range(2):
        out = tmp_path / f"t{k}.bin"
        r = subprocess.run([SEAL, "3", "4", "--bids", "11,6,13", "--transcript", str(out), "--quiet"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append(out.read_bytes())
    t0, t1 = (seal_flow.parse_transcript(o) for o in outs)
    assert t0["seed"] == 0 and t1["seed"] == 0
    assert t0["max_bid"] == t1["max_bid"] == [13, 13, 13]
    assert t0["commit"] != t1["commit"]
    assert seal_flow.transcript_ok(outs[0]) and seal_flow.transcript_ok(outs[1])
