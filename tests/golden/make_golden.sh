#!/bin/sh
# Regenerates tests/golden/seal_*.bin by running the UNMODIFIED reference SEAL
# classes (oracle/_ref/seal_ref, built by `make -C oracle ref` from
# /root/reference with the seeded RNG shim force-included).  Build container only.
# args of seal_ref: <n> <c> <seed> <bids|-> <out>
set -e
cd "$(dirname "$0")/../.."
make -s -C oracle ref
for cfg in "3 4 42 -" "1 3 1 5" "2 3 1 -" "4 6 7 -" "5 5 11 0,0,0,0,0" "3 8 5 255,255,3" "6 4 99 -"; do
  set -- $cfg
  ./oracle/_ref/seal_ref "$1" "$2" "$3" "$4" "tests/golden/seal_n$1_c$2_s$3.bin"
done
# tests/golden/seal_reference_summary.json holds, per transcript, the JSON line seal_ref
# prints on stderr (sha256 of the transcript, max bid, the reference's DataTracker byte totals).
# CCS22: ccs22_ref <n> <c> <seed> <evaluatorId> <bids|-> <out>
for cfg in "4 5 9 2 -" "2 3 1 0 -" "3 6 4 1 0,0,0" "5 4 8 4 -" "1 3 2 0 5" "4 8 3 0 200,7,200,13"; do
  set -- $cfg
  ./oracle/_ref/ccs22_ref "$1" "$2" "$3" "$4" "$5" "tests/golden/ccs22_n$1_c$2_s$3_e$4.bin"
done
# BASELINE configs 1 and 2 at full size (digests only, tests/golden/baseline_config_digests.json):
#   ./oracle/_ref/seal_ref 10 20 2024 - /tmp/seal_10_20.bin          (~52 s)
#   ./oracle/_ref/ccs22_ref 20 32 2024 7 - /tmp/ccs22_20_32.bin ; ./oracle/_ref/ccs22_ref 20 31 2025 3 - /tmp/ccs22_20_31.bin
