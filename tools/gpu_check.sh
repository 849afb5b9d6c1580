#!/bin/bash
# One GPU box visit: parity tests, the bench line, the ncu launch list and one full capture of the two
# scalar-multiplication kernels.  usage (from the repo root, under gpurun):  bash tools/gpu_check.sh <tag> [what...]
# what: tests bench launches ncu   (default: all four)
tag=${1:-r02}; shift
what=${*:-tests bench launches ncu}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_smi.txt 2>&1
for w in $what; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > $out/${tag}_gpu_tests.log 2>&1
      echo "tests rc=$?" >> $out/${tag}_gpu_tests.log; tail -15 $out/${tag}_gpu_tests.log ;;
    bench)
      timeout 900 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
      echo "bench rc=$?"; tail -3 $out/${tag}_bench.err; head -c 1500 $out/${tag}_bench.json; echo ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-seal --no-cpu-baseline > $out/${tag}_launches.log 2>&1
      echo "launches rc=$?" ;;
    ncu)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_var_base|k_fixed_base' -s 6 -c 2 -f -o $out/${tag}_full \
        python bench.py --steps 2 --warmup 1 --no-seal --no-cpu-baseline > $out/${tag}_ncu.log 2>&1
      echo "ncu rc=$?"; ls -la $out/${tag}_full.ncu-rep ;;
  esac
done
