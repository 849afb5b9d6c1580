#!/bin/sh
# A/B of the END-TO-END figure (host buffers through the C ABI) between the in-tree build and the builds under
# privacy-auction_b200/experiments (development aid, GPU box); two repetitions each.
for lib in default privacy-auction_b200/experiments/libpa_engine_*.so; do
  [ "$lib" = default ] || [ -f "$lib" ] || continue
  if [ "$lib" = default ]; then unset PA_ENGINE_LIB; else export PA_ENGINE_LIB=$PWD/$lib; fi
  for rep in 1 2; do
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-seal | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value']/1e6,2), 'M/s device', round(d['e2e']['value']/1e6,2), 'M/s e2e')"
  done
done
