#!/bin/sh
# A/B of engine builds under privacy-auction_b200/experiments (development aid, GPU box)
for lib in privacy-auction_b200/experiments/libpa_engine_*.so; do
  export PA_ENGINE_LIB=$PWD/$lib
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-seal | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib', round(d['value']/1e6,2), 'M/s', {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})"
  python tools/seal_profile.py config4 2 2>&1 | grep -E "^---" | tail -1
done
