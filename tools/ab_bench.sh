#!/bin/bash
# A/B timing of library builds (tools/ab_*.so, built with other -D flags) on the bench set (config 2) and config 3: ms per step
run() { python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra "$@" 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-50s %.4f ms' % ('$LABEL', d['ms_per_step']))"; }
for rep in 1 2 3; do
LABEL="default build" run
for f in tools/ab_*.so; do LABEL="$f" TREEDIST_LIB=$f run; done
done
LABEL="cfg3 default build" run --config 2 --steps 5
for f in tools/ab_*.so; do LABEL="cfg3 $f" TREEDIST_LIB=$f run --config 2 --steps 5; done
LABEL="cfg2 rf+wrf+euc default" run --metrics rf,wrf,euc
for f in tools/ab_*.so; do LABEL="cfg2 rf+wrf+euc $f" TREEDIST_LIB=$f run --metrics rf,wrf,euc; done
