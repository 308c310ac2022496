#!/bin/bash
# A/B timing of library builds / launch options on the bench set (config 2) and config 3: prints ms per step
run() { python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra "$@" 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-60s %.4f ms' % ('$LABEL', d['ms_per_step']))"; }
for rep in 1 2; do
LABEL="default (pdl, parts auto, diet)" run
LABEL="no pdl" TREEDIST_NO_PDL=1 run
LABEL="parts=1" TREEDIST_EVENT_PARTS=1 run
LABEL="parts=4" TREEDIST_EVENT_PARTS=4 run
LABEL="no pdl, parts=1" TREEDIST_NO_PDL=1 TREEDIST_EVENT_PARTS=1 run
LABEL="no diet lib" TREEDIST_LIB=tools/ab_nodiet.so run
LABEL="round-1 lib" TREEDIST_LIB=tools/ab_libtreedist_r1.so run
done
LABEL="cfg3 default" run --config 2 --steps 5
LABEL="cfg3 no pdl parts=1" TREEDIST_NO_PDL=1 TREEDIST_EVENT_PARTS=1 run --config 2 --steps 5
LABEL="cfg2 rf only" run --metrics rf
LABEL="cfg2 rf only no pdl parts=1" TREEDIST_NO_PDL=1 TREEDIST_EVENT_PARTS=1 run --metrics rf
