# A/B helper: times the bench workload with the default build, then with every extra nvcc flag set given as an argument
# (e.g. "-DPS_EVWALK=0" "-DPS_NSTAGE=2 -DPS_MINB=4"); rebuilds kernels_pairsplit.o on the GPU box for each.
cd /root/repo
run() { for m in rf rf,euc wrf,euc; do TREEDIST_PAIR_KERNEL=split timeout 200 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --metrics $m 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', '$m', d['ms_per_step'])"; done; }
timeout 200 python tools/debug_compare.py 5000 2>&1 | grep -E "split" | head -5
run "default"
for v in "$@"; do rm -f treecl_b200/csrc/build/kernels_pairsplit.o; make -s -C treecl_b200/csrc PTXAS="$v" >/dev/null 2>&1; run "$v"; done
