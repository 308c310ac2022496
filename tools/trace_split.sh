cd /root/repo
rm -f treecl_b200/csrc/build/kernels_pairsplit.o; make -s -C treecl_b200/csrc PTXAS="-DPS_TRACE $1" >/dev/null 2>&1
TREEDIST_PAIR_KERNEL=split timeout 200 python - <<'PY'
import os, sys
sys.path.insert(0, '.')
from treecl_b200 import _lib, synth
trees = synth.uniform_trees(5000, 50, seed=0xCA55E77E + 1)
ts = _lib.TreeSet(trees)
ts.pairwise(('rf', 'euc'))
os.environ['TREEDIST_TRACE'] = '/tmp/trace.bin'
ts.pairwise(('rf', 'euc'))
PY
python tools/trace_split.py /tmp/trace.bin
