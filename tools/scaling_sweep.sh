# 1/2/4/8-GPU weak-scaling sweep of bench.py on one box (run under gpurun --gpus 8); one JSON line per N in gpurun_out/scale_*.json
cd /root/repo
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_$n.json
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_$n.json; fi
  python -c "import json,sys; d=json.load(open('gpurun_out/scale_$n.json')); print($n, d['config']['workload'][:40], d['config']['pairs'], round(d['ms_per_step'],4), '%.3e'%d['value'], '%.3e'%d['e2e']['value'])"
done
