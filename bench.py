#!/usr/bin/env python
"""bench.py -- tree-pairs/sec of the all-pairs inter-tree distance matrix (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm   (CUDA kernels through libtreedist_cuda)
    python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's CPU algorithm for the path,
                                                             timed on this box's host cores (oracle port; the real
                                                             treeCl + tree_distance stack cannot be installed offline)

Workload (weak scaling).  N=1: BASELINE configs[1] = 5,000 uniform random 50-taxon trees, rf + euc condensed matrices
produced by ONE fused launch per step.  N>1: the tree count grows as 5000*sqrt(N) so that every GPU keeps the same
12.5M pairs; rows are sharded over ranks with balanced pair counts, the encoded set is replicated, no collective sits
on the data path.  A "step" is one pass over the whole matrix.  `value` = pairs of the whole job / max-over-ranks
device time, inputs resident in HBM.  `e2e` = same metric through the public call with HOST buffers: Newick strings in
host memory -> encode -> H2D -> kernels -> D2H of both condensed matrices into pinned host memory, wall clock.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRICS = ('rf', 'euc')
BASE_TREES, N_TAXA = 5000, 50
SEED = 0xCA55E77E + 1          # treeCl RANDOM_SEED + config index


# BASELINE.json configs (index -> trees, taxa, metrics); configs[1] is the bench line, the others are for experiments
WORKLOAD_FMT = '{} uniform random {}-taxon trees, {} condensed matrices'      # both arms name the workload identically
CONFIGS = {0: (200, 20, ('rf', 'wrf', 'euc')), 1: (5000, 50, ('rf', 'euc')), 2: (20000, 100, ('wrf', 'euc')),
           4: (2000, 30, ('geo',))}


def workload(n_gpus, n_trees=None, n_taxa=None):
    from treecl_b200 import synth
    nt = n_trees or int(round(BASE_TREES * math.sqrt(n_gpus)))
    nx = n_taxa or N_TAXA
    return synth.uniform_trees(nt, nx, seed=SEED), nt, nx


def algorithmic_bytes(n_trees, n_taxa, n_splits, pairs, n_metrics):
    """inputs once (4 B id + 8 B length per internal edge, 8 B per leaf edge) + 8 B per pair per metric."""
    return n_trees * (12 * n_splits + 8 * n_taxa) + pairs * 8 * n_metrics


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for k, name in enumerate(names):
                if len(r) > 2 + k and r[2 + k].lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference(trees, metrics, budget_pairs, threads, reparse):
    """Oracle port on a bounded sample: the first `budget_pairs` pairs of the same workload, all metrics."""
    import oracle
    oracle.build()
    n = len(trees)
    total = n * (n - 1) // 2
    take = min(total, budget_pairs)
    t0 = time.perf_counter()
    for m in metrics:
        oracle.matrix_range(trees, m, 0, take, nthreads=threads, reparse_per_pair=reparse)
    dt = time.perf_counter() - t0
    return take / dt, take, dt


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    trees, nt, nx = workload(args.gpus, args.trees, args.taxa)
    threads = os.cpu_count() or 1
    pairs_total = nt * (nt - 1) // 2
    # reference-shaped: every pair parses both Newick strings (treeCl/tasks.py:41-75); size the sample for ~10 s/step
    probe, _, _ = cpu_reference(trees, METRICS, 20000, threads, True)
    sample = int(min(pairs_total, max(20000, probe * 8.0)))
    for _ in range(args.warmup):
        cpu_reference(trees, METRICS, min(sample, 20000), threads, True)
    times = []
    for _ in range(args.steps):
        v, take, dt = cpu_reference(trees, METRICS, sample, threads, True)
        times.append(dt)
    value = sample * args.steps / sum(times)
    best, take_b, dt_b = cpu_reference(trees, METRICS, min(pairs_total, 4000000), threads, False)
    desc = ('oracle port (plain C restatement of the reference path; real treeCl/tree_distance not installable offline), '
            'reference-shaped: both Newick strings re-parsed for every pair as treeCl/tasks.py:41-75 does, {} threads, '
            'first {} of {} pairs, metrics {}'.format(threads, sample, pairs_total, '+'.join(METRICS)))
    line = {
        'impl': 'reference', 'metric': 'tree-pairs/sec (rf+euc matrices)', 'value': value, 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(times) / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': WORKLOAD_FMT.format(nt, nx, '+'.join(METRICS)),
                   'sample_pairs_per_step': sample},
        'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port', 'sample': desc,
                         'best_cpu_value': best,
                         'best_cpu_sample': 'same oracle, trees encoded once, {} threads, first {} pairs'.format(threads, take_b)},
        'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from treecl_b200 import _lib, engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = max(world, 1)

    trees, nt, nx = workload(n_gpus, args.trees, args.taxa)
    pairs_total = nt * (nt - 1) // 2
    r0, r1 = engine.balanced_row_ranges(nt, n_gpus)[rank]
    p0, p1 = _lib.condensed_offset(nt, r0), _lib.condensed_offset(nt, r1)
    my_pairs = p1 - p0

    stream = torch.cuda.current_stream()
    ts = _lib.TreeSet(trees)
    ts.upload(local, stream.cuda_stream)
    info = ts.info
    outs = {m: torch.empty(max(my_pairs, 1), dtype=torch.float64, device=dev) for m in METRICS}
    ptrs = {m: outs[m].data_ptr() for m in METRICS}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev)    # 256 MiB, read (reduced) after the memset

    def step():
        ts.pairwise_device(METRICS, ptrs, row_begin=r0, row_end=r1, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.time()
    barrier()
    for k in range(args.steps):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair):
        flush_rd.sum()                     # write 256 MiB, then read 256 MiB so that L2 is cold AND holds no dirty lines
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    barrier()
    t_wall1 = time.time()
    launches = _lib.launch_count() - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = pairs_total * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public call, host buffers --------------------------------------------------------
    host = {m: torch.empty(max(my_pairs, 1), dtype=torch.float64).pin_memory() for m in METRICS}
    host_np = {m: host[m].numpy() for m in METRICS}

    def e2e_step():
        t = _lib.TreeSet(trees)                                   # parse + encode on the host
        t.pairwise(METRICS, row_begin=r0, row_end=r1, device=local, out=host_np)   # H2D, kernels, D2H
        h2d = upload_bytes(t.info, nt, METRICS)
        t.close()
        return h2d

    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 5))
    h2d_bytes = 0
    if e2e_steps:
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h2d_bytes = e2e_step()
    barrier()
    e2e_dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = pairs_total * e2e_steps / float(e2e_dt.item())

    if rank == 0:
        n_splits = info.max_splits
        alg = algorithmic_bytes(nt, nx, n_splits, my_pairs, len(METRICS))     # this rank's launch
        kernel_s = (sum(step_ms) / args.steps) * 1e-3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            pass
        peak = peaks.get('hbm_gbs', 6650.0)
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
            traffic = tr.get('pairwise_rf_euc_{}x{}'.format(nt, nx))
        except (OSError, ValueError):
            pass
        achieved = alg / kernel_s / 1e9
        threads = os.cpu_count() or 1
        cpu = None
        if n_gpus == 1 and not args.no_cpu_baseline:
            shaped, take_s, _ = cpu_reference(trees, METRICS, 200000, threads, True)
            best, take_b, _ = cpu_reference(trees, METRICS, min(pairs_total, 4000000), threads, False)
            cpu = {'value': shaped, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                   'sample': 'oracle port, reference-shaped (both Newick strings re-parsed per pair, treeCl/tasks.py:41-75), '
                             '{} threads, first {} pairs, rf+euc'.format(threads, take_s),
                   'best_cpu_value': best,
                   'best_cpu_sample': 'same oracle, trees encoded once, {} threads, first {} pairs'.format(threads, take_b)}
        line = {
            'metric': 'tree-pairs/sec ({} matrices)'.format('+'.join(METRICS)), 'value': value, 'unit': 'pairs/s', 'n_gpus': n_gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD_FMT.format(nt, nx, '+'.join(METRICS)), 'launches': 'one call = 3 kernels (events, tiles, exact fix-up) per step',
                       'pairs': pairs_total, 'pairs_per_gpu': my_pairs, 'parallelism': 'rows sharded x{}'.format(n_gpus),
                       'l2': '256 MiB memset + 256 MiB read between timed steps (outside the event pairs): L2 cold and clean; outputs (200 MB/step) exceed L2',
                       'dict_size': int(info.dict_size), 'dict_shared': int(info.dict_shared), 'max_shared': int(info.max_shared)},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback',
                         'algorithmic_bytes_per_launch': alg, 'kernel': 'pair_dense_kernel<{}> (fp64 tensor-core tile kernel; the step also runs pair_events_kernel and exact_fixup_kernel, and the '
                                   'duration used here is the whole step, so frac is a lower bound for the tile kernel)'.format('|'.join(METRICS))},
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(h2d_bytes),
                    'd2h_bytes_per_step': int(my_pairs * 8 * len(METRICS)),
                    'includes': 'Newick parse + split encode on host, H2D, kernels, D2H into pinned host memory'},
            'gpu_launches': int(launches), 'clocks': clocks,
            'step_ms': [round(x, 4) for x in step_ms],
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def upload_bytes(info, nt, metrics=('rf', 'euc')):
    sh = ((info.max_shared + 1) + 1) & ~1
    s = max(info.max_splits, 1)
    # leaf lengths, shared lists (+ posting index), per-tree scalars, postings (tree, end, length) ...
    b = nt * (info.n_taxa * 8 + sh * (16 + 4) + 4 + 7 * 8) + info.n_postings * 16
    if 'geo' in metrics or not info.uniform:
        b += nt * (s * (4 + 8 + 8 * info.words) + 8 * info.words + 4)      # ... and, for the kernels that need them, the full lists
    return b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--trees', type=int, default=None, help='override the tree count (testing)')
    ap.add_argument('--taxa', type=int, default=None, help='override the taxon count (testing)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--metrics', default=None, help='comma list overriding rf,euc (experiments)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--config', type=int, default=1, help='BASELINE.json configs index (experiments; default 1 = bench line)')
    args = ap.parse_args()
    global METRICS, BASE_TREES, N_TAXA, SEED
    if args.config != 1:
        BASE_TREES, N_TAXA, METRICS = CONFIGS[args.config]
        SEED = 0xCA55E77E + args.config
    if args.metrics:
        METRICS = tuple(args.metrics.split(','))
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
