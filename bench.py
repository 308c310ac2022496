#!/usr/bin/env python
"""bench.py -- tree-pairs/sec of the all-pairs inter-tree distance matrix (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm   (CUDA kernels through libtreedist_cuda)
    python bench.py --impl reference --gpus N --steps K ...  reference arm: the reference's CPU algorithm for the path,
                                                             timed on this box's host cores (oracle port; the real
                                                             treeCl + tree_distance stack cannot be installed offline)

Workloads (BASELINE.json `configs`, generators in treecl_b200/synth.py):
  N = 1   main line = configs[1]: 5,000 uniform random 50-taxon trees, rf + euc condensed matrices, one call per step.
  N > 1   main line = configs[2]: 20,000 uniform random 100-taxon trees, wrf + euc, STRONG scaling: the fixed matrix is split
          over the ranks in folded row blocks (engine.folded_row_ranges: every rank a block of long rows and a block of short
          rows, equal pair counts), the set is encoded once on rank 0 and its serialised form broadcast; no collective on the
          data path.
  `config.extra` carries the other configs on the same GPUs, each with pairs/s, ms per step and a spot parity flag against
  the oracle:  N = 1: configs[2], [3], [4] on one GPU;  N > 1: configs[1] weak-scaled (5000*sqrt(N) trees, the round-1 line),
  configs[3] (100,000 x 200 clustered trees, RF as uint16 through the int8 tcgen05 incidence contraction, row blocks streamed
  to pinned host memory in the e2e figure) and configs[4] (2,000 x 30, geodesic) strong-scaled.
A "step" is one pass over the whole matrix.  `value` = pairs of the whole job / max-over-ranks device time, inputs resident in
HBM.  `e2e` = the same through the public call with HOST buffers: Newick strings in host memory -> encode -> (broadcast) -> H2D
-> kernels -> D2H of the condensed matrices into pinned host memory, wall clock, max over ranks.
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

SEED0 = 0xCA55E77E              # treeCl RANDOM_SEED; config k uses SEED0 + k
WORKLOAD_FMT = '{} {} {}-taxon trees, {} condensed matrices'      # both arms name the workload identically


def config_of(index):
    from treecl_b200 import synth
    return synth.BASELINE_CONFIGS[index]


def workload_name(cfg, n_trees=None):
    kind = 'uniform random' if cfg['kind'] == 'uniform' else 'clustered (simulator-style)'
    return WORKLOAD_FMT.format(n_trees or cfg['n_trees'], kind, cfg['n_taxa'], '+'.join(cfg['metrics']))


def make_strings(index, n_trees=None, n_taxa=None):
    """Newick strings of BASELINE config `index` (NumPy generator for the small configs, the native one for configs 2 and 3)."""
    from treecl_b200 import synth
    cfg = config_of(index)
    nt, nx = n_trees or cfg['n_trees'], n_taxa or cfg['n_taxa']
    if index in (2, 3) or nt > 6000:
        kw = {k: cfg[k] for k in ('n_classes', 'class_nni', 'lam') if k in cfg}
        return synth.native_trees(nt, nx, cfg['kind'], seed=SEED0 + index, **kw)
    return synth.uniform_trees(nt, nx, seed=SEED0 + index)


def algorithmic_bytes(n_trees, n_taxa, n_splits, pairs, out_bytes_per_pair):
    """inputs once (4 B id + 8 B length per internal edge, 8 B per leaf edge) + the output bytes per pair."""
    return n_trees * (12 * n_splits + 8 * n_taxa) + pairs * out_bytes_per_pair


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


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle port on a bounded sample
# ---------------------------------------------------------------------------------------------------------------------
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
    index = args.config if args.config is not None else (1 if args.gpus <= 1 else 2)
    cfg = config_of(index)
    metrics = tuple(args.metrics.split(',')) if args.metrics else cfg['metrics']
    trees = make_strings(index, args.trees, args.taxa)
    nt = len(trees)
    threads = os.cpu_count() or 1
    pairs_total = nt * (nt - 1) // 2
    # reference-shaped: every pair parses both Newick strings (treeCl/tasks.py:41-75); size the sample for ~8 s/step
    probe, _, _ = cpu_reference(trees, metrics, 20000, threads, True)
    sample = int(min(pairs_total, max(20000, probe * 8.0)))
    for _ in range(args.warmup):
        cpu_reference(trees, metrics, min(sample, 20000), threads, True)
    times = []
    for _ in range(args.steps):
        v, take, dt = cpu_reference(trees, metrics, sample, threads, True)
        times.append(dt)
    value = sample * args.steps / sum(times)
    best, take_b, dt_b = cpu_reference(trees, metrics, min(pairs_total, 4000000), threads, False)
    desc = ('oracle port (plain C restatement of the reference path; real treeCl/tree_distance not installable offline), '
            'reference-shaped: both Newick strings re-parsed for every pair as treeCl/tasks.py:41-75 does, {} threads, '
            'first {} of {} pairs, metrics {}'.format(threads, sample, pairs_total, '+'.join(metrics)))
    line = {
        'impl': 'reference', 'metric': 'tree-pairs/sec ({} matrices)'.format('+'.join(metrics)), 'value': value, 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(times) / args.steps,
        'higher_is_better': True, 'scaling': 'weak' if args.gpus <= 1 else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(cfg, nt), 'sample_pairs_per_step': sample},
        'cpu_baseline': {'value': value, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port', 'sample': desc,
                         'best_cpu_value': best,
                         'best_cpu_sample': 'same oracle, trees encoded once, {} threads, first {} pairs'.format(threads, take_b)},
        'e2e': {'value': value, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
class Ctx(object):
    """Rank / device / collectives of this process."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; the product has no CPU fallback')
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)        # > 126 MB L2
        self.flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=self.dev)    # 256 MiB, read (reduced) after the memset

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def flush_l2(self):
        self.flush.zero_()                 # write 256 MiB, then read 256 MiB: L2 cold AND no dirty lines left
        self.flush_rd.sum()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def share_treeset(self, make):
        """Encode ONCE: rank 0 calls make() -> TreeSet; its serialised bytes are broadcast (NCCL) and the other ranks deserialise
        (td_treeset_serialize / td_treeset_deserialize).  Returns (TreeSet, serialised bytes)."""
        from treecl_b200 import _lib
        torch, dist = self.torch, self.dist
        if self.world == 1:
            return make(), 0
        size = torch.zeros(1, dtype=torch.int64, device=self.dev)
        ts = None
        if self.rank == 0:
            ts = make()
            blob = torch.from_numpy(ts.serialize()).to(self.dev, non_blocking=False)
            size[0] = blob.numel()
        dist.broadcast(size, 0)
        if self.rank != 0:
            blob = torch.empty(int(size.item()), dtype=torch.uint8, device=self.dev)
        dist.broadcast(blob, 0)
        if self.rank != 0:
            ts = _lib.TreeSet.deserialize(blob.cpu().numpy())
        return ts, int(size.item())


def my_blocks(ctx, n_trees, align=64):
    """This rank's folded row blocks.  align: 64 rows = a tile row of the pair kernels; 512 = a tile of the 2-SM incidence kernel (a block
    boundary inside a tile makes both neighbours compute it)."""
    from treecl_b200 import engine
    return engine.folded_row_ranges(n_trees, ctx.world, align=align)[ctx.rank]


def block_pairs(n, blocks):
    from treecl_b200 import _lib
    return sum(_lib.condensed_offset(n, b) - _lib.condensed_offset(n, a) for a, b in blocks)


def timed_steps(ctx, step, steps, warmup):
    """`warmup` untimed steps, then `steps` steps each bracketed by CUDA events on the launching stream, L2 flushed in between
    (outside the event pairs).  Returns (per-step ms of this rank, max over ranks of the total)."""
    torch = ctx.torch
    for _ in range(max(warmup, 0)):
        step()
    ctx.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        ctx.flush_l2()
        ev[k][0].record(ctx.stream)
        step()
        ev[k][1].record(ctx.stream)
    ctx.barrier()
    ms = [a.elapsed_time(b) for a, b in ev]
    return ms, ctx.max_over_ranks(sum(ms))


def device_outputs(ctx, n, blocks, metrics, dtype=None):
    """Per block and metric a device buffer; returns [(row_begin, row_end, {metric: tensor})]."""
    from treecl_b200 import _lib
    torch = ctx.torch
    out = []
    for a, b in blocks:
        pairs = _lib.condensed_offset(n, b) - _lib.condensed_offset(n, a)
        out.append((a, b, {m: torch.empty(max(pairs, 1), dtype=dtype or torch.float64, device=ctx.dev) for m in metrics}))
    return out


def spot_parity(trees, n, blocks, outs, metrics, n_rows, normalise=False):
    """A few random rows of this rank's blocks against the oracle (rf bit-exact, the others to 1e-9 relative).  trees: Newick strings."""
    import oracle
    rng = np.random.default_rng(7)
    ok, checked = True, 0
    threads = os.cpu_count() or 1
    rows = []
    for a, b, _ in outs:
        if b - a > 1:
            rows += [int(x) for x in rng.choice(np.arange(a, min(b, n - 1)), size=min(n_rows, max(1, min(b, n - 1) - a)), replace=False)]
    if not rows:
        return True, 0
    for m in metrics:
        want = oracle.matrix_rows(trees, m, rows, normalise=normalise, nthreads=threads)
        for k, r in enumerate(rows):
            a, b, bufs = next(o for o in outs if o[0] <= r < o[1])
            off = (r * n - r * (r + 1) // 2) - (a * n - a * (a + 1) // 2)
            got = bufs[m][off:off + (n - 1 - r)].cpu().numpy().astype(np.float64)
            w = want[k, r + 1:]
            good = np.array_equal(got, w) if m == 'rf' else bool(np.all(np.abs(got - w) <= 1e-9 * np.abs(w) + 1e-12))
            ok &= good
            checked += w.size
    return bool(ok), int(checked)


def run_config(ctx, index, steps, warmup, n_trees=None, parity_rows=2, u16=False, weak=False):
    """One BASELINE config on all ranks: encode once (rank 0) + broadcast, folded row blocks, device-timed steps, spot parity on rank 0."""
    from treecl_b200 import _lib, synth
    cfg = config_of(index)
    nt = n_trees or cfg['n_trees']
    metrics = cfg['metrics']
    holder = {}

    def make():
        if index in (2, 3) or nt > 6000:
            kw = {k: cfg[k] for k in ('n_classes', 'class_nni', 'lam') if k in cfg}
            holder['g'] = synth.NativeTrees(nt, cfg['n_taxa'], cfg['kind'], seed=SEED0 + index, **kw)
            return holder['g'].treeset()
        holder['s'] = synth.uniform_trees(nt, cfg['n_taxa'], seed=SEED0 + index)
        return _lib.TreeSet(holder['s'])

    t0 = time.perf_counter()
    ts, blob_bytes = ctx.share_treeset(make)
    t_setup = time.perf_counter() - t0
    ts.upload(ctx.local, ctx.stream.cuda_stream)
    info = ts.info
    blocks = my_blocks(ctx, nt, 512 if u16 else 64)
    pairs_total = nt * (nt - 1) // 2
    outs = device_outputs(ctx, nt, blocks, metrics, ctx.torch.int16 if u16 else None)

    blocks_dev = [(a, b, {m: bufs[m].data_ptr() for m in metrics}) for a, b, bufs in outs]

    def step():
        if u16:
            for a, b, bufs in outs:
                ts.rf_incidence_device(bufs['rf'].data_ptr(), out_is_u16=True, row_begin=a, row_end=b, stream=ctx.stream.cuda_stream)
        elif blocks_dev:
            ts.pairwise_device_blocks(metrics, blocks_dev, stream=ctx.stream.cuda_stream)      # both folded blocks of this rank in one call

    ms, total_ms = timed_steps(ctx, step, steps, warmup)
    ts.check_errors(ctx.local)
    res = {'workload': workload_name(cfg, nt), 'pairs': pairs_total, 'pairs_this_rank': block_pairs(nt, blocks),
           'ms_per_step': total_ms / steps, 'pairs_per_s': pairs_total * steps / (total_ms * 1e-3), 'steps': steps,
           'scaling': 'weak' if weak else 'strong', 'output': 'u16' if u16 else 'f64',
           'dict_size': int(info.dict_size), 'dict_shared': int(info.dict_shared), 'mean_common_splits': round(float(info.mean_common), 3),
           'setup_s': round(t_setup, 2), 'serialised_set_bytes': blob_bytes}
    if index == 4:
        st = ts.geo_stats()
        if st['pairs']:
            res['gtp'] = {'maxflow_per_pair': round(st['maxflow_solves'] / st['pairs'], 2), 'ratios_per_pair': round(st['ratios'] / st['pairs'], 2)}
    if ctx.rank == 0 and parity_rows:
        strings = holder['g'].strings() if 'g' in holder else holder['s']
        ok, checked = spot_parity(strings, nt, blocks, outs, metrics, parity_rows)
        res['parity_vs_oracle'] = {'ok': ok, 'pairs_checked': checked, 'rows_per_block': parity_rows}
    if 'g' in holder:
        holder['g'].close()
    return res, ts, outs, holder


def run_ours(args):
    from treecl_b200 import _lib, engine, synth
    ctx = Ctx()
    torch = ctx.torch
    n_gpus = ctx.world
    index = args.config if args.config is not None else (1 if n_gpus == 1 else 2)
    cfg = dict(config_of(index))
    if args.metrics:
        cfg['metrics'] = tuple(args.metrics.split(','))
    metrics = cfg['metrics']
    nt, nx = args.trees or cfg['n_trees'], args.taxa or cfg['n_taxa']
    pairs_total = nt * (nt - 1) // 2

    # ---- the set: Newick strings live in host memory on rank 0 (the e2e leg starts from them) -----------------------------------
    strings, gen = None, None
    if ctx.rank == 0:
        if index in (2, 3) or nt > 6000:
            kw = {k: cfg[k] for k in ('n_classes', 'class_nni', 'lam') if k in cfg}
            gen = synth.NativeTrees(nt, nx, cfg['kind'], seed=SEED0 + index, **kw)
        else:
            strings = synth.uniform_trees(nt, nx, seed=SEED0 + index)

    def encode():
        return gen.treeset() if gen is not None else _lib.TreeSet(strings)

    ts, blob_bytes = ctx.share_treeset(encode)
    ts.upload(ctx.local, ctx.stream.cuda_stream)
    info = ts.info
    blocks = my_blocks(ctx, nt)
    my_pairs = block_pairs(nt, blocks)
    outs = device_outputs(ctx, nt, blocks, metrics)

    blocks_dev = [(a, b, {m: bufs[m].data_ptr() for m in metrics}) for a, b, bufs in outs]

    def step():
        if blocks_dev:
            ts.pairwise_device_blocks(metrics, blocks_dev, stream=ctx.stream.cuda_stream)          # both folded blocks of this rank in one call

    for _ in range(max(args.warmup, 0)):
        step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    launches0 = _lib.launch_count()
    t_wall0 = time.time()
    step_ms, total_ms = timed_steps(ctx, step, args.steps, 0)
    t_wall1 = time.time()
    launches = _lib.launch_count() - launches0
    ts.check_errors(ctx.local)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = pairs_total * args.steps / (total_ms * 1e-3)

    # ---- N > 1: the same matrix on ONE GPU of this box (rank 0 alone, the others wait), so that the strong-scaling factor can be read from
    # this record alone: value / (pairs / one_gpu_ms)
    one_gpu = None
    if n_gpus > 1 and not args.no_extra:
        one_ms = 0.0
        if ctx.rank == 0:
            full = {m: torch.empty(pairs_total, dtype=torch.float64, device=ctx.dev) for m in metrics}
            f = lambda: ts.pairwise_device(metrics, {m: full[m].data_ptr() for m in metrics}, stream=ctx.stream.cuda_stream)
            f(); f()
            torch.cuda.synchronize()
            ms = []
            for _ in range(3):
                ctx.flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(ctx.stream); f(); e1.record(ctx.stream)
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            one_ms = float(np.median(ms))
            del full
            torch.cuda.empty_cache()
        ctx.barrier()
        if ctx.rank == 0:
            one_gpu = {'ms_per_step': one_ms, 'pairs_per_s': pairs_total / (one_ms * 1e-3),
                       'what': 'the whole matrix of this workload on rank 0 alone, same box, same kernels (median of 3 steps)'}

    # ---- end to end through the public call, host buffers ------------------------------------------------------------------------
    host = [(a, b, {m: _lib.pinned_empty(max(_lib.condensed_offset(nt, b) - _lib.condensed_offset(nt, a), 1)) for m in metrics}) for a, b in blocks]

    def e2e_step():
        t, _ = ctx.share_treeset(encode)                          # parse + encode on rank 0 (+ serialise, broadcast, deserialise)
        for a, b, bufs in host:
            t.pairwise(metrics, row_begin=a, row_end=b, device=ctx.local, out=bufs)   # H2D (first call), kernels, banded D2H
        h2d = upload_bytes(t.info, nt, metrics)
        t.close()
        return h2d

    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 5))
    h2d_bytes, e2e_value = 0, None
    if e2e_steps:
        e2e_step()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            h2d_bytes = e2e_step()
        ctx.barrier()
        e2e_value = pairs_total * e2e_steps / ctx.max_over_ranks(time.perf_counter() - t0)

    # ---- the drop-in call itself (N = 1): Collection.get_inter_tree_distances, one metric per call as in the reference, N x N
    # DistanceMatrix returned (squareform on the device, DataFrame wrapped round the library's page-locked buffer)
    api = None
    if n_gpus == 1 and strings is not None and not args.no_e2e:
        import treecl_b200
        coll = treecl_b200.Collection(records=[treecl_b200.Record('g%05d' % k, ml_tree=t) for k, t in enumerate(strings)])
        api = {}
        for m in metrics:
            for _ in range(2):                                   # (warm the library's page-locked result pool: 200 MB per matrix)
                coll.get_inter_tree_distances(m, show_progress=False)
            t0 = time.perf_counter()
            for _ in range(3):
                dm = coll.get_inter_tree_distances(m, show_progress=False)
                del dm
            api[m + '_ms'] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
        api['what'] = 'Collection.get_inter_tree_distances(metric): encode + kernels + squareform on the device + D2H of the N x N matrix + DataFrame; wall clock per call'

    # ---- the other BASELINE configs on the same GPUs --------------------------------------------------------------------------------
    extra = {}
    if not args.no_extra:
        del outs, host
        ts.close()
        torch.cuda.empty_cache()
        plan = ([(2, {}), (3, {'u16': True, 'parity_rows': 2}), (4, {'parity_rows': 8})] if n_gpus == 1 else
                [(1, {'n_trees': int(round(5000 * math.sqrt(n_gpus))), 'weak': True}), (3, {'u16': True, 'parity_rows': 1}), (4, {'parity_rows': 4})])
        for idx, kw in plan:
            key = 'config{}_{}'.format(idx + 1, '+'.join(config_of(idx)['metrics']))
            try:
                res, t2, o2, h2 = run_config(ctx, idx, max(2, min(args.steps, 5)), 2, **kw)
                if idx == 3 and not args.no_e2e:
                    res['e2e'] = config4_e2e(ctx, t2, res['pairs'])
                extra[key] = res
                del o2, h2
                t2.close()
                _lib.trim_cache(ctx.local)
                torch.cuda.empty_cache()
            except Exception as err:          # noqa - an extra must not take the main line down
                extra[key] = {'error': '{}: {}'.format(type(err).__name__, str(err)[:300])}

    if ctx.rank == 0:
        n_splits = info.max_splits
        alg = algorithmic_bytes(nt, nx, n_splits, my_pairs, 8 * len(metrics))     # this rank's launches
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
            traffic = tr.get('pairwise_{}_{}x{}'.format('_'.join(metrics), nt, nx))
        except (OSError, ValueError):
            pass
        achieved = alg / kernel_s / 1e9
        threads = os.cpu_count() or 1
        cpu = None
        if n_gpus == 1 and not args.no_cpu_baseline and strings is not None:
            shaped, take_s, _ = cpu_reference(strings, metrics, 200000, threads, True)
            best, take_b, _ = cpu_reference(strings, metrics, min(pairs_total, 4000000), threads, False)
            cpu = {'value': shaped, 'unit': 'pairs/s', 'cores': threads, 'kind': 'port',
                   'sample': 'oracle port, reference-shaped (both Newick strings re-parsed per pair, treeCl/tasks.py:41-75), '
                             '{} threads, first {} pairs, {}'.format(threads, take_s, '+'.join(metrics)),
                   'best_cpu_value': best,
                   'best_cpu_sample': 'same oracle, trees encoded once, {} threads, first {} pairs'.format(threads, take_b)}
        line = {
            'metric': 'tree-pairs/sec ({} matrices)'.format('+'.join(metrics)), 'value': value, 'unit': 'pairs/s', 'n_gpus': n_gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak' if n_gpus == 1 else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': workload_name(cfg, nt),
                       'launches': 'one call per step: events kernel, one tile kernel per row block ({} per rank), exact fix-up'.format(len(blocks)),
                       'pairs': pairs_total, 'pairs_per_gpu': my_pairs,
                       'parallelism': 'folded row blocks x{} (rank r: a block of long rows + a block of short rows), set encoded once and broadcast'.format(n_gpus) if n_gpus > 1 else 'one GPU',
                       'l2': '256 MiB memset + 256 MiB read between timed steps (outside the event pairs): L2 cold and clean; outputs exceed L2',
                       'dict_size': int(info.dict_size), 'dict_shared': int(info.dict_shared), 'max_shared': int(info.max_shared),
                       'note': None if n_gpus == 1 else 'N > 1 measures configs[2] strong-scaled; the N = 1 line measures configs[1] - compare this value with '
                               'config.extra.config3_wrf+euc of the N = 1 line (same workload on one GPU), not with its main value',
                       'same_workload_on_one_gpu': one_gpu,
                       'extra': extra},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': traffic, 'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback',
                         'algorithmic_bytes_per_launch': alg, 'kernel': 'pair_dense_kernel<{}> (fp64 tensor-core tile kernel; the step also runs pair_events_kernel and exact_fixup_kernel, and the '
                                   'duration used here is the whole step of rank 0, so frac is a lower bound for the tile kernel)'.format('|'.join(metrics))},
            'cpu_baseline': cpu,
            'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': int(h2d_bytes + blob_bytes),
                    'd2h_bytes_per_step': int(my_pairs * 8 * len(metrics)),
                    'includes': 'Newick parse + split encode on host (rank 0){}, H2D, kernels, banded D2H into pinned host memory'.format(
                        ', serialise + NCCL broadcast + deserialise' if n_gpus > 1 else ''),
                    'api_call': api},
            'gpu_launches': int(launches), 'clocks': clocks,
            'step_ms': [round(x, 4) for x in step_ms],
        }
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()
    return 0


def config4_e2e(ctx, ts, pairs_total):
    """BASELINE configs[3] 'tiles streamed to pinned host': every rank streams its uint16 row blocks into page-locked host memory
    (td_rf_incidence: row bands, kernels under the copies).  Set already resident; wall clock, max over ranks."""
    from treecl_b200 import _lib
    n = ts.n_trees
    blocks = my_blocks(ctx, n, 512)
    if block_pairs(n, blocks) * 2 > (3 << 30):
        return {'skipped': 'this rank alone would need {:.1f} GB of page-locked host memory; measured at N >= 4'.format(block_pairs(n, blocks) * 2 / 1e9)}
    host = [(a, b, _lib.pinned_empty(max(_lib.condensed_offset(n, b) - _lib.condensed_offset(n, a), 1), np.uint16)) for a, b in blocks]

    def go():
        for a, b, buf in host:
            ts.rf_incidence(row_begin=a, row_end=b, device=ctx.local, out=buf)

    go()
    ctx.barrier()
    t0 = time.perf_counter()
    go()
    ctx.barrier()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    return {'pairs_per_s': pairs_total / dt, 'ms': dt * 1e3, 'd2h_bytes_this_rank': int(sum(h[2].nbytes for h in host)),
            'what': 'set resident -> last byte of the uint16 condensed blocks in pinned host memory'}


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
    ap.add_argument('--metrics', default=None, help='comma list overriding the config\'s metrics (experiments)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the other BASELINE configs (config.extra)')
    ap.add_argument('--config', type=int, default=None, help='BASELINE.json configs index of the main line (default: 1 at N = 1, 2 at N > 1)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)
    return run_ours(args)


if __name__ == '__main__':
    sys.exit(main())
