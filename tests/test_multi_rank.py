"""World-size-2/3 gloo tests of the multi-GPU plan (CPU only).  (1) every rank takes a balanced row range of the upper triangle,
fills its slice of the condensed vector, and the slices tile the whole vector with no gaps or overlaps; (2) bench.py's plan: rank 0
encodes the set once, its serialised bytes are broadcast, every rank deserialises and takes its FOLDED row blocks.  The per-rank
"device" here is the CPU oracle (tests may use it); on the GPU box the same plan drives td_pairwise_device per rank
(bench.py, tests/test_gpu_parity.py::test_row_sharding_is_byte_identical)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]; s.close()
    return port


def _worker(rank, world, port, n_trees, n_taxa, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle
    from treecl_b200 import _lib, engine, synth
    trees = synth.uniform_trees(n_trees, n_taxa, seed=5)            # every rank builds the same replicated set
    r0, r1 = engine.balanced_row_ranges(n_trees, world)[rank]
    p0, p1 = _lib.condensed_offset(n_trees, r0), _lib.condensed_offset(n_trees, r1)
    mine = oracle.matrix_range(trees, 'euc', p0, p1)
    # the timing reduction bench.py uses: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    # no collective on the data path: each rank writes its own slice (here: a file per rank); rank 0 checks the tiling
    np.save(os.path.join(out_dir, 'slice_%d.npy' % rank), np.concatenate([[p0, p1], mine]))
    dist.barrier()
    if rank == 0:
        full = oracle.matrix(trees, 'euc')
        pos, parts = 0, []
        for r in range(world):
            a = np.load(os.path.join(out_dir, 'slice_%d.npy' % r))
            assert int(a[0]) == pos                                   # contiguous, no gap / overlap
            pos = int(a[1]); parts.append(a[2:])
        assert pos == full.size
        assert np.array_equal(np.concatenate(parts), full)
        counts = [p.size for p in parts]
        assert max(counts) - min(counts) <= 2 * n_trees               # balanced pair counts
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_row_sharding_plan_gloo(tmp_path, world):
    import oracle
    oracle.build()
    mp.spawn(_worker, args=(world, _free_port(), 41, 9, str(tmp_path)), nprocs=world, join=True)


def _worker_broadcast(rank, world, port, n_trees, n_taxa, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import oracle
    from treecl_b200 import _lib, engine, synth
    trees = synth.uniform_trees(n_trees, n_taxa, seed=6)
    # encode once on rank 0, broadcast the serialised set (bench.py does the same over NCCL)
    size = torch.zeros(1, dtype=torch.int64)
    if rank == 0:
        blob = torch.from_numpy(_lib.TreeSet(trees).serialize())
        size[0] = blob.numel()
    dist.broadcast(size, 0)
    if rank != 0:
        blob = torch.empty(int(size.item()), dtype=torch.uint8)
    dist.broadcast(blob, 0)
    ts = _lib.TreeSet.deserialize(blob.numpy())
    assert ts.n_trees == n_trees and ts.info.n_taxa == n_taxa
    # the deserialised encoding is the encoding: masks / lengths equal what a local encode gives
    local = _lib.TreeSet(trees)
    for what, dt in ((_lib.TD_GET_MASKS, np.uint64), (_lib.TD_GET_LENGTHS, np.float64), (_lib.TD_GET_IDS, np.uint32), (_lib.TD_GET_LEAFLEN, np.float64)):
        assert np.array_equal(ts.get(what, dt), local.get(what, dt))
    # folded row blocks; the per-rank "device" is the oracle
    mine = []
    for a, b in engine.folded_row_ranges(n_trees, world, align=8)[rank]:
        p0, p1 = _lib.condensed_offset(n_trees, a), _lib.condensed_offset(n_trees, b)
        mine.append(np.concatenate([[p0, p1], oracle.matrix_range(trees, 'rf', p0, p1)]))
    np.save(os.path.join(out_dir, 'fold_%d.npy' % rank), np.concatenate(mine) if mine else np.zeros(0))
    dist.barrier()
    if rank == 0:
        full = oracle.matrix(trees, 'rf')
        got = np.full(full.size, -1.0)
        counts = []
        for r in range(world):
            a = np.load(os.path.join(out_dir, 'fold_%d.npy' % r))
            at, n_r = 0, 0
            while at < a.size:
                p0, p1 = int(a[at]), int(a[at + 1])
                assert (got[p0:p1] == -1).all()                       # no overlap
                got[p0:p1] = a[at + 2:at + 2 + p1 - p0]
                at += 2 + p1 - p0; n_r += p1 - p0
            counts.append(n_r)
        assert np.array_equal(got, full)                              # no gap
        assert max(counts) <= 1.15 * (full.size / world)             # balanced (small set, coarse alignment)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_encode_once_broadcast_and_folded_blocks_gloo(tmp_path, world):
    import oracle
    oracle.build()
    mp.spawn(_worker_broadcast, args=(world, _free_port(), 97, 11, str(tmp_path)), nprocs=world, join=True)
