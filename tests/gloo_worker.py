"""world_size-2 gloo worker for test_host_logic.py: LPT shards -> uneven all-gather -> gathered order ->
row blocks.  The 'signatures' are synthetic rows that encode their scaffold id, so the test checks the
plumbing (the real counting kernel is covered by the gpu tests)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from dbb_b200 import parallel


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    lengths = rng.randint(1000, 100000, 501)
    parts = parallel.lpt_partition(lengths, world)
    order, inv = parallel.gathered_order(parts)
    mine = parts[rank]
    local = torch.from_numpy(np.stack([mine.astype(np.float32) * 3 + c for c in range(5)], axis=1))
    full = parallel.all_gather_rows(local, [len(p) for p in parts])
    assert full.shape == (501, 5)
    assert np.array_equal(full[:, 0].numpy(), order.astype(np.float32) * 3), 'gathered order is rank-major'
    # scalars permuted on the host line up with the gathered rows
    assert np.array_equal((full[:, 1].numpy() - 1) / 3, order)
    # the scalar hand-off: every rank contributes the metadata of its own shard only
    gc = rng.uniform(0.2, 0.8, 501)
    cov = rng.lognormal(3.0, 1.0, 501)
    ids, ln, g, c = parallel.all_gather_scalars(mine, lengths[mine], gc[mine], cov[mine], counts=[len(p) for p in parts])
    assert np.array_equal(ids, order) and np.array_equal(ln, lengths[order])
    assert np.array_equal(g, gc[order]) and np.array_equal(c, cov[order])
    r0, r1 = parallel.row_block(501, rank, world)
    rows = torch.tensor([float(r1 - r0)])
    dist.all_reduce(rows)
    assert int(rows.item()) == 501
    dist.barrier()
    dist.destroy_process_group()
    print('gloo ok rank %d' % rank)


if __name__ == '__main__':
    sys.exit(main())
