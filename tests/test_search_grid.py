"""The rank grid of the configs[2]-size search (bench_dist.py): reference shards x query groups cover every (query, reference) pair
exactly once at every world size -- host logic only, no GPU."""
import numpy as np
import pytest

from public_kssd_b200 import parallel


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_grid_covers_every_pair_once(world):
    n_ref, n_qry = 1003, 97
    for Gr in sorted({1, 2, world} & {g for g in (1, 2, world) if world % g == 0}):
        Gq = world // Gr
        seen = np.zeros((n_qry, n_ref), dtype=np.uint8)
        for rank in range(world):
            gr, gq = rank % Gr, rank // Gr
            r = parallel.genome_shard(n_ref, Gr, gr)
            q = parallel.genome_shard(n_qry, Gq, gq)
            seen[q.start:q.stop, r.start:r.stop] += 1
        assert (seen == 1).all(), (world, Gr)


def test_code_ranges_partition_the_code_space():
    for world in (1, 2, 3, 8):
        edges = [parallel.code_range(r, world, 28) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 1 << 28
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
