"""The N>1 host path on CPU: two gloo ranks render their sample shards (with the oracle as the stand-in
renderer, this test is about sharding + reduce + rescale), reduce to rank 0, and must reproduce the
single-process film."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))
    import torch
    import torch.distributed as dist
    from oracle import bindings as ob
    from lmb200py import scenedesc, distributed
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = scenedesc.cornell_box(24, 24)
    P = ob.PortPT(sc)
    b, e = distributed.shard_range(n, rank, world)
    # unscaled film of this rank's slice (PortPT.render scales by W*H/n: undo it)
    img, _ = P.render(1, n, seed=5, begin=b, end=e)
    film = torch.zeros((24, 24, 4), dtype=torch.float32)
    film[..., :3] = torch.from_numpy(img / np.float32(distributed.film_scale(24, 24, n)))
    distributed.reduce_film(film, dist)
    if rank == 0:
        np.save(out_path, (film[..., :3] * distributed.film_scale(24, 24, n)).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_tiles_exactly():
    sys.path.insert(0, os.path.join(ROOT, "lightmetrica-v2_b200"))
    from lmb200py import distributed
    for n in (0, 1, 7, 1000, 2**33 + 5):
        for w in (1, 2, 3, 8):
            edges = [distributed.shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        distributed.shard_range(10, 2, 2)


def test_two_rank_reduce_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import bindings as ob
    from lmb200py import scenedesc
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n = 24 * 24 * 32
    out = str(tmp_path / "film.npy")
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = np.load(out)
    ref, _ = ob.PortPT(scenedesc.cornell_box(24, 24)).render(1, n, seed=5)
    assert np.allclose(got, ref, rtol=1e-4, atol=1e-5)
