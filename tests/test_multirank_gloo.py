"""N>1 host logic on CPU (gloo, world_size 2): the chunk partition + all-gather of per-rank partials + fold reproduces the
single-rank MSM.  On the GPU the per-rank partial comes from snarkv_g1_msm_partial and the fold from
snarkv_g1_fold_partials_device; here the CPU oracle stands in for both so that the sharding/collective logic is what is tested."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from snark_verifier_b200.sharding import chunk_bounds

N_TERMS = 1000
SEED = 3


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, cnt = chunk_bounds(N_TERMS, world, rank)
    if cnt:
        part = oracle.msm_pippenger(oracle.synth_scalars(SEED, lo, cnt), oracle.synth_points(SEED, lo, cnt, 1), cnt, 1)
    else:
        part = bytes(64)
    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8)
    gathered = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(gathered, mine)
    acc = bytes(64)
    for g in gathered:                                   # results.iter().fold(identity, |acc, r| acc + r)   util/msm.rs:333-335
        acc = oracle.g1_add(acc, bytes(g.numpy()))
    ret[rank] = acc
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_chunked_msm_with_allgather_fold(world):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29611 + world
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    full = oracle.msm_pippenger(oracle.synth_scalars(SEED, 0, N_TERMS), oracle.synth_points(SEED, 0, N_TERMS, 2), N_TERMS, 2)
    assert all(ret[r] == full for r in range(world))


def test_chunk_bounds_cover_all_terms():
    for n in (1, 2, 7, 1000, 1 << 24):
        for world in (1, 2, 3, 4, 8):
            spans = [chunk_bounds(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for lo, c in spans:
                assert lo == pos or c == 0
                pos += c
