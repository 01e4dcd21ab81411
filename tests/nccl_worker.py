"""Worker of tests/test_gpu_multirank_nccl.py: one rank per GPU under torchrun.  Every rank runs the CUDA pipeline on its chunk
(snarkv_g1_msm_device -> 96-byte Jacobian partial), the partials are all-gathered over NCCL, every rank folds them
(snarkv_g1_fold_partials_device) and compares the affine result with the CPU oracle's MSM over the WHOLE input.  The pairing
batch is sharded check i -> rank i mod G with an all-gather of the accept bytes."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m
from snark_verifier_b200.sharding import chunk_bounds


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = sv.CudaLoader(local)
    # one explicit stream for torch, NCCL and the library (as bench.py does): the legacy default stream handle is 0, which
    # snarkv_set_stream reads as "use the context's own stream" — the library would then race with torch's fills and NCCL
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    L.set_stream(stream.cuda_stream)
    ok = True
    for n in (7, 1000, (1 << 15) + 3):
        s, p = oracle.synth_scalars(77, 0, n), oracle.synth_points(77, 0, n, 2)
        lo, cnt = chunk_bounds(n, world, rank)
        part = torch.zeros(96, dtype=torch.uint8, device=dev)
        if cnt:
            ds = torch.frombuffer(bytearray(s[32 * lo:32 * (lo + cnt)]), dtype=torch.uint8).to(dev)
            dp = torch.frombuffer(bytearray(p[64 * lo:64 * (lo + cnt)]), dtype=torch.uint8).to(dev)
            L.msm_device(ds.data_ptr(), dp.data_ptr(), cnt, d_out_jacobian=part.data_ptr())
        else:
            part[32] = 1                                 # Jacobian identity (0, 1, 0); Z == 0 is what counts
        parts = torch.zeros(96 * world, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(parts, part)
        out = torch.zeros(64, dtype=torch.uint8, device=dev)
        L.fold_partials_device(parts.data_ptr(), world, out.data_ptr())
        got = bytes(out.cpu().numpy())
        exp = oracle.msm_pippenger(s, p, n, 2)
        if got != exp:
            ok = False
            print("rank %d: MSM mismatch at n=%d" % (rank, n), flush=True)
    # pairing batch: check i -> rank i mod G, accept bytes all-gathered
    g2 = oracle.g2_generator()
    sk = 0x5EC2E7
    s_g2 = oracle.g2_mul(g2, m.fe_to_le(sk))
    gen = m.g1_to_bytes(m.G1_GEN)
    kz = sv.KzgAs(L, sv.KzgDecidingKey(gen, g2, s_g2))
    N = 8 * world
    lhs = [oracle.g1_mul(gen, m.fe_to_le((3 + i) * sk + (1 if i % 3 == 1 else 0))) for i in range(N)]
    rhs = [oracle.g1_mul(gen, m.fe_to_le(3 + i)) for i in range(N)]
    mine = list(range(rank, N, world))
    acc, _ = kz.decide_batch(b"".join(lhs[i] for i in mine), b"".join(rhs[i] for i in mine), len(mine))
    t = torch.frombuffer(bytearray(acc), dtype=torch.uint8).to(dev)
    allacc = torch.zeros(N, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allacc, t)
    allacc = allacc.view(world, N // world).t().reshape(-1).cpu().numpy()      # back to check order
    exp = bytes(0 if i % 3 == 1 else 1 for i in range(N))
    if bytes(allacc) != exp:
        ok = False
        print("rank %d: pairing shard mismatch" % rank, flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    L.close()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_WORKER_OK" if int(flag.item()) == 1 else "NCCL_WORKER_FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
