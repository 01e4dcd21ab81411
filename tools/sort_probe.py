"""Developer probe: sort-phase stage times (digits / partition / buckets) for SNARKV_SORT_TILE = 4096 | 8192."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snark_verifier_b200 as sv
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,22,24").split(",")]
nmax = 1 << max(sizes)
stream = torch.cuda.Stream()
res = {}
for tile, blocks in ((4096, 3), (4096, 4), (4096, 6), (8192, 2), (8192, 3), (8192, 4)):
    os.environ["SNARKV_SORT_TILE"] = str(tile)
    os.environ["SNARKV_SORT_BLOCKS"] = str(blocks)
    L = sv.CudaLoader(0, fmt=sv.MONTGOMERY)
    L.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        ds = torch.empty(nmax * 32, dtype=torch.uint8, device="cuda")
        dp = torch.empty(nmax * 64, dtype=torch.uint8, device="cuda")
        out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        L.synth_scalars_device(5, 0, nmax, ds.data_ptr())
        L.synth_points_device(5, 0, nmax, dp.data_ptr())
    stream.synchronize()
    for lg in sizes:
        n = 1 << lg
        best = {}
        for rep in range(4):
            L.profile(True)
            with torch.cuda.stream(stream):
                L.msm_device(ds.data_ptr(), dp.data_ptr(), n, d_out_affine=out.data_ptr())
            stream.synchronize()
            for a, b, _ in L.stage_times():
                if rep and ("sort" in a or "digits" in a):
                    best[a] = min(best.get(a, 1e9), b)
        res[(tile, lg)] = bytes(out.cpu().numpy())
        print("tile=%d blocks/SM=%d n=2^%d " % (tile, blocks, lg) + " ".join("%s=%.3f" % kv for kv in best.items()), flush=True)
    L.close()
print("results equal:", all(res[(4096, lg)] == res[(8192, lg)] for lg in sizes))
