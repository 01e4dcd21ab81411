"""Developer probe (f3): batched per-proof PLONK scalar evaluation (snarkv_fr_program_eval_batch_device) on a StandardPlonk-shaped
protocol, inputs resident in HBM; CUDA events.  usage: plonk_eval_probe.py [k] [m,m,...]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import snark_verifier_b200 as sv
from snark_verifier_b200 import plonk_eval as pe

k = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ms_list = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,4096,65536,1048576").split(",")]
p = pe.standard_plonk_like_protocol(k, num_instance=1)
prog = pe.compile_quotient_evaluation(p)
tot = p.input_layout()["total"]
print("program: %d instructions %s, %d registers, %d inputs, %d outputs" % (len(prog.instrs), prog.op_histogram(), prog.n_regs, tot, len(prog.outputs)))
stream = torch.cuda.Stream()
L = sv.CudaLoader(0)
L.set_stream(stream.cuda_stream)
for m in ms_list:
    with torch.cuda.stream(stream):
        d_in = torch.empty(m * tot * 32, dtype=torch.uint8, device="cuda")
        d_out = torch.zeros(m * len(prog.outputs) * 32, dtype=torch.uint8, device="cuda")
        L.synth_scalars_device(9, 0, m * tot, d_in.data_ptr())
    ts = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            L.fr_program_eval(prog, None, m, d_inputs=d_in.data_ptr(), d_outputs=d_out.data_ptr())
            e1.record(stream)
        stream.synchronize()
        if rep >= 2:
            ts.append(e0.elapsed_time(e1))
    t = min(ts)
    n_mul = prog.op_histogram()["mul"]
    print("m=%8d  %.3f ms  %.2f M proofs/s  %.1f G Fr-mul/s  (%.0f bytes in+out per proof)" % (m, t, m / t / 1e3, m * n_mul / t / 1e6, (tot + len(prog.outputs)) * 32), flush=True)
L.close()
