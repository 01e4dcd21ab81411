"""compute-sanitizer target (developer tool): every kernel of the library once at small sizes, results checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
import snark_verifier_b200 as sv
from oracle import bn254_model as m
le = m.fe_to_le
L = sv.CudaLoader(0)
for n in (1, 3, 300, 5000):
    s = oracle.synth_scalars(3, 0, n); p = oracle.synth_points(3, 0, n, 4)
    assert L.msm(s, p, n, flags=sv.CHECK_INPUTS) == oracle.msm_pippenger(s, p, n, 4), n
s = le(1) * 3000; p = oracle.synth_points(4, 0, 3000, 4)
assert L.msm(s, p, 3000) == oracle.msm_pippenger(s, p, 3000, 4)                      # task split + block merge
offs = [0, 21, 24, 44]
s = oracle.synth_scalars(5, 0, 44); p = oracle.synth_points(5, 0, 44, 4)
L.msm_batch(s, p, offs); L.msm_batch_rlc(s, p, offs, le(12345))
ds = torch.empty(2000 * 32, dtype=torch.uint8, device="cuda"); dp = torch.empty(2000 * 64, dtype=torch.uint8, device="cuda")
L.synth_scalars_device(1, 0, 2000, ds.data_ptr()); L.synth_points_device(1, 0, 2000, dp.data_ptr())
part = torch.zeros(192, dtype=torch.uint8, device="cuda"); out = torch.zeros(64, dtype=torch.uint8, device="cuda")
L.msm_device(ds.data_ptr(), dp.data_ptr(), 1000, d_out_jacobian=part.data_ptr())
L.msm_device(ds.data_ptr() + 32000, dp.data_ptr() + 64000, 1000, d_out_jacobian=part.data_ptr() + 96)
L.fold_partials_device(part.data_ptr(), 2, out.data_ptr()); torch.cuda.synchronize()
g2 = oracle.g2_generator(); sk = 77; s_g2 = oracle.g2_mul(g2, le(sk)); gen = m.g1_to_bytes(m.G1_GEN)
kz = sv.KzgAs(L, sv.KzgDecidingKey(gen, g2, s_g2))
lhs = b"".join(oracle.g1_mul(gen, le(a * sk + (a == 3))) for a in range(1, 6)); rhs = b"".join(oracle.g1_mul(gen, le(a)) for a in range(1, 6))
for mode in (1, 3, 4):
    L.set_pairing_mode(mode)
    acc, gt = kz.decide_batch(lhs, rhs, 5, want_gt=True)
    assert acc == b"\x01\x01\x00\x01\x01", (mode, acc)
L.set_pairing_mode(0)
accs = [sv.KzgAccumulator(lhs[64 * i:64 * i + 64], rhs[64 * i:64 * i + 64]) for i in (0, 1, 3, 4)]
kz.decide(kz.verify(accs, le(99)))
ok, _ = kz.decide_all_fused(b"".join(a.lhs for a in accs), b"".join(a.rhs for a in accs), 4, le(31337))
assert ok
L.field_op(0, 3, le(5) * 4, le(5) * 4, 4)
# batched-affine accumulation (forced: the sizes above choose XYZZ), with the task-level self-check against XYZZ (mode 3)
for mode in (2, 3):
    L.set_accumulate_mode(mode)
    n = 6000
    s = oracle.synth_scalars(8, 0, n); p = oracle.synth_points(8, 0, n, 4)
    L.set_window_bits(6)                                                               # 32 buckets per window: long lists
    assert L.msm(s, p, n) == oracle.msm_pippenger(s, p, n, 4), mode
    L.set_window_bits(0)
L.set_accumulate_mode(0)
# rows f1-f3, a13: scalar preparation, transcript, scalar-evaluation program, limb decoding
L.powers(le(7), 100); L.batch_invert(le(3) * 50 + le(0) * 2, 52); L.fr_mul_vec(le(3) * 10, le(4) * 10, 10)
L.evm_transcript_challenges(bytes(range(64)) * 3, 64, [32, 64], 3)
from snark_verifier_b200 import pcs, plonk_eval as pe
proto = pe.standard_plonk_like_protocol(6, num_instance=2)
prog = pe.compile_quotient_evaluation(proto)
tot = proto.input_layout()["total"]
L.fr_program_eval(prog, oracle.synth_scalars(6, 0, 70 * tot), 70)
enc = pcs.LimbsEncoding(4, 68)
rows = enc.to_repr(accs[0]) + enc.to_repr(accs[1])
rows[3] ^= 1
_, _, valid = enc.from_repr_batch(L, b"".join(le(v) for v in rows), 2)
assert valid == b"\x00\x01"
print("sanitizer target ok; launches:", L.launch_count)
