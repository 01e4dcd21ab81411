import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import snark_verifier_b200 as sv
from snark_verifier_b200 import plonk
import plonk_toy as T
import oracle
srs = T.Srs(3); circuit = T.Circuit(4, 11, [5, 7])
L = sv.CudaLoader(0); kz = sv.KzgAs(L, sv.KzgDecidingKey(T.GEN, srs.g2, srs.s_g2))
protocol = T.make_protocol(circuit, srs, None)
bv = plonk.PlonkBatchVerifier(L, kz, T.GEN, protocol, "gwc19")
proof = T.prove(circuit, protocol, srs, "gwc19")
inst = [circuit.public]
a1 = bv.accumulate_new([inst], [proof], 1)
a2 = bv.accumulate_new([inst], [proof], 1)
kz.decide_all([a1])
a3 = bv.accumulate_new([inst], [proof], 1)
bv.use_device_plan = False
py = bv.accumulate_new([inst], [proof], 1)
bv.use_device_plan = True
a4 = bv.accumulate_new([inst], [proof], 1)
print("a1==py", a1.lhs == py.lhs, "a2==py", a2.lhs == py.lhs, "a3(after decide)==py", a3.lhs == py.lhs, "a4(after python path)==py", a4.lhs == py.lhs)
two = oracle.g1_add(py.lhs, py.lhs)
print("a3 == 2*py:", a3.lhs == two, " a2 == 2*py:", a2.lhs == two)
# fresh verifier object on the same loader (new plan)
bv2 = plonk.PlonkBatchVerifier(L, kz, T.GEN, protocol, "gwc19")
b1 = bv2.accumulate_new([inst], [proof], 1)
print("second verifier object b1==py", b1.lhs == py.lhs)
v = plonk.PlonkVerifier(L, kz, T.GEN, protocol, "gwc19")
v.verify(inst, proof)
g = v.succinct_verify(inst, proof)[0]
print("PlonkVerifier succinct after verify == py", g.lhs == py.lhs)
