"""SURVEY §8 f3 (third "next" row): per-proof PLONK scalar evaluation — `CommonPolynomialEvaluation` (protocol.rs:211-283),
`Expression::evaluate` (protocol.rs:336-392), instance / quotient evaluations (proof.rs:298-349) — compiled once into a
straight-line Fr program and run for a batch of proofs by `snarkv_fr_program_eval_batch`.

CPU tests: the host compiler (snark_verifier_b200/plonk_eval.py) against the big-integer oracle (oracle/plonk_eval_model.py) through
the oracle's own program interpreter.  GPU tests: the device kernel against the oracle, bit-exact, through the C ABI."""
import random

import numpy as np
import pytest

from oracle import plonk_eval_model as om
from snark_verifier_b200 import plonk_eval as pe

R = om.R


def random_expression(rng, depth, n_polys, n_chal, lagranges):
    E = pe.Expression
    if depth == 0 or rng.random() < 0.15:
        k = rng.randrange(4)
        if k == 0:
            return E.constant(rng.choice([0, 1, 2, R - 1, rng.randrange(R)]))
        if k == 1:
            return E.common_polynomial(pe.CommonPolynomial.identity() if rng.random() < 0.4 else pe.CommonPolynomial.lagrange(rng.choice(lagranges)))
        if k == 2:
            return E.polynomial(pe.Query(rng.randrange(n_polys), pe.Rotation(rng.choice([-1, 0, 0, 1, 2]))))
        return E.challenge(rng.randrange(n_chal))
    k = rng.randrange(6)
    sub = lambda: random_expression(rng, depth - 1, n_polys, n_chal, lagranges)
    if k == 0:
        return E.negated(sub())
    if k == 1:
        return E.sum(sub(), sub())
    if k == 2:
        return E.product(sub(), sub())
    if k == 3:
        return E.scaled(sub(), rng.randrange(R))
    if k == 4:
        return sub() - sub()
    return E.distribute_powers([sub() for _ in range(rng.randrange(1, 5))], sub())


def random_protocol(rng, k=6):
    """No instance columns: every query of the numerator is an evaluation the proof carries."""
    num = random_expression(rng, 6, n_polys=5, n_chal=3, lagranges=[-3, -1, 0, 1, 4])
    evals = sorted(num.used_query())
    return pe.QuotientProtocol(domain=pe.Domain(k), num_preprocessed=5, num_instance=[], evaluations=evals, num_challenge=3, numerator=num)


def oracle_row(p, row):
    lay = p.input_layout()
    inst, off = [], lay["instances"]
    for cnt in p.num_instance:
        inst.append(row[off:off + cnt])
        off += cnt
    return om.quotient_evaluation(p.domain.k, p.num_preprocessed, p.num_instance, [(q.poly, q.rotation.value) for q in p.evaluations],
                                  p.numerator.to_tuple(), row[0], row[1:1 + p.num_challenge], row[lay["evaluations"]:lay["instances"]], inst)


def rows_for(p, rng, m):
    tot = p.input_layout()["total"]
    rows = [[rng.randrange(R) for _ in range(tot)] for _ in range(m)]
    if m > 2:
        rows[1][0] = pe.root_of_unity(p.domain.k)     # z on the domain: z^n - 1 = 0 and one z - omega^i = 0 (zero stays zero)
        rows[2][0] = 0
    return rows


# ---- CPU: host logic ---------------------------------------------------------------------------------------------------------
def test_domain_and_constants():
    assert pe.FR_ROOT_OF_UNITY == pow(7, (R - 1) >> 28, R)
    for k in (1, 8, 20, 28):
        w = pe.root_of_unity(k)
        assert pow(w, 1 << k, R) == 1 and pow(w, 1 << (k - 1), R) == R - 1 and w == om.root_of_unity(k)
    d = pe.Domain(8)
    assert d.n * d.n_inv % R == 1 and d.gen * d.gen_inv % R == 1
    assert d.rotate_scalar(5, pe.Rotation(3)) == 5 * pow(d.gen, 3, R) % R
    assert d.rotate_scalar(5, pe.Rotation(-2)) == 5 * pow(d.gen_inv, 2, R) % R == 5 * om.rotate(8, -2) % R


def test_pow_const_is_the_reference_sequence():
    for exp in (1, 2, 3, 5, 6, 255, 256, 1 << 20, (1 << 20) + 77):
        b = pe.ProgramBuilder()
        out = b.pow_const(b.input(0), exp)
        prog = b.finish([out])
        x = 0x1234567 + exp
        assert om.run_program(prog.instrs, prog.n_regs, prog.consts, [x], prog.outputs) == [pow(x, exp, R)]
        # loader.rs:52-69: trailing zeros squarings, then one squaring per remaining bit and one product per set bit
        tz = (exp & -exp).bit_length() - 1
        rest = exp >> tz
        assert prog.op_histogram()["mul"] == tz + (rest.bit_length() - 1) + (bin(rest).count("1") - 1)


def test_standard_plonk_program_matches_oracle():
    rng = random.Random(11)
    for k, ninst in ((4, 1), (8, 3), (17, 2)):
        p = pe.standard_plonk_like_protocol(k, num_instance=ninst)
        prog = pe.compile_quotient_evaluation(p)
        assert prog.n_inputs == p.input_layout()["total"]
        assert prog.n_regs < 48, "liveness-based register reuse keeps the register file small"
        assert prog.op_histogram()["inv"] == 1, "all denominators of a proof share ONE inversion (util/arithmetic.rs:47-69)"
        for row in rows_for(p, rng, 6):
            assert om.run_program(prog.instrs, prog.n_regs, prog.consts, row, prog.outputs) == oracle_row(p, row)


def test_random_expression_programs_match_oracle():
    rng = random.Random(5)
    for _ in range(25):
        p = random_protocol(rng)
        prog = pe.compile_quotient_evaluation(p)
        for row in rows_for(p, rng, 4):
            assert om.run_program(prog.instrs, prog.n_regs, prog.consts, row, prog.outputs) == oracle_row(p, row)


def test_register_allocation_on_random_programs():
    """ProgramBuilder.finish maps SSA values onto registers by liveness; whatever the dataflow, the allocated program must compute
    what the SSA program computes (values read after their register was recycled would show up here)."""
    rng = random.Random(77)
    for trial in range(60):
        b = pe.ProgramBuilder()
        vals = [b.input(i) for i in range(4)] + [b.const(rng.choice([0, 1, 5, R - 1]))]
        ref = None
        for _ in range(rng.randrange(5, 80)):
            k = rng.randrange(8)
            x, y = rng.choice(vals), rng.choice(vals)
            if k == 0:
                vals.append(b.add(x, y))
            elif k == 1:
                vals.append(b.sub(x, y))
            elif k in (2, 3):
                vals.append(b.mul(x, y))
            elif k == 4:
                vals.append(b.neg(x))
            elif k == 5:
                vals.append(b.inv(x))
            elif k == 6:
                vals.extend(b.batch_invert(rng.sample(vals, rng.randrange(1, min(5, len(vals)) + 1))))
            else:
                vals.append(b.pow_const(x, rng.randrange(1, 40)))
        outs = rng.sample(vals, rng.randrange(1, 6))
        prog = b.finish(outs)
        row = [rng.choice([0, 1, rng.randrange(R)]) for _ in range(4)]
        # direct SSA evaluation (one slot per value, never recycled)
        ssa = []
        for op, a, c in b.ssa:
            if op == pe.OP_INPUT:
                ssa.append(row[a] % R)
            elif op == pe.OP_CONST:
                ssa.append(b.consts[a])
            elif op == pe.OP_ADD:
                ssa.append((ssa[a] + ssa[c]) % R)
            elif op == pe.OP_SUB:
                ssa.append((ssa[a] - ssa[c]) % R)
            elif op == pe.OP_MUL:
                ssa.append(ssa[a] * ssa[c] % R)
            elif op == pe.OP_NEG:
                ssa.append((-ssa[a]) % R)
            elif op == pe.OP_INV:
                ssa.append(om.inv_or_zero(ssa[a]))
            elif op == pe.OP_NZ:
                ssa.append(ssa[a] if ssa[a] else 1)
            else:
                ssa.append(ssa[a] if ssa[c] else 0)
        assert om.run_program(prog.instrs, prog.n_regs, prog.consts, row, prog.outputs) == [ssa[o] for o in outs], trial
        assert prog.n_regs <= len(b.ssa)


def test_batch_invert_matches_individual_inverses_with_zeros():
    """util/arithmetic.rs:47-74 as straight-line code: zeros stay zero and do not poison the shared inversion."""
    for values in ([3, 0, 7], [0, 0], [0], [5], [1, 2, 3, 4, 0, 6]):
        b = pe.ProgramBuilder()
        ins = [b.input(i) for i in range(len(values))]
        prog = b.finish(b.batch_invert(ins))
        assert prog.op_histogram()["inv"] == 1
        assert om.run_program(prog.instrs, prog.n_regs, prog.consts, values, prog.outputs) == [om.inv_or_zero(v) for v in values]


def test_missing_query_and_challenge_are_errors():
    E = pe.Expression
    p = pe.QuotientProtocol(pe.Domain(4), 1, [], [], 1, E.polynomial(pe.Query(0)))
    with pytest.raises(KeyError):       # Error::InvalidProtocol("Missing query ..") (proof.rs:232)
        pe.compile_quotient_evaluation(p)
    p = pe.QuotientProtocol(pe.Domain(4), 1, [], [], 1, E.challenge(1))
    with pytest.raises(KeyError):       # Error::InvalidProtocol("Missing challenge ..") (proof.rs:239)
        pe.compile_quotient_evaluation(p)


def test_used_sets_and_distribute_powers_single():
    E, CP, Q = pe.Expression, pe.CommonPolynomial, pe.Query
    e = E.distribute_powers([E.polynomial(Q(2)) * E.common_polynomial(CP.lagrange(-1))], E.challenge(0))
    assert e.used_langrange() == {-1} and e.used_query() == {Q(2)}
    assert om.eval_expr(e.to_tuple(), 3, {-1: 7}, {(2, 0): 5}, [9]) == 35     # a single expression ignores the scalar (protocol.rs:381-383)


def _tup(x):
    return tuple(_tup(e) for e in x) if isinstance(x, list) else x


def golden_cases(golden):
    g = golden("plonk_eval")
    for case in g["cases"]:
        p = pe.standard_plonk_like_protocol(case["k"], num_instance=case["num_instance"], blinding_factors=g["blinding_factors"])
        rows = [[int(v, 16) for v in r["inputs"]] for r in case["rows"]]
        exp = [[int(v, 16) for v in r["outputs"]] for r in case["rows"]]
        yield g, p, rows, exp


def test_golden_fixture_pins_structure_and_values(golden):
    """tests/golden/plonk_eval.json (oracle/gen_golden_plonk.py): the numerator written out there as plain tuples is the one the
    product builds, and the compiled program reproduces the committed evaluations."""
    for g, p, rows, exp in golden_cases(golden):
        assert p.numerator.to_tuple() == _tup(g["numerator"])
        assert [[q.poly, q.rotation.value] for q in p.evaluations] == g["evaluation_queries"]
        prog = pe.compile_quotient_evaluation(p)
        for row, e in zip(rows, exp):
            assert om.run_program(prog.instrs, prog.n_regs, prog.consts, row, prog.outputs) == e


def test_cpp_compiler_emits_the_same_program(tmp_path):
    """snark_verifier_b200/host/plonk_eval.hpp (the C++ host mirror: Expression, CommonPolynomialEvaluation, ProgramBuilder with
    the same liveness-based register allocation) must emit, instruction for instruction, the program the Python compiler emits."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "plonk_compile_test"
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", str(exe), os.path.join(root, "tests", "plonk_compile_test.cpp")])
    for k, ninst in ((4, 1), (8, 3), (12, 2), (17, 2)):
        out = subprocess.run([str(exe), str(k), str(ninst)], capture_output=True, text=True, check=True).stdout.splitlines()
        prog = pe.compile_quotient_evaluation(pe.standard_plonk_like_protocol(k, num_instance=ninst))
        hdr = out[0].split()
        assert (int(hdr[1]), int(hdr[3])) == (prog.n_regs, prog.n_inputs)
        assert [tuple(int(x) for x in l.split()[1:]) for l in out if l.startswith("i ")] == [tuple(i) for i in prog.instrs]
        assert [int(l.split()[1], 16) for l in out if l.startswith("c ")] == prog.consts
        assert [int(x) for x in [l for l in out if l.startswith("o")][0].split()[1:]] == prog.outputs


# ---- GPU: the device kernel, bit-exact -------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_device_matches_golden_fixture(golden):
    import snark_verifier_b200 as sv
    L = sv.CudaLoader(0)
    try:
        for g, p, rows, exp in golden_cases(golden):
            prog = pe.compile_quotient_evaluation(p)
            buf = b"".join(v.to_bytes(32, "little") for row in rows for v in row)
            out = L.fr_program_eval(prog, buf, len(rows))
            got = [int.from_bytes(out[i:i + 32], "little") for i in range(0, len(out), 32)]
            assert got == [v for e in exp for v in e]
    finally:
        L.close()



def pack_rows(rows, montgomery=False):
    conv = (lambda v: (v << 256) % R) if montgomery else (lambda v: v)
    return b"".join(conv(v).to_bytes(32, "little") for row in rows for v in row)


def unpack(buf, n_out, montgomery=False):
    rinv = pow(1 << 256, -1, R)
    vals = [int.from_bytes(buf[i:i + 32], "little") for i in range(0, len(buf), 32)]
    if montgomery:
        vals = [v * rinv % R for v in vals]
    return [vals[i:i + n_out] for i in range(0, len(vals), n_out)]


@pytest.mark.gpu
@pytest.mark.parametrize("fmt_mont", [False, True])
def test_device_standard_plonk_batch_matches_oracle(fmt_mont):
    import snark_verifier_b200 as sv
    rng = random.Random(21)
    L = sv.CudaLoader(0, fmt=sv.MONTGOMERY if fmt_mont else sv.CANONICAL)
    try:
        for k, ninst, m in ((8, 1, 1), (8, 3, 33), (12, 2, 300)):
            p = pe.standard_plonk_like_protocol(k, num_instance=ninst)
            prog = pe.compile_quotient_evaluation(p)
            rows = rows_for(p, rng, m)
            got = unpack(L.fr_program_eval(prog, pack_rows(rows, fmt_mont), m), len(prog.outputs), fmt_mont)
            assert got == [oracle_row(p, row) for row in rows]
    finally:
        L.close()


@pytest.mark.gpu
def test_device_random_programs_match_oracle():
    import snark_verifier_b200 as sv
    rng = random.Random(31)
    L = sv.CudaLoader(0)
    try:
        for _ in range(12):
            p = random_protocol(rng)
            prog = pe.compile_quotient_evaluation(p)
            rows = rows_for(p, rng, 40)
            got = unpack(L.fr_program_eval(prog, pack_rows(rows), 40), len(prog.outputs))
            assert got == [oracle_row(p, row) for row in rows]
    finally:
        L.close()


@pytest.mark.gpu
def test_device_resident_variant_and_large_batch_property():
    """4096 proofs with inputs and outputs resident in HBM; checked against the oracle on a sample and, for every proof, through
    the identity quotient_eval * (z^n - 1) == numerator-independent outputs' consistency zn_minus_one * zn_minus_one_inv == 1."""
    import torch

    import snark_verifier_b200 as sv
    rng = random.Random(41)
    L = sv.CudaLoader(0)
    try:
        m = 4096
        p = pe.standard_plonk_like_protocol(10, num_instance=2)
        prog = pe.compile_quotient_evaluation(p)
        tot = p.input_layout()["total"]
        d_in = torch.empty(m * tot * 32, dtype=torch.uint8, device="cuda")
        L.synth_scalars_device(77, 0, m * tot, d_in.data_ptr())            # uniformly random canonical scalars
        d_out = torch.zeros(m * len(prog.outputs) * 32, dtype=torch.uint8, device="cuda")
        L.fr_program_eval(prog, None, m, d_inputs=d_in.data_ptr(), d_outputs=d_out.data_ptr())
        torch.cuda.synchronize()
        host_in = bytes(d_in.cpu().numpy())
        got = unpack(bytes(d_out.cpu().numpy()), len(prog.outputs))
        vals = [int.from_bytes(host_in[i:i + 32], "little") for i in range(0, len(host_in), 32)]
        for j in list(range(0, m, 257)) + [m - 1]:
            assert got[j] == oracle_row(p, vals[j * tot:(j + 1) * tot])
        for j in range(m):
            z = vals[j * tot]
            assert got[j][1] == pow(z, 1 << 10, R) and got[j][2] * got[j][3] % R == 1
    finally:
        L.close()


@pytest.mark.gpu
def test_device_rejects_malformed_programs():
    import snark_verifier_b200 as sv
    L = sv.CudaLoader(0)
    try:
        good = pe.Program([(pe.OP_INPUT, 0, 0, 0), (pe.OP_MUL, 1, 0, 0)], 2, [], 1, [1], ["sq"])
        x = 12345
        assert unpack(L.fr_program_eval(good, pack_rows([[x]]), 1), 1) == [[x * x % R]]
        for bad in (pe.Program([(pe.OP_MUL, 1, 0, 0)], 2, [], 1, [1], ["x"]),                                   # read before write
                    pe.Program([(pe.OP_INPUT, 0, 3, 0)], 1, [], 1, [0], ["x"]),                                 # input slot out of range
                    pe.Program([(pe.OP_CONST, 0, 0, 0)], 1, [], 0, [0], ["x"]),                                 # no such constant
                    pe.Program([(99, 0, 0, 0)], 1, [], 1, [0], ["x"]),                                          # unknown opcode
                    pe.Program([(pe.OP_INPUT, 5, 0, 0)], 2, [], 1, [0], ["x"]),                                 # dst out of range
                    pe.Program([(pe.OP_INPUT, 0, 0, 0)], 2, [], 1, [1], ["x"])):                                # output never written
            with pytest.raises(sv.Error):
                L.fr_program_eval(bad, pack_rows([[x]]), 1)
        # the context stays usable after a rejected call
        assert unpack(L.fr_program_eval(good, pack_rows([[x]]), 1), 1) == [[x * x % R]]
    finally:
        L.close()
