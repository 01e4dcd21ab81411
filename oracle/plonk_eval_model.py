"""TEST INFRASTRUCTURE (oracle) — plain big-integer restatement of the reference's per-proof PLONK scalar evaluation; used only
by tests/ as the checker of the device path (snarkv_fr_program_eval_batch) and of the host compiler.  It shares no code with
snark_verifier_b200/: expressions arrive as nested tuples.

Follows (snark-verifier/src): verifier/plonk/protocol.rs:211-283 (CommonPolynomialEvaluation), :336-392 (Expression::evaluate),
verifier/plonk/proof.rs:298-349 (instance evaluations, quotient evaluation), util/arithmetic.rs:83-160 (root_of_unity, Domain),
loader.rs:255-262 + util/arithmetic.rs:47-74 (batch_invert leaves zeros untouched).  "Parity unpinned": the reference ships no
known-answer vectors for this path; the values are canonical mathematics in Fr.
"""
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
S = 28
GENERATOR = 7   # Fr::MULTIPLICATIVE_GENERATOR; ROOT_OF_UNITY = GENERATOR^((r - 1) / 2^S)


def root_of_unity(k):
    return pow(pow(GENERATOR, (R - 1) >> S, R), 1 << (S - k), R)


def inv_or_zero(x):
    return pow(x, -1, R) if x % R else 0


def rotate(k, i):
    """omega^i for the domain of size 2^k (Domain::rotate_scalar(1, Rotation(i)))"""
    w = root_of_unity(k)
    return pow(w, i, R) if i >= 0 else pow(pow(w, -1, R), -i, R)


def common_poly_eval(k, lagranges, z):
    """-> (zn, zn_minus_one, zn_minus_one_inv, {i: L_i(z)}) with L_i(z) = (z^n - 1) / n * omega^i / (z - omega^i)"""
    n = 1 << k
    zn = pow(z, n, R)
    zn_m1 = (zn - 1) % R
    numer = zn_m1 * pow(n, -1, R) % R
    lag = {}
    for i in sorted(set(lagranges)):
        om = rotate(k, i)
        lag[i] = numer * om % R * inv_or_zero((z - om) % R) % R
    return zn, zn_m1, inv_or_zero(zn_m1), lag


def eval_expr(e, z, lag, evals, challenges):
    """e: nested tuples (Expression.to_tuple()); evals: {(poly, rotation): value}"""
    t = e[0]
    if t == "constant":
        return e[1] % R
    if t == "common":
        return z % R if e[1] == "identity" else lag[e[2]]
    if t == "poly":
        return evals[(e[1], e[2])]
    if t == "challenge":
        return challenges[e[1]]
    if t == "neg":
        return (-eval_expr(e[1], z, lag, evals, challenges)) % R
    if t == "sum":
        return (eval_expr(e[1], z, lag, evals, challenges) + eval_expr(e[2], z, lag, evals, challenges)) % R
    if t == "product":
        return eval_expr(e[1], z, lag, evals, challenges) * eval_expr(e[2], z, lag, evals, challenges) % R
    if t == "scaled":
        return eval_expr(e[1], z, lag, evals, challenges) * e[2] % R
    assert t == "powers"
    exprs, scalar = e[1], e[2]
    acc = eval_expr(exprs[0], z, lag, evals, challenges)
    if len(exprs) == 1:
        return acc
    s = eval_expr(scalar, z, lag, evals, challenges)
    for x in exprs[1:]:
        acc = (acc * s + eval_expr(x, z, lag, evals, challenges)) % R
    return acc


def queries_of(e, out=None):
    out = set() if out is None else out
    if e[0] == "poly":
        out.add((e[1], e[2]))
    elif e[0] in ("neg", "scaled"):
        queries_of(e[1], out)
    elif e[0] in ("sum", "product"):
        queries_of(e[1], out); queries_of(e[2], out)
    elif e[0] == "powers":
        for x in e[1]:
            queries_of(x, out)
        queries_of(e[2], out)
    return out


def lagranges_of(e, out=None):
    out = set() if out is None else out
    if e[0] == "common" and e[1] == "lagrange":
        out.add(e[2])
    elif e[0] in ("neg", "scaled"):
        lagranges_of(e[1], out)
    elif e[0] in ("sum", "product"):
        lagranges_of(e[1], out); lagranges_of(e[2], out)
    elif e[0] == "powers":
        for x in e[1]:
            lagranges_of(x, out)
        lagranges_of(e[2], out)
    return out


def quotient_evaluation(k, num_preprocessed, num_instance, evaluation_queries, numerator, z, challenges, evaluations, instances):
    """One proof.  evaluation_queries: [(poly, rotation)] in protocol.evaluations order, `evaluations` the proof's values in that
    order, instances: list of columns.  -> [quotient_eval, zn, zn - 1, 1/(zn - 1), instance evaluations (sorted by query)...]"""
    inst_q = sorted(q for q in queries_of(numerator) if num_preprocessed <= q[0] < num_preprocessed + len(num_instance))
    lo = hi = 0
    for _, rot in inst_q:      # protocol.rs:88-96
        if rot < lo:
            lo = rot
        elif rot > hi:
            hi = rot
    lagr = lagranges_of(numerator) | set(range(-hi, (max(num_instance) if num_instance else 0) + abs(lo)))
    zn, zn_m1, zn_m1_inv, lag = common_poly_eval(k, lagr, z)
    evals = {}
    inst_vals = []
    for poly, rot in inst_q:   # proof.rs:313-334
        col = instances[poly - num_preprocessed]
        v = sum(x * lag[j - rot] for j, x in enumerate(col)) % R
        evals[(poly, rot)] = v
        inst_vals.append(v)
    for q, v in zip(evaluation_queries, evaluations):
        evals[tuple(q)] = v % R
    num = eval_expr(numerator, z, lag, evals, challenges)
    return [num * zn_m1_inv % R, zn, zn_m1, zn_m1_inv] + inst_vals


def run_program(instrs, n_regs, consts, inputs, out_regs):
    """Reference interpreter of the register program (include/snarkv_cuda.h SNARKV_FR_OP_*) for ONE proof: checks the host
    compiler without a GPU."""
    reg = [None] * n_regs
    for op, dst, a, b in instrs:
        if op == 0:
            v = inputs[a] % R
        elif op == 1:
            v = consts[a] % R
        elif op == 2:
            v = (reg[a] + reg[b]) % R
        elif op == 3:
            v = (reg[a] - reg[b]) % R
        elif op == 4:
            v = reg[a] * reg[b] % R
        elif op == 5:
            v = (-reg[a]) % R
        elif op == 6:
            v = inv_or_zero(reg[a])
        elif op == 7:
            v = reg[a] if reg[a] else 1
        elif op == 8:
            v = reg[a] if reg[b] else 0
        else:
            raise ValueError(op)
        reg[dst] = v
    return [reg[r] for r in out_regs]
