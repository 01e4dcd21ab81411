"""CPU: the Pallas model (oracle/pasta_model.py) against the definitions it restates — h_coeffs (pcs/ipa.rs:401-417) literal vs the
closed product form the device kernel uses, IpaAs::decide (pcs/ipa/decider.rs:47-59) accept / reject, group-order identities."""
import random

from oracle import pasta_model as pm


def test_curve_parameters_are_self_consistent():
    assert pm.P.bit_length() == 255 and pm.Q.bit_length() == 255
    assert pm.is_on_curve(pm.GEN) and pm.mul(pm.GEN, pm.Q) is None
    g2 = pm.add(pm.GEN, pm.GEN)
    assert g2 == pm.mul(pm.GEN, 2) and pm.is_on_curve(g2)
    assert pm.add(pm.mul(pm.GEN, 5), pm.neg(pm.mul(pm.GEN, 5))) is None
    assert pm.mul(pm.GEN, 7) == pm.msm_naive([3, 4], [pm.GEN, pm.GEN])


def test_h_coeffs_literal_equals_bit_product_form():
    rnd = random.Random(1)
    for k in (1, 2, 5, 8):
        xi = [rnd.randrange(pm.Q) for _ in range(k)]
        s = rnd.randrange(pm.Q)
        h = pm.h_coeffs(xi, s)
        for j in range(1 << k):
            want = s
            for i in range(k):
                if (j >> i) & 1:
                    want = want * xi[k - 1 - i] % pm.Q
            assert h[j] == want


def test_ipa_decide_accepts_and_rejects():
    rnd = random.Random(2)
    k = 4
    xi = [rnd.randrange(1, pm.Q) for _ in range(k)]
    g = [pm.mul(pm.GEN, rnd.randrange(1, pm.Q)) for _ in range(1 << k)]
    u = pm.msm_naive(pm.h_coeffs(xi), g)
    assert pm.ipa_decide(g, u, xi)
    assert not pm.ipa_decide(g, pm.add(u, pm.GEN), xi)
    assert not pm.ipa_decide(g, u, xi[:-1] + [(xi[-1] + 1) % pm.Q])
