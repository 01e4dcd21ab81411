"""Host mirror of the callers that SHAPE the hot path's inputs: the KZG multi-open verifiers that turn a
proof's queries into the (lhs, rhs) `Msm` pair of a `KzgAccumulator`, and the limb encoding that feeds old accumulators back in.

Mirrored reference items (paths relative to snark-verifier/src):
  pcs.rs:20-49                         Query { poly, shift, eval }
  pcs/kzg/multiopen/gwc19.rs:45-82     Gwc19::verify          gwc19.rs:114-160  QuerySet::msm, query_sets
  pcs/kzg/multiopen/bdfg21.rs:51-83    Bdfg21::verify         bdfg21.rs:123-371 query_sets, query_set_coeffs, QuerySetCoeff
  pcs/kzg/accumulator.rs:57-81         LimbsEncoding::from_repr (native)        util/arithmetic.rs:270-298 fe_from_limbs / fe_to_limbs
  loader.rs:71-78, 255-262             powers, batch_invert (zero stays zero)

Scalars are Python ints mod r (µs of Fr work per proof, like the reference's NativeLoader); the two `Msm::evaluate` calls at the end
of each `verify` are the loader's `multi_scalar_multiplication` — on a CudaLoader, the GPU MSM.  For m proofs at once the same
`Msm`s go through `CudaLoader.msm_batch_rlc`, and `LimbsEncoding.from_repr_batch` runs on the device.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence

from . import CHECK_INPUTS, KzgAccumulator, Msm, R_MODULUS

Q_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47   # BN254 base field


def _inv(x: int) -> int:
    """LoadedScalar::invert().unwrap_or(value): batch_invert leaves a zero untouched (loader.rs:255-262)"""
    return pow(x, -1, R_MODULUS) if x % R_MODULUS else 0


def powers(x: int, n: int) -> List[int]:
    """loader.rs:71-78"""
    out, acc = [], 1
    for _ in range(n):
        out.append(acc)
        acc = acc * x % R_MODULUS
    return out


@dataclass
class Query:
    """pcs.rs:20-49: polynomial index, evaluation point = shift * z, claimed evaluation"""
    poly: int
    shift: int
    eval: int


# ----------------------------------------------------------------------------------------------------------------------
# GWC19 (gwc19.rs)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class Gwc19Proof:
    v: int
    ws: List[bytes]
    u: int


class Gwc19:
    @staticmethod
    def query_sets(queries: Sequence[Query]):
        """gwc19.rs:140-160: one set per distinct shift, in order of first appearance"""
        sets = []
        for q in queries:
            for s in sets:
                if s["shift"] == q.shift:
                    s["polys"].append(q.poly); s["evals"].append(q.eval)
                    break
            else:
                sets.append({"shift": q.shift, "polys": [q.poly], "evals": [q.eval]})
        return sets

    @staticmethod
    def verify(loader, svk_g: bytes, commitments: Sequence[Msm], z: int, queries: Sequence[Query], proof: Gwc19Proof) -> KzgAccumulator:
        """gwc19.rs:45-82"""
        sets = Gwc19.query_sets(queries)
        powers_of_u = powers(proof.u, len(sets))
        powers_of_v = powers(proof.v, max(len(s["polys"]) for s in sets))
        f = Msm(loader)
        for s, pu in zip(sets, powers_of_u):
            set_msm = Msm(loader)                                           # QuerySet::msm (gwc19.rs:120-137)
            for poly, ev, pv in zip(s["polys"], s["evals"], powers_of_v):
                set_msm = set_msm + (commitments[poly] - Msm.constant_(loader, ev)) * pv
            f = f + set_msm * pu
        z_omegas = [s["shift"] * z % R_MODULUS for s in sets]
        rhs = [Msm.base(loader, w) * pu for w, pu in zip(proof.ws, powers_of_u)]
        lhs = f + Msm.sum(loader, [uw * zo for uw, zo in zip(rhs, z_omegas)])
        return KzgAccumulator(lhs.evaluate(svk_g), Msm.sum(loader, rhs).evaluate(svk_g))


# ----------------------------------------------------------------------------------------------------------------------
# BDFG21 / SHPLONK (bdfg21.rs)
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class Bdfg21Proof:
    mu: int
    gamma: int
    w: bytes
    z_prime: int
    w_prime: bytes


class _Fraction:
    """util/arithmetic.rs:162-240 (numer / denom, inverted in a batch, then evaluated)"""

    def __init__(self, numer: Optional[int], denom: int):
        self.numer, self.denom, self.eval, self.inv = numer, denom, None, False

    def evaluate(self):
        assert self.inv
        if self.eval is None:
            self.eval = self.denom if self.numer is None else self.numer * self.denom % R_MODULUS


class _QuerySetCoeff:
    """bdfg21.rs:260-371"""

    def __init__(self, shifts, powers_of_z, z_prime, z_prime_minus_z_shift_i, z_s_1):
        R = R_MODULUS
        ell = []
        for j, sj in enumerate(shifts):
            acc = 1
            for i, si in enumerate(shifts):
                if i != j:
                    acc = acc * (sj - si) % R
            ell.append(acc)
        z = powers_of_z[1]
        z_pow = powers_of_z[len(shifts) - 1]
        # sum_products_with_coeff([(ell, z^(k-1), z'), (-(ell * shift), z^(k-1), z)])
        self.eval_coeffs = [_Fraction(None, (e * z_pow % R * z_prime - e * s % R * z_pow % R * z) % R) for s, e in zip(shifts, ell)]
        z_s = 1
        for s in shifts:
            z_s = z_s * z_prime_minus_z_shift_i[s] % R
        self.z_s = z_s
        self.commitment_coeff = None if z_s_1 is None else _Fraction(z_s_1, z_s)
        self.r_eval_coeff = None

    def denoms(self):
        """bdfg21.rs:331-363: first call -> barycentric weights (+ z_s_1 / z_s), second call -> the r_eval coefficient"""
        fr = self.eval_coeffs + ([self.commitment_coeff] if self.commitment_coeff is not None else [])
        if not self.eval_coeffs[0].inv:
            for f in fr:
                f.inv = True
            return fr
        assert self.r_eval_coeff is None
        for f in fr:
            f.evaluate()
        wsum = sum(f.eval for f in self.eval_coeffs) % R_MODULUS
        self.r_eval_coeff = _Fraction(self.commitment_coeff.eval if self.commitment_coeff is not None else None, wsum)
        self.r_eval_coeff.inv = True
        return [self.r_eval_coeff]


class Bdfg21:
    @staticmethod
    def query_sets(queries: Sequence[Query]):
        """bdfg21.rs:123-175: polynomials grouped by their SET of shifts; evals re-ordered to the set's shift order"""
        poly_shifts = []
        for q in queries:
            for ps in poly_shifts:
                if ps[0] == q.poly:
                    if q.shift not in ps[1]:
                        ps[1].append(q.shift); ps[2].append(q.eval)
                    break
            else:
                poly_shifts.append((q.poly, [q.shift], [q.eval]))
        sets = []
        for poly, shifts, evals in poly_shifts:
            for s in sets:
                if set(s["shifts"]) == set(shifts):
                    if poly not in s["polys"]:
                        s["polys"].append(poly)
                        s["evals"].append([evals[shifts.index(lhs)] for lhs in s["shifts"]])
                    break
            else:
                sets.append({"shifts": shifts, "polys": [poly], "evals": [evals]})
        return sets

    @staticmethod
    def query_set_coeffs(sets, z: int, z_prime: int):
        """bdfg21.rs:177-222"""
        superset = sorted({s for st in sets for s in st["shifts"]})
        size = max([len(st["shifts"]) for st in sets] + [2])
        powers_of_z = powers(z, size)
        zp_minus = {s: (z_prime - z * s) % R_MODULUS for s in superset}
        z_s_1, coeffs = None, []
        for st in sets:
            c = _QuerySetCoeff(st["shifts"], powers_of_z, z_prime, zp_minus, z_s_1)
            if z_s_1 is None:
                z_s_1 = c.z_s
            coeffs.append(c)
        for _ in range(2):                                                  # two rounds of L::batch_invert over the denominators
            for c in coeffs:
                for f in c.denoms():
                    f.denom = _inv(f.denom)
        for c in coeffs:
            c.r_eval_coeff.evaluate()
        return coeffs

    @staticmethod
    def verify(loader, svk_g: bytes, commitments: Sequence[Msm], z: int, queries: Sequence[Query], proof: Bdfg21Proof) -> KzgAccumulator:
        """bdfg21.rs:51-83"""
        R = R_MODULUS
        sets = Bdfg21.query_sets(queries)
        coeffs = Bdfg21.query_set_coeffs(sets, z, proof.z_prime)
        powers_of_mu = powers(proof.mu, max(len(s["polys"]) for s in sets))
        f = Msm(loader)
        for st, co, pg in zip(sets, coeffs, powers(proof.gamma, len(sets))):
            set_msm = Msm(loader)                                           # QuerySet::msm (bdfg21.rs:231-258)
            for poly, evals, pm in zip(st["polys"], st["evals"], powers_of_mu):
                commitment = commitments[poly] * co.commitment_coeff.eval if co.commitment_coeff is not None else commitments[poly]
                r_eval = sum(c.eval * e for c, e in zip(co.eval_coeffs, evals)) % R * co.r_eval_coeff.eval % R
                set_msm = set_msm + (commitment - Msm.constant_(loader, r_eval)) * pm
            f = f + set_msm * pg
        f = f - Msm.base(loader, proof.w) * coeffs[0].z_s
        rhs = Msm.base(loader, proof.w_prime)
        lhs = f + rhs * proof.z_prime
        return KzgAccumulator(lhs.evaluate(svk_g), rhs.evaluate(svk_g))


# ----------------------------------------------------------------------------------------------------------------------
# LimbsEncoding (accumulator.rs:28-82)
# ----------------------------------------------------------------------------------------------------------------------
class InvalidAccumulator(ValueError):
    """The reference panics (`from_repr(..).unwrap()`, `from_xy(..).unwrap()`); the mirror raises."""


class LimbsEncoding:
    """`LimbsEncoding<LIMBS, BITS>`: a KZG accumulator as 4 * LIMBS scalar-field limbs (lhs.x, lhs.y, rhs.x, rhs.y)."""

    def __init__(self, limbs: int = 4, bits: int = 68):
        self.limbs, self.bits = limbs, bits

    def fe_to_limbs(self, fe: int) -> List[int]:
        """util/arithmetic.rs:284-298"""
        mask = (1 << self.bits) - 1
        return [(fe >> (self.bits * i)) & mask for i in range(self.limbs)]

    def fe_from_limbs(self, limbs: Sequence[int]) -> int:
        """util/arithmetic.rs:270-282: sum limb_i << (BITS i); must fit 32 bytes and be a canonical Fq (fe_from_big)"""
        v = sum(l << (self.bits * i) for i, l in enumerate(limbs))
        if v >= Q_MODULUS:
            raise InvalidAccumulator("limbs do not encode a base-field element")
        return v

    def to_repr(self, acc: KzgAccumulator) -> List[int]:
        coords = [int.from_bytes(b[i:i + 32], "little") for b in (acc.lhs, acc.rhs) for i in (0, 32)]
        return [l for c in coords for l in self.fe_to_limbs(c)]

    def from_repr(self, limbs: Sequence[int]) -> KzgAccumulator:
        """accumulator.rs:57-81"""
        assert len(limbs) == 4 * self.limbs
        c = [self.fe_from_limbs(limbs[i * self.limbs:(i + 1) * self.limbs]) for i in range(4)]
        for x, y in ((c[0], c[1]), (c[2], c[3])):                            # C::from_xy(..).unwrap(): on the curve, or the identity (0, 0)
            if (x, y) != (0, 0) and (y * y - x * x * x - 3) % Q_MODULUS:
                raise InvalidAccumulator("point is not on the curve")
        le = lambda v: v.to_bytes(32, "little")
        return KzgAccumulator(le(c[0]) + le(c[1]), le(c[2]) + le(c[3]))

    def from_repr_batch(self, loader, limbs: bytes, m: int):
        """m accumulators at once on the device (snarkv_kzg_accumulators_from_limbs): limbs = m x 4 LIMBS x 32 B scalars in the
        loader's format -> (lhs m x 64 B, rhs m x 64 B, valid m bytes); valid[i] = 0 where the reference would panic."""
        return loader.accumulators_from_limbs(limbs, m, self.limbs, self.bits)


# ----------------------------------------------------------------------------------------------------------------------
# m proofs of ONE protocol at once (BASELINE config 3): scalars by a device program, one fused MSM per side, one pairing
# ----------------------------------------------------------------------------------------------------------------------
class KzgBatchVerifier:
    """`Gwc19::verify` / `Bdfg21::verify` + the RLC `decide_all` of pcs/kzg/decider.rs:146-185 for a batch of proofs that share a
    protocol:
      1. the MSM scalars of every proof — the verifier's `Msm` algebra (gwc19.rs:52-81 / bdfg21.rs:58-82 with its coefficient
         computation) compiled once to a straight-line Fr program (plonk_eval.py) — on the device;
      2. lhs = sum_j rho^j lhs_j and rhs = sum_j rho^j rhs_j, each as ONE Pippenger pass (snarkv_g1_msm_batch_rlc);
      3. one pairing check e(lhs, g2) e(rhs, -s g2) == 1.
    Per-proof data (dict): "commitments", "evals" (query order), "ws" (opening-proof points: W_0.. for GWC19, [W, W'] for
    SHPLONK) and the scalars "z", "v", "u" (GWC19) or "z", "mu", "gamma", "z_prime" (SHPLONK)."""

    SCALARS = {"gwc19": ("z", "v", "u"), "bdfg21": ("z", "mu", "gamma", "z_prime")}

    def __init__(self, loader, kzg, svk_g: bytes, queries: Sequence[tuple], num_polys: int, scheme: str = "gwc19"):
        from .plonk_eval import compile_bdfg21_msm_scalars, compile_gwc19_msm_scalars
        self.loader, self.kzg, self.g, self.scheme = loader, kzg, bytes(svk_g), scheme
        self.flags = CHECK_INPUTS   # checked by default; a caller that has already validated every point may clear it
        self.compiled = {"gwc19": compile_gwc19_msm_scalars, "bdfg21": compile_bdfg21_msm_scalars}[scheme](queries, num_polys)

    def _points(self, slots, proof):
        out = []
        for s in slots:
            out.append(self.g if s == ("g",) else proof["commitments"][s[1]] if s[0] == "c" else proof["ws"][s[1]])
        return b"".join(out)

    def accumulate(self, proofs: Sequence[dict], rho: int) -> KzgAccumulator:
        cp, m = self.compiled, len(proofs)
        le = lambda v: (v % R_MODULUS).to_bytes(32, "little")
        rows = b"".join(b"".join(le(p[k]) for k in self.SCALARS[self.scheme]) + b"".join(le(e) for e in p["evals"]) for p in proofs)
        lhs_p = b"".join(self._points(cp.lhs_slots, p) for p in proofs)
        rhs_p = b"".join(self._points(cp.rhs_slots, p) for p in proofs)
        return self.accumulate_packed(rows, lhs_p, rhs_p, m, rho)

    def accumulate_packed(self, rows, lhs_points, rhs_points, m: int, rho: int) -> KzgAccumulator:
        """The same on packed arrays (bytes or numpy uint8): `rows` = m x program inputs x 32 B, `lhs_points` / `rhs_points` = the
        bases of every proof in `compiled.lhs_slots` / `rhs_slots` order, m x slots x 64 B."""
        import numpy as np
        cp = self.compiled
        nl, nr = len(cp.lhs_slots), len(cp.rhs_slots)
        scal = np.frombuffer(self.loader.fr_program_eval(cp.program, rows, m), dtype=np.uint8).reshape(m, nl + nr, 32)
        lhs_s = np.ascontiguousarray(scal[:, :nl]).reshape(-1)
        rhs_s = np.ascontiguousarray(scal[:, nl:]).reshape(-1)
        rho_b = (rho % R_MODULUS).to_bytes(32, "little")
        # proof-supplied points are validated on the device like the reference's `read_ec_point` / `from_xy` does before any
        # arithmetic (canonical coordinates, on the curve): an invalid one surfaces as Error (SNARKV_ERR_BAD_POINT)
        lhs = self.loader.msm_batch_rlc(lhs_s, lhs_points, np.arange(m + 1, dtype=np.uint64) * nl, rho_b, flags=self.flags)
        rhs = self.loader.msm_batch_rlc(rhs_s, rhs_points, np.arange(m + 1, dtype=np.uint64) * nr, rho_b, flags=self.flags)
        return KzgAccumulator(lhs, rhs)

    def verify_batch(self, proofs: Sequence[dict], rho: int):
        """Raises AssertionFailure (decider.rs:81) unless the random linear combination of all proofs decides."""
        self.kzg.decide(self.accumulate(proofs, rho))


class Gwc19BatchVerifier(KzgBatchVerifier):
    def __init__(self, loader, kzg, svk_g, queries, num_polys):
        super().__init__(loader, kzg, svk_g, queries, num_polys, "gwc19")


class Bdfg21BatchVerifier(KzgBatchVerifier):
    def __init__(self, loader, kzg, svk_g, queries, num_polys):
        super().__init__(loader, kzg, svk_g, queries, num_polys, "bdfg21")
