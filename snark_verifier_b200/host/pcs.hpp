// pcs.hpp — C++ host mirror of the SHPLONK multi-open verifier `Bdfg21::verify` (snark-verifier/src/pcs/kzg/multiopen/bdfg21.rs):
//   bdfg21.rs:51-83    verify: (lhs, rhs) Msm pair of the accumulator
//   bdfg21.rs:123-175  query_sets: polynomials grouped by their SET of shifts, evaluations re-ordered to the set's shift order
//   bdfg21.rs:177-222  query_set_coeffs: two rounds of L::batch_invert over the Fraction denominators
//   bdfg21.rs:224-371  QuerySet::msm, QuerySetCoeff::{new, denoms, evaluate}
// GWC19 lives in cuda_loader.hpp (`Gwc19`); snark_verifier_b200/pcs.py is the Python twin of both.  Scalars are canonical 32-byte
// values handled through the caller's FrOps (add, mul, neg, inv with inv(0) = 0 — `batch_invert` leaves zeros untouched,
// loader.rs:255-262); the two `evaluate` calls at the end are the loader's multi_scalar_multiplication, i.e. the GPU MSM.
// tests/pcs_mirror_test.cpp runs this header on the CPU over a recording loader and the Python mirror checks every (scalar, base) pair.
#pragma once
#include <algorithm>
#include <map>
#include <optional>

#include "cuda_loader.hpp"

namespace snarkv {

struct Bdfg21Proof { Fr mu, gamma; G1Affine w; Fr z_prime; G1Affine w_prime; };

template <class Loader>
struct BasicBdfg21 {
    using Msm = BasicMsm<Loader>;
    struct Fraction {   // util/arithmetic.rs:162-240
        std::optional<Fr> numer;
        Fr denom;
        std::optional<Fr> eval;
        bool inv = false;
        void evaluate(const FrOps& o) {
            if (!inv) throw Error("Fraction::evaluate before inversion");
            if (!eval) eval = numer ? o.mul(*numer, denom) : denom;
        }
    };
    struct QuerySet { std::vector<Fr> shifts; std::vector<size_t> polys; std::vector<std::vector<Fr>> evals; };
    struct Coeff {   // QuerySetCoeff, bdfg21.rs:260-371
        Fr z_s;
        std::vector<Fraction> eval_coeffs;
        std::optional<Fraction> commitment_coeff, r_eval_coeff;
    };

    static Fr sub(const FrOps& o, const Fr& a, const Fr& b) { return o.add(a, o.neg(b)); }
    static bool fr_less(const Fr& a, const Fr& b) {   // canonical little-endian bytes compared as integers (BTreeSet<F> order)
        for (int i = 31; i >= 0; --i) if (a[i] != b[i]) return a[i] < b[i];
        return false;
    }

    static std::vector<QuerySet> query_sets(const std::vector<Query>& queries) {   // bdfg21.rs:123-175
        struct PolyShifts { size_t poly; std::vector<Fr> shifts, evals; };
        std::vector<PolyShifts> ps;
        for (const Query& q : queries) {
            auto it = std::find_if(ps.begin(), ps.end(), [&](const PolyShifts& p) { return p.poly == q.poly; });
            if (it == ps.end()) ps.push_back({q.poly, {q.shift}, {q.eval}});
            else if (std::find(it->shifts.begin(), it->shifts.end(), q.shift) == it->shifts.end()) { it->shifts.push_back(q.shift); it->evals.push_back(q.eval); }
        }
        std::vector<QuerySet> sets;
        for (const PolyShifts& p : ps) {
            auto same_set = [&](const QuerySet& s) {
                if (s.shifts.size() != p.shifts.size()) return false;
                for (const Fr& x : p.shifts) if (std::find(s.shifts.begin(), s.shifts.end(), x) == s.shifts.end()) return false;
                return true;
            };
            auto it = std::find_if(sets.begin(), sets.end(), same_set);
            if (it == sets.end()) { sets.push_back({p.shifts, {p.poly}, {p.evals}}); continue; }
            if (std::find(it->polys.begin(), it->polys.end(), p.poly) != it->polys.end()) continue;
            it->polys.push_back(p.poly);
            std::vector<Fr> ordered;
            for (const Fr& lhs : it->shifts) ordered.push_back(p.evals[std::find(p.shifts.begin(), p.shifts.end(), lhs) - p.shifts.begin()]);
            it->evals.push_back(ordered);
        }
        return sets;
    }

    static std::vector<Coeff> query_set_coeffs(const FrOps& o, const std::vector<QuerySet>& sets, const Fr& z, const Fr& z_prime) {   // bdfg21.rs:177-222
        if (!o.neg || !o.inv) throw Error("Bdfg21: FrOps::neg and FrOps::inv are required");
        std::vector<Fr> superset;
        for (const QuerySet& s : sets) for (const Fr& x : s.shifts) if (std::find(superset.begin(), superset.end(), x) == superset.end()) superset.push_back(x);
        std::sort(superset.begin(), superset.end(), fr_less);
        size_t size = 2;
        for (const QuerySet& s : sets) size = std::max(size, s.shifts.size());
        const std::vector<Fr> powers_of_z = BasicGwc19<Loader>::powers(o, z, size);
        std::vector<std::pair<Fr, Fr>> zp_minus;   // shift -> z' - z shift
        for (const Fr& x : superset) zp_minus.push_back({x, sub(o, z_prime, o.mul(z, x))});
        auto zp_of = [&](const Fr& x) { for (auto& kv : zp_minus) if (kv.first == x) return kv.second; throw Error("shift not in superset"); };
        std::optional<Fr> z_s_1;
        std::vector<Coeff> coeffs;
        for (const QuerySet& s : sets) {   // QuerySetCoeff::new, bdfg21.rs:268-326
            Coeff c;
            const Fr& zz = powers_of_z[1];
            const Fr& z_pow = powers_of_z[s.shifts.size() - 1];
            for (size_t j = 0; j < s.shifts.size(); ++j) {
                Fr ell = o.one;   // normalized_ell_prime_j = prod_{i != j} (shift_j - shift_i)
                for (size_t i = 0; i < s.shifts.size(); ++i) if (i != j) ell = o.mul(ell, sub(o, s.shifts[j], s.shifts[i]));
                // sum_products_with_coeff([(ell, z^(k-1), z'), (-(ell shift_j), z^(k-1), z)])
                const Fr t1 = o.mul(o.mul(ell, z_pow), z_prime);
                const Fr t2 = o.mul(o.mul(o.neg(o.mul(ell, s.shifts[j])), z_pow), zz);
                c.eval_coeffs.push_back(Fraction{std::nullopt, o.add(t1, t2), std::nullopt, false});   // Fraction::one_over
            }
            Fr z_s = o.one;
            for (const Fr& x : s.shifts) z_s = o.mul(z_s, zp_of(x));
            c.z_s = z_s;
            if (z_s_1) c.commitment_coeff = Fraction{*z_s_1, z_s, std::nullopt, false};   // Fraction::new(z_s_1, z_s)
            else z_s_1 = z_s;
            coeffs.push_back(c);
        }
        // first batch_invert: barycentric weights (+ z_s_1 / z_s); second: the r_eval coefficient (bdfg21.rs:218-220, 331-363)
        for (Coeff& c : coeffs) {
            for (Fraction& f : c.eval_coeffs) { f.denom = o.inv(f.denom); f.inv = true; }
            if (c.commitment_coeff) { c.commitment_coeff->denom = o.inv(c.commitment_coeff->denom); c.commitment_coeff->inv = true; }
        }
        for (Coeff& c : coeffs) {
            for (Fraction& f : c.eval_coeffs) f.evaluate(o);
            if (c.commitment_coeff) c.commitment_coeff->evaluate(o);
            Fr wsum = *c.eval_coeffs[0].eval;
            for (size_t i = 1; i < c.eval_coeffs.size(); ++i) wsum = o.add(wsum, *c.eval_coeffs[i].eval);
            c.r_eval_coeff = c.commitment_coeff ? Fraction{*c.commitment_coeff->eval, wsum, std::nullopt, false} : Fraction{std::nullopt, wsum, std::nullopt, false};
            c.r_eval_coeff->denom = o.inv(c.r_eval_coeff->denom);
            c.r_eval_coeff->inv = true;
        }
        for (Coeff& c : coeffs) c.r_eval_coeff->evaluate(o);
        return coeffs;
    }

    static KzgAccumulator verify(Loader& loader, const FrOps& o, const G1Affine& svk_g, const std::vector<Msm>& commitments, const Fr& z,
                                 const std::vector<Query>& queries, const Bdfg21Proof& proof) {   // bdfg21.rs:51-83
        const std::vector<QuerySet> sets = query_sets(queries);
        const std::vector<Coeff> coeffs = query_set_coeffs(o, sets, z, proof.z_prime);
        size_t max_polys = 0;
        for (const QuerySet& s : sets) max_polys = std::max(max_polys, s.polys.size());
        const std::vector<Fr> powers_of_mu = BasicGwc19<Loader>::powers(o, proof.mu, max_polys);
        const std::vector<Fr> powers_of_gamma = BasicGwc19<Loader>::powers(o, proof.gamma, sets.size());
        Msm f(loader, o);
        for (size_t k = 0; k < sets.size(); ++k) {
            const QuerySet& st = sets[k];
            const Coeff& co = coeffs[k];
            Msm set_msm(loader, o);   // QuerySet::msm, bdfg21.rs:231-258
            for (size_t i = 0; i < st.polys.size(); ++i) {
                const Msm commitment = co.commitment_coeff ? commitments[st.polys[i]] * *co.commitment_coeff->eval : commitments[st.polys[i]];
                Fr r_eval = o.mul(*co.eval_coeffs[0].eval, st.evals[i][0]);   // loader.sum_products
                for (size_t j = 1; j < co.eval_coeffs.size(); ++j) r_eval = o.add(r_eval, o.mul(*co.eval_coeffs[j].eval, st.evals[i][j]));
                r_eval = o.mul(r_eval, *co.r_eval_coeff->eval);
                set_msm = set_msm + (commitment - Msm::constant(loader, o, r_eval)) * powers_of_mu[i];
            }
            f = f + set_msm * powers_of_gamma[k];
        }
        f = f - Msm::base(loader, o, proof.w) * coeffs[0].z_s;
        const Msm rhs = Msm::base(loader, o, proof.w_prime);
        const Msm lhs = f + rhs * proof.z_prime;
        return KzgAccumulator{lhs.evaluate(svk_g), rhs.evaluate(svk_g)};
    }
};
using Bdfg21 = BasicBdfg21<CudaLoader>;

}  // namespace snarkv
