// plonk_eval.hpp — C++ host mirror of snark-verifier's per-proof PLONK scalar evaluation and its compiler to the straight-line
// Fr register program that snarkv_fr_program_eval_batch (csrc/fr_program.cu) runs for a batch of proofs.
//
// Mirrored reference items (paths relative to snark-verifier/src):
//   util/arithmetic.rs:83-160            root_of_unity, Rotation, Domain::{new, rotate_scalar}
//   verifier/plonk/protocol.rs:77-106    PlonkProtocol::langranges
//   verifier/plonk/protocol.rs:186-283   CommonPolynomial, CommonPolynomialEvaluation::{new, denoms, evaluate}
//   verifier/plonk/protocol.rs:304-455   Query, Expression, Expression::evaluate, used_langrange, used_query
//   verifier/plonk/proof.rs:298-349      instance evaluations, quotient evaluation (linearization: None)
//   loader.rs:52-69                      LoadedScalar::pow_const      util/arithmetic.rs:47-74  batch_invert (zero stays zero)
// The protocol of a batch is fixed, so `Expression::evaluate` runs ONCE with closures that emit instructions (exactly how the
// reference drives its EVM / Halo2 loaders); the result is a snarkv::FrProgram.  snark_verifier_b200/plonk_eval.py is the same
// compiler in Python; tests/test_plonk_eval.py::test_cpp_compiler_emits_the_same_program checks that both emit identical programs.
#pragma once
#include <algorithm>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <vector>

#include "cuda_loader.hpp"

namespace snarkv {
namespace plonk {

// ---- Fr on the host: canonical values as 4 x u64, little endian (setup-time arithmetic only: a few hundred operations) --------
struct Fe {
    uint64_t v[4] = {0, 0, 0, 0};
    bool operator==(const Fe& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    bool operator<(const Fe& o) const {
        for (int i = 3; i >= 0; --i) if (v[i] != o.v[i]) return v[i] < o.v[i];
        return false;
    }
    bool is_zero() const { return !(v[0] | v[1] | v[2] | v[3]); }
};
inline const Fe& fe_modulus() {
    static const Fe r{{0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}};
    return r;
}
inline Fe fe_from_u64(uint64_t x) { Fe r; r.v[0] = x; return r; }
inline Fe fe_add(const Fe& a, const Fe& b) {
    Fe r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (unsigned __int128)a.v[i] + b.v[i]; r.v[i] = (uint64_t)c; c >>= 64; }
    const Fe& m = fe_modulus();
    if (!(r < m)) {
        unsigned __int128 br = 0;
        for (int i = 0; i < 4; ++i) { unsigned __int128 d = (unsigned __int128)r.v[i] - m.v[i] - br; r.v[i] = (uint64_t)d; br = (d >> 64) & 1; }
    }
    return r;
}
inline Fe fe_neg(const Fe& a) {
    if (a.is_zero()) return a;
    Fe r;
    const Fe& m = fe_modulus();
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; ++i) { unsigned __int128 d = (unsigned __int128)m.v[i] - a.v[i] - br; r.v[i] = (uint64_t)d; br = (d >> 64) & 1; }
    return r;
}
inline Fe fe_sub(const Fe& a, const Fe& b) { return fe_add(a, fe_neg(b)); }
inline Fe fe_mul(const Fe& a, const Fe& b) {   // double-and-add: slow and obviously right; the compiler needs a handful of these
    Fe acc;
    for (int bit = 255; bit >= 0; --bit) {
        acc = fe_add(acc, acc);
        if ((b.v[bit >> 6] >> (bit & 63)) & 1) acc = fe_add(acc, a);
    }
    return acc;
}
inline Fe fe_pow(const Fe& a, const Fe& e) {
    Fe acc = fe_from_u64(1);
    for (int bit = 255; bit >= 0; --bit) {
        acc = fe_mul(acc, acc);
        if ((e.v[bit >> 6] >> (bit & 63)) & 1) acc = fe_mul(acc, a);
    }
    return acc;
}
inline Fe fe_pow_u64(const Fe& a, uint64_t e) { return fe_pow(a, fe_from_u64(e)); }
inline Fe fe_inv(const Fe& a) {   // a^(r-2)
    Fe e = fe_modulus();
    e.v[0] -= 2;   // r ends in ...0001: no borrow
    return fe_pow(a, e);
}
inline Fr fe_to_bytes(const Fe& a) { Fr r; memcpy(r.data(), a.v, 32); return r; }

constexpr int FR_S = 28;
inline Fe fr_root_of_unity() {   // Fr::ROOT_OF_UNITY = 7^((r-1)/2^28)
    static const Fe w{{0xd34f1ed960c37c9cull, 0x3215cf6dd39329c8ull, 0x98865ea93dd31f74ull, 0x03ddb9f5166d18b7ull}};
    return w;
}
inline Fe root_of_unity(int k) {   // util/arithmetic.rs:83-90
    if (k > FR_S) throw Error("root_of_unity: k > S");
    Fe w = fr_root_of_unity();
    for (int i = 0; i < FR_S - k; ++i) w = fe_mul(w, w);
    return w;
}
inline Fe fr_delta() {   // Fr::DELTA = 7^(2^28)
    Fe d = fe_from_u64(7);
    for (int i = 0; i < FR_S; ++i) d = fe_mul(d, d);
    return d;
}

struct Domain {   // util/arithmetic.rs:123-160
    int k;
    uint64_t n;
    Fe n_inv, gen, gen_inv;
    explicit Domain(int k_) : k(k_), n(1ull << k_) {
        gen = root_of_unity(k);
        n_inv = fe_inv(fe_from_u64(n));
        gen_inv = fe_inv(gen);
    }
    Fe rotate_scalar(const Fe& scalar, int32_t rotation) const {
        if (rotation == 0) return scalar;
        if (rotation > 0) return fe_mul(scalar, fe_pow_u64(gen, (uint64_t)rotation));
        return fe_mul(scalar, fe_pow_u64(gen_inv, (uint64_t)(-(int64_t)rotation)));
    }
};

struct Query {   // protocol.rs:304-320
    size_t poly;
    int32_t rotation;
    bool operator<(const Query& o) const { return poly != o.poly ? poly < o.poly : rotation < o.rotation; }
    bool operator==(const Query& o) const { return poly == o.poly && rotation == o.rotation; }
};
struct CommonPolynomial {   // protocol.rs:186-197
    bool identity;
    int32_t index;
};

// protocol.rs:322-334
struct Expression;
using Expr = std::shared_ptr<const Expression>;
struct Expression {
    enum Tag { Constant, Common, Polynomial, Challenge, Negated, Sum, Product, Scaled, DistributePowers } tag;
    Fe scalar;                   // Constant / Scaled
    CommonPolynomial common{};   // Common
    Query query{};               // Polynomial
    size_t challenge = 0;        // Challenge
    std::vector<Expr> kids;      // Negated (1), Sum / Product (2), Scaled (1), DistributePowers (exprs..., scalar last)
};
inline Expr constant(const Fe& c) { auto e = std::make_shared<Expression>(); e->tag = Expression::Constant; e->scalar = c; return e; }
inline Expr common(bool identity, int32_t i = 0) { auto e = std::make_shared<Expression>(); e->tag = Expression::Common; e->common = {identity, i}; return e; }
inline Expr lagrange(int32_t i) { return common(false, i); }
inline Expr polynomial(size_t poly, int32_t rot = 0) { auto e = std::make_shared<Expression>(); e->tag = Expression::Polynomial; e->query = {poly, rot}; return e; }
inline Expr challenge(size_t i) { auto e = std::make_shared<Expression>(); e->tag = Expression::Challenge; e->challenge = i; return e; }
inline Expr negated(Expr a) { auto e = std::make_shared<Expression>(); e->tag = Expression::Negated; e->kids = {a}; return e; }
inline Expr operator+(Expr a, Expr b) { auto e = std::make_shared<Expression>(); e->tag = Expression::Sum; e->kids = {a, b}; return e; }
inline Expr operator*(Expr a, Expr b) { auto e = std::make_shared<Expression>(); e->tag = Expression::Product; e->kids = {a, b}; return e; }
inline Expr operator-(Expr a, Expr b) { return a + negated(b); }   // impl_expression_ops!(Sub, sub, Sum, .., Neg::neg)
inline Expr scaled(Expr a, const Fe& c) { auto e = std::make_shared<Expression>(); e->tag = Expression::Scaled; e->scalar = c; e->kids = {a}; return e; }
inline Expr distribute_powers(std::vector<Expr> exprs, Expr scalar) {
    auto e = std::make_shared<Expression>(); e->tag = Expression::DistributePowers; e->kids = std::move(exprs); e->kids.push_back(scalar); return e;
}

// protocol.rs:336-392: the eight-closure fold
template <class T>
struct Fold {
    std::function<T(const Fe&)> constant;
    std::function<T(const CommonPolynomial&)> common_poly;
    std::function<T(const Query&)> poly;
    std::function<T(size_t)> challenge;
    std::function<T(const T&)> negated;
    std::function<T(const T&, const T&)> sum, product;
    std::function<T(const T&, const Fe&)> scaled;
};
template <class T>
T evaluate(const Expression& e, const Fold<T>& f) {
    switch (e.tag) {
        case Expression::Constant: return f.constant(e.scalar);
        case Expression::Common: return f.common_poly(e.common);
        case Expression::Polynomial: return f.poly(e.query);
        case Expression::Challenge: return f.challenge(e.challenge);
        case Expression::Negated: return f.negated(evaluate(*e.kids[0], f));
        case Expression::Sum: { T a = evaluate(*e.kids[0], f); T b = evaluate(*e.kids[1], f); return f.sum(a, b); }
        case Expression::Product: { T a = evaluate(*e.kids[0], f); T b = evaluate(*e.kids[1], f); return f.product(a, b); }
        case Expression::Scaled: return f.scaled(evaluate(*e.kids[0], f), e.scalar);
        default: {   // DistributePowers: Horner in the scalar, first expression = highest power
            const size_t n = e.kids.size() - 1;
            if (n == 0) throw Error("DistributePowers: empty");
            if (n == 1) return evaluate(*e.kids[0], f);
            T acc = evaluate(*e.kids[0], f);
            T s = evaluate(*e.kids[n], f);
            for (size_t i = 1; i < n; ++i) {
                T scaled_acc = f.product(acc, s);          // sequenced explicitly: Rust evaluates `sum(product(..), evaluate(..))`
                T next = evaluate(*e.kids[i], f);          // left to right, C++ leaves the order of arguments unspecified
                acc = f.sum(scaled_acc, next);
            }
            return acc;
        }
    }
}
inline std::set<int32_t> used_langrange(const Expression& e) {   // protocol.rs:417-435
    using S = std::set<int32_t>;
    auto merge = [](const S& a, const S& b) { S r = a; r.insert(b.begin(), b.end()); return r; };
    Fold<S> f{[](const Fe&) { return S{}; }, [](const CommonPolynomial& p) { return p.identity ? S{} : S{p.index}; }, [](const Query&) { return S{}; },
              [](size_t) { return S{}; }, [](const S& a) { return a; }, merge, merge, [](const S& a, const Fe&) { return a; }};
    return evaluate<S>(e, f);
}
inline std::set<Query> used_query(const Expression& e) {   // protocol.rs:437-455
    using S = std::set<Query>;
    auto merge = [](const S& a, const S& b) { S r = a; r.insert(b.begin(), b.end()); return r; };
    Fold<S> f{[](const Fe&) { return S{}; }, [](const CommonPolynomial&) { return S{}; }, [](const Query& q) { return S{q}; },
              [](size_t) { return S{}; }, [](const S& a) { return a; }, merge, merge, [](const S& a, const Fe&) { return a; }};
    return evaluate<S>(e, f);
}

// ---- program builder: the "loader" whose scalars are virtual registers ------------------------------------------------------
class ProgramBuilder {
  public:
    using Val = uint32_t;   // SSA id
    Val input(uint32_t slot) {
        auto it = input_val_.find(slot);
        if (it != input_val_.end()) return it->second;
        n_inputs = std::max<size_t>(n_inputs, slot + 1);
        return input_val_[slot] = emit(SNARKV_FR_OP_INPUT, slot, 0);
    }
    Val constant(const Fe& c) {   // ScalarLoader::load_const, one register per distinct value
        uint32_t ix;
        auto it = const_ix_.find(c);
        if (it == const_ix_.end()) { ix = (uint32_t)consts_.size(); const_ix_[c] = ix; consts_.push_back(c); }
        else ix = it->second;
        auto jt = const_val_.find(ix);
        if (jt != const_val_.end()) return jt->second;
        return const_val_[ix] = emit(SNARKV_FR_OP_CONST, ix, 0);
    }
    Val add(Val a, Val b) { return emit(SNARKV_FR_OP_ADD, a, b); }
    Val sub(Val a, Val b) { return emit(SNARKV_FR_OP_SUB, a, b); }
    Val mul(Val a, Val b) { return emit(SNARKV_FR_OP_MUL, a, b); }
    Val neg(Val a) { return emit(SNARKV_FR_OP_NEG, a, 0); }
    Val inv(Val a) { return emit(SNARKV_FR_OP_INV, a, 0); }
    Val nz(Val a) { return emit(SNARKV_FR_OP_NZ, a, 0); }
    Val keepz(Val a, Val b) { return emit(SNARKV_FR_OP_KEEPZ, a, b); }
    Val pow_const(Val a, uint64_t exp) {   // loader.rs:52-69, the same square-and-multiply order
        if (exp == 0) throw Error("pow_const: exp must be > 0");
        Val base = a;
        while ((exp & 1) == 0) { base = mul(base, base); exp >>= 1; }
        Val acc = base;
        while (exp > 1) {
            exp >>= 1;
            base = mul(base, base);
            if (exp & 1) acc = mul(acc, base);
        }
        return acc;
    }
    std::vector<Val> batch_invert(const std::vector<Val>& values) {   // util/arithmetic.rs:47-74 as straight-line code
        if (values.empty()) return {};
        std::vector<Val> nzv, products, out(values.size());
        for (Val v : values) nzv.push_back(nz(v));
        products.push_back(nzv[0]);
        for (size_t i = 1; i < nzv.size(); ++i) products.push_back(mul(products.back(), nzv[i]));
        Val all_inv = inv(products.back());
        for (size_t i = values.size(); i-- > 0;) {
            const Val inv_i = i > 0 ? mul(all_inv, products[i - 1]) : all_inv;
            if (i > 0) all_inv = mul(all_inv, nzv[i]);
            out[i] = keepz(inv_i, values[i]);
        }
        return out;
    }
    // SSA -> registers by liveness (a value's register is reused after its last use)
    FrProgram finish(const std::vector<Val>& outputs) const {
        const size_t n = ssa_.size();
        std::vector<size_t> last(n);
        for (size_t i = 0; i < n; ++i) last[i] = i;
        auto two = [](uint32_t op) { return op == SNARKV_FR_OP_ADD || op == SNARKV_FR_OP_SUB || op == SNARKV_FR_OP_MUL || op == SNARKV_FR_OP_KEEPZ; };
        auto one = [](uint32_t op) { return op == SNARKV_FR_OP_NEG || op == SNARKV_FR_OP_INV || op == SNARKV_FR_OP_NZ; };
        for (size_t i = 0; i < n; ++i) {
            const Ssa& s = ssa_[i];
            if (two(s.op)) { last[s.a] = std::max(last[s.a], i); last[s.b] = std::max(last[s.b], i); }
            else if (one(s.op)) last[s.a] = std::max(last[s.a], i);
        }
        for (Val o : outputs) last[o] = n;
        std::vector<uint32_t> free_regs, reg(n, 0);
        std::map<size_t, std::vector<size_t>> expire;
        uint32_t n_regs = 0;
        FrProgram p;
        for (size_t i = 0; i < n; ++i) {
            const Ssa& s = ssa_[i];
            uint32_t ra, rb = 0;
            if (s.op == SNARKV_FR_OP_INPUT || s.op == SNARKV_FR_OP_CONST) ra = s.a;
            else if (one(s.op)) ra = reg[s.a];
            else { ra = reg[s.a]; rb = reg[s.b]; }
            auto it = expire.find(i);
            if (it != expire.end()) { for (size_t v : it->second) free_regs.push_back(reg[v]); expire.erase(it); }
            uint32_t r;
            if (!free_regs.empty()) { r = free_regs.back(); free_regs.pop_back(); }
            else r = n_regs++;
            reg[i] = r;
            if (last[i] == i) free_regs.push_back(r);
            else if (last[i] < n) expire[last[i]].push_back(i);
            p.instrs.push_back({s.op, r, ra, rb});
        }
        p.n_regs = std::max<uint32_t>(n_regs, 1);
        for (const Fe& c : consts_) p.consts.push_back(fe_to_bytes(c));
        p.n_inputs = n_inputs;
        for (Val o : outputs) p.outputs.push_back(reg[o]);
        return p;
    }
    size_t n_inputs = 0;

  private:
    struct Ssa { uint32_t op, a, b; };
    Val emit(uint32_t op, uint32_t a, uint32_t b) { ssa_.push_back({op, a, b}); return (Val)(ssa_.size() - 1); }
    std::vector<Ssa> ssa_;
    std::vector<Fe> consts_;
    std::map<Fe, uint32_t> const_ix_;
    std::map<uint32_t, Val> const_val_, input_val_;
};

// protocol.rs:199-283: `new` collects the fractions, all denominators of a proof are inverted together (Lagrange denominators in
// index order, then z^n - 1), then the fractions are evaluated
struct CommonPolynomialEvaluation {
    ProgramBuilder::Val zn, zn_minus_one, zn_minus_one_inv, identity;
    std::map<int32_t, ProgramBuilder::Val> lagrange;
    CommonPolynomialEvaluation(ProgramBuilder& b, const Domain& domain, const std::set<int32_t>& langranges, ProgramBuilder::Val z) {
        zn = b.pow_const(z, domain.n);
        const auto one = b.constant(fe_from_u64(1));
        zn_minus_one = b.sub(zn, one);
        const auto n_inv = b.constant(domain.n_inv);
        const auto numer = b.mul(zn_minus_one, n_inv);
        identity = z;
        std::vector<ProgramBuilder::Val> numers, denoms;
        for (int32_t i : langranges) {
            const auto omega = b.constant(domain.rotate_scalar(fe_from_u64(1), i));
            numers.push_back(b.mul(numer, omega));
            denoms.push_back(b.sub(z, omega));
        }
        denoms.push_back(zn_minus_one);
        const auto inv = b.batch_invert(denoms);
        zn_minus_one_inv = inv.back();
        size_t k = 0;
        for (int32_t i : langranges) { lagrange[i] = b.mul(numers[k], inv[k]); ++k; }
    }
    ProgramBuilder::Val get(const CommonPolynomial& p) const { return p.identity ? identity : lagrange.at(p.index); }
};

// the slice of PlonkProtocol (protocol.rs:22-67) the scalar path reads
struct QuotientProtocol {
    Domain domain;
    size_t num_preprocessed;
    std::vector<size_t> num_instance;
    std::vector<Query> evaluations;
    size_t num_challenge;
    Expr numerator;
    // per-proof input row: [z | challenges | evaluations (protocol.evaluations order) | instances (column-major)]
    size_t off_challenges() const { return 1; }
    size_t off_evaluations() const { return 1 + num_challenge; }
    size_t off_instances() const { return off_evaluations() + evaluations.size(); }
    size_t total_inputs() const { size_t t = off_instances(); for (size_t c : num_instance) t += c; return t; }
};

// outputs per proof: [quotient evaluation, z^n, z^n - 1, 1/(z^n - 1), instance evaluations...]
inline FrProgram compile_quotient_evaluation(const QuotientProtocol& p) {
    using Val = ProgramBuilder::Val;
    ProgramBuilder b;
    const Val z = b.input(0);
    const size_t offset = p.num_preprocessed;
    std::vector<Query> inst_queries;
    for (const Query& q : used_query(*p.numerator))
        if (q.poly >= offset && q.poly < offset + p.num_instance.size()) inst_queries.push_back(q);   // std::set iterates sorted
    std::set<int32_t> lang = used_langrange(*p.numerator);   // PlonkProtocol::langranges (protocol.rs:77-106)
    size_t max_inst = 0;
    for (size_t c : p.num_instance) max_inst = std::max(max_inst, c);
    int32_t min_rot = 0, max_rot = 0;
    for (const Query& q : inst_queries) {
        if (q.rotation < min_rot) min_rot = q.rotation;
        else if (q.rotation > max_rot) max_rot = q.rotation;
    }
    for (int32_t i = -max_rot; i < (int32_t)max_inst + std::abs(min_rot); ++i) lang.insert(i);
    CommonPolynomialEvaluation cpe(b, p.domain, lang, z);
    std::map<Query, Val> evals;
    std::vector<size_t> inst_base;
    { size_t o = p.off_instances(); for (size_t c : p.num_instance) { inst_base.push_back(o); o += c; } }
    std::vector<Val> inst_eval_regs;
    for (const Query& q : inst_queries) {   // proof.rs:313-334 (loader.sum_products)
        const size_t col = q.poly - offset;
        bool have = false;
        Val acc = 0;
        for (size_t j = 0; j < p.num_instance[col]; ++j) {
            const Val inst = b.input((uint32_t)(inst_base[col] + j));
            const Val term = b.mul(inst, cpe.get({false, (int32_t)j - q.rotation}));
            acc = have ? b.add(acc, term) : term;
            have = true;
        }
        if (!have) acc = b.constant(Fe{});
        evals[q] = acc;
        inst_eval_regs.push_back(acc);
    }
    for (size_t k = 0; k < p.evaluations.size(); ++k) evals[p.evaluations[k]] = b.input((uint32_t)(p.off_evaluations() + k));   // proof.rs:336-346
    Fold<Val> f{
        [&](const Fe& c) { return b.constant(c); },
        [&](const CommonPolynomial& cp) { return cpe.get(cp); },
        [&](const Query& q) { auto it = evals.find(q); if (it == evals.end()) throw Error("Missing query"); return it->second; },   // Error::InvalidProtocol
        [&](size_t i) { if (i >= p.num_challenge) throw Error("Missing challenge"); return b.input((uint32_t)(p.off_challenges() + i)); },
        [&](const Val& a) { return b.neg(a); },
        [&](const Val& a, const Val& c) { return b.add(a, c); },
        [&](const Val& a, const Val& c) { return b.mul(a, c); },
        [&](const Val& a, const Fe& c) { const Val k = b.constant(c); return b.mul(a, k); }};
    const Val numerator = evaluate<Val>(*p.numerator, f);
    const Val quotient_eval = b.mul(numerator, cpe.zn_minus_one_inv);   // proof.rs:298-303
    std::vector<Val> outs{quotient_eval, cpe.zn, cpe.zn_minus_one, cpe.zn_minus_one_inv};
    outs.insert(outs.end(), inst_eval_regs.begin(), inst_eval_regs.end());
    b.n_inputs = p.total_inputs();
    return b.finish(outs);
}

// synthetic protocol of StandardPlonk shape (tests / bench); see snark_verifier_b200/plonk_eval.py::standard_plonk_like_protocol
inline QuotientProtocol standard_plonk_like_protocol(int k, size_t num_instance = 1, int blinding_factors = 5) {
    const Expr q_a = polynomial(0), q_b = polynomial(1), q_c = polynomial(2), q_ab = polynomial(3), constant_ = polynomial(4);
    const std::vector<Expr> sigmas{polynomial(5), polynomial(6), polynomial(7)};
    const Expr instance = polynomial(8);
    const std::vector<Expr> advice{polynomial(9), polynomial(10), polynomial(11)};
    const Expr a = advice[0], bb = advice[1], c = advice[2];
    const Expr z = polynomial(12), z_omega = polynomial(12, 1);
    const Expr beta = challenge(1), gamma = challenge(2), alpha = challenge(3);
    const Expr one = constant(fe_from_u64(1));
    const int32_t rotation_last = -(blinding_factors + 1);
    const Expr l_0 = lagrange(0), l_last = lagrange(rotation_last);
    Expr l_blind;
    for (int32_t i = rotation_last + 1; i < 0; ++i) l_blind = l_blind ? l_blind + lagrange(i) : lagrange(i);
    const Expr l_active = one - (l_last + l_blind);
    const Expr identity = common(true);
    const Expr gate = q_a * a + q_b * bb + q_c * c + q_ab * a * bb + constant_ + instance;
    Expr left = z_omega;
    for (size_t i = 0; i < 3; ++i) left = left * (advice[i] + beta * sigmas[i] + gamma);
    Expr right = z;
    Fe delta = fe_from_u64(1);
    const Fe DELTA = fr_delta();
    for (size_t i = 0; i < 3; ++i) {
        right = right * (advice[i] + beta * constant(delta) * identity + gamma);
        delta = fe_mul(delta, DELTA);
    }
    std::vector<Expr> constraints{gate, l_0 * (one - z), l_last * (z * z - z), l_active * (left - right)};
    std::vector<Query> evaluations;
    for (size_t i = 0; i < 8; ++i) evaluations.push_back({i, 0});
    for (size_t i = 0; i < 3; ++i) evaluations.push_back({9 + i, 0});
    evaluations.push_back({12, 0});
    evaluations.push_back({12, 1});
    return QuotientProtocol{Domain(k), 8, {num_instance}, evaluations, 4, distribute_powers(constraints, alpha)};
}

// ---- the multi-open verifier's MSM scalars as a program: `Msm` (util/msm.rs) over virtual registers -----------------------------
// Bases are SLOTS — ('g') the SRS generator carrying the Msm constant, ('c', j) the commitment of polynomial j, ('w', i) the i-th
// opening-proof point: the bases differ from proof to proof, the slot a base occupies in the final MSM does not.  Every statement
// below emits in the order of snark_verifier_b200/plonk_eval.py (SymbolicMsm, compile_*_msm_scalars), which follows the reference's
// evaluation order; tests/test_pcs_mirror.py::test_cpp_msm_scalar_programs_equal_python compares the programs instruction by instruction.
struct Slot {
    char kind;      // 'g', 'c', 'w'
    uint32_t idx;
    bool operator==(const Slot& o) const { return kind == o.kind && idx == o.idx; }
};
class SymbolicMsm {
  public:
    using Val = ProgramBuilder::Val;
    explicit SymbolicMsm(ProgramBuilder& b) : b_(&b) {}
    static SymbolicMsm base(ProgramBuilder& b, Slot s) { SymbolicMsm m(b); m.terms_.push_back({s, b.constant(fe_from_u64(1))}); return m; }
    static SymbolicMsm constant_(ProgramBuilder& b, Val v) { SymbolicMsm m(b); m.has_const_ = true; m.const_ = v; return m; }
    SymbolicMsm operator*(Val k) const {   // scale: constant first, then the terms in order (msm.rs:100-107)
        SymbolicMsm r(*b_);
        if (has_const_) { r.has_const_ = true; r.const_ = b_->mul(const_, k); }
        for (const auto& t : terms_) r.terms_.push_back({t.first, b_->mul(t.second, k)});
        return r;
    }
    SymbolicMsm operator-() const {
        SymbolicMsm r(*b_);
        if (has_const_) { r.has_const_ = true; r.const_ = b_->neg(const_); }
        for (const auto& t : terms_) r.terms_.push_back({t.first, b_->neg(t.second)});
        return r;
    }
    SymbolicMsm operator+(const SymbolicMsm& o) const {   // extend: push dedupes by slot (msm.rs:109-128)
        SymbolicMsm r = *this;
        if (o.has_const_) {
            if (r.has_const_) r.const_ = b_->add(r.const_, o.const_);
            else { r.has_const_ = true; r.const_ = o.const_; }
        }
        for (const auto& t : o.terms_) {
            auto it = std::find_if(r.terms_.begin(), r.terms_.end(), [&](const std::pair<Slot, Val>& x) { return x.first == t.first; });
            if (it == r.terms_.end()) r.terms_.push_back(t);
            else it->second = b_->add(it->second, t.second);
        }
        return r;
    }
    SymbolicMsm operator-(const SymbolicMsm& o) const { const SymbolicMsm n = -o; return *this + n; }
    static SymbolicMsm sum(ProgramBuilder& b, const std::vector<SymbolicMsm>& ms) {
        SymbolicMsm acc(b);
        for (const auto& m : ms) acc = acc + m;
        return acc;
    }
    bool has_const_ = false;
    Val const_ = 0;
    std::vector<std::pair<Slot, Val>> terms_;

  private:
    ProgramBuilder* b_;
};

inline std::vector<ProgramBuilder::Val> powers(ProgramBuilder& b, ProgramBuilder::Val x, size_t n) {   // loader.rs:71-78
    std::vector<ProgramBuilder::Val> out{b.constant(fe_from_u64(1))};
    if (n > 1) out.push_back(x);
    for (size_t i = 2; i < n; ++i) out.push_back(b.mul(out.back(), x));
    out.resize(std::min(out.size(), n));
    return out;
}

// program outputs, per proof: the scalars of the lhs MSM followed by those of the rhs MSM; `lhs_slots` / `rhs_slots` name the bases
struct MsmScalarProgram {
    FrProgram program;
    std::vector<Slot> lhs_slots, rhs_slots;
};
inline MsmScalarProgram finish_msm_program(ProgramBuilder& b, const SymbolicMsm& lhs, const SymbolicMsm& rhs, size_t total_inputs) {
    MsmScalarProgram out;
    std::vector<ProgramBuilder::Val> vals;
    auto flat = [&](const SymbolicMsm& m, std::vector<Slot>& slots) {
        if (m.has_const_) { slots.push_back({'g', 0}); vals.push_back(m.const_); }   // evaluate(Some(gen)): (constant, gen) goes first
        for (const auto& t : m.terms_) { slots.push_back(t.first); vals.push_back(t.second); }
    };
    flat(lhs, out.lhs_slots);
    flat(rhs, out.rhs_slots);
    b.n_inputs = total_inputs;
    out.program = b.finish(vals);
    return out;
}

struct ShiftQuery { size_t poly; Fe shift; };   // (poly, shift) of one query, protocol order

// `Gwc19::verify` (pcs/kzg/multiopen/gwc19.rs:45-82).  Per-proof input row: [z | v | u | one evaluation per query, in query order]
inline MsmScalarProgram compile_gwc19_msm_scalars(const std::vector<ShiftQuery>& queries, size_t num_polys) {
    using Val = ProgramBuilder::Val;
    ProgramBuilder b;
    const Val z = b.input(0), v = b.input(1), u = b.input(2);
    struct Set { Fe shift; std::vector<size_t> polys; std::vector<Val> evals; };
    std::vector<Set> sets;   // gwc19.rs:140-160
    for (size_t k = 0; k < queries.size(); ++k) {
        const Val ev = b.input((uint32_t)(3 + k));
        auto it = std::find_if(sets.begin(), sets.end(), [&](const Set& s) { return s.shift == queries[k].shift; });
        if (it == sets.end()) { sets.push_back({queries[k].shift, {}, {}}); it = sets.end() - 1; }
        it->polys.push_back(queries[k].poly);
        it->evals.push_back(ev);
    }
    size_t max_polys = 0;
    for (const Set& s : sets) max_polys = std::max(max_polys, s.polys.size());
    const std::vector<Val> pu = powers(b, u, sets.size());
    const std::vector<Val> pv = powers(b, v, max_polys);
    std::vector<SymbolicMsm> commitments;
    for (size_t j = 0; j < num_polys; ++j) commitments.push_back(SymbolicMsm::base(b, {'c', (uint32_t)j}));
    SymbolicMsm f(b);
    for (size_t k = 0; k < sets.size(); ++k) {
        SymbolicMsm set_msm(b);   // QuerySet::msm (gwc19.rs:120-137)
        for (size_t i = 0; i < sets[k].polys.size(); ++i) {
            const SymbolicMsm diff = commitments[sets[k].polys[i]] - SymbolicMsm::constant_(b, sets[k].evals[i]);
            const SymbolicMsm scaled = diff * pv[i];
            set_msm = set_msm + scaled;
        }
        const SymbolicMsm su = set_msm * pu[k];
        f = f + su;
    }
    std::vector<Val> z_omegas;
    for (const Set& s : sets) { const Val sh = b.constant(s.shift); z_omegas.push_back(b.mul(sh, z)); }
    std::vector<SymbolicMsm> rhs;
    for (size_t i = 0; i < sets.size(); ++i) { const SymbolicMsm w = SymbolicMsm::base(b, {'w', (uint32_t)i}); rhs.push_back(w * pu[i]); }
    std::vector<SymbolicMsm> uwz;
    for (size_t i = 0; i < sets.size(); ++i) uwz.push_back(rhs[i] * z_omegas[i]);
    const SymbolicMsm uwz_sum = SymbolicMsm::sum(b, uwz);
    const SymbolicMsm lhs = f + uwz_sum;
    const SymbolicMsm rhs_sum = SymbolicMsm::sum(b, rhs);
    return finish_msm_program(b, lhs, rhs_sum, 3 + queries.size());
}

// `Bdfg21::verify` (bdfg21.rs:51-83; query sets :123-175, coefficients :177-371).  Per-proof input row:
// [z | mu | gamma | z' | one evaluation per query, in query order]; ('w', 0) = W, ('w', 1) = W'.
inline MsmScalarProgram compile_bdfg21_msm_scalars(const std::vector<ShiftQuery>& queries, size_t num_polys) {
    using Val = ProgramBuilder::Val;
    ProgramBuilder b;
    const Val z = b.input(0), mu = b.input(1), gamma = b.input(2), z_prime = b.input(3);
    struct PolyShifts { size_t poly; std::vector<Fe> shifts; std::vector<Val> evals; };
    std::vector<PolyShifts> ps;
    for (size_t k = 0; k < queries.size(); ++k) {
        const Val ev = b.input((uint32_t)(4 + k));
        auto it = std::find_if(ps.begin(), ps.end(), [&](const PolyShifts& p) { return p.poly == queries[k].poly; });
        if (it == ps.end()) ps.push_back({queries[k].poly, {queries[k].shift}, {ev}});
        else if (std::find(it->shifts.begin(), it->shifts.end(), queries[k].shift) == it->shifts.end()) { it->shifts.push_back(queries[k].shift); it->evals.push_back(ev); }
    }
    struct Set { std::vector<Fe> shifts; std::vector<size_t> polys; std::vector<std::vector<Val>> evals; };
    std::vector<Set> sets;
    for (const PolyShifts& p : ps) {
        auto same = [&](const Set& s) {
            if (s.shifts.size() != p.shifts.size()) return false;
            for (const Fe& x : p.shifts) if (std::find(s.shifts.begin(), s.shifts.end(), x) == s.shifts.end()) return false;
            return true;
        };
        auto it = std::find_if(sets.begin(), sets.end(), same);
        if (it == sets.end()) { sets.push_back({p.shifts, {p.poly}, {p.evals}}); continue; }
        if (std::find(it->polys.begin(), it->polys.end(), p.poly) != it->polys.end()) continue;
        it->polys.push_back(p.poly);
        std::vector<Val> ordered;
        for (const Fe& lhs : it->shifts) ordered.push_back(p.evals[std::find(p.shifts.begin(), p.shifts.end(), lhs) - p.shifts.begin()]);
        it->evals.push_back(ordered);
    }
    // query_set_coeffs
    std::vector<Fe> superset;
    for (const Set& s : sets) for (const Fe& x : s.shifts) if (std::find(superset.begin(), superset.end(), x) == superset.end()) superset.push_back(x);
    std::sort(superset.begin(), superset.end());
    size_t size = 2;
    for (const Set& s : sets) size = std::max(size, s.shifts.size());
    const std::vector<Val> powers_of_z = powers(b, z, size);
    std::vector<std::pair<Fe, Val>> zp_minus;
    for (const Fe& x : superset) { const Val c = b.constant(x); const Val zx = b.mul(z, c); zp_minus.push_back({x, b.sub(z_prime, zx)}); }
    auto zp_of = [&](const Fe& x) { for (auto& kv : zp_minus) if (kv.first == x) return kv.second; throw Error("shift not in superset"); };
    struct Coeff { std::vector<Val> bary, eval_coeffs; Val z_s = 0; bool has_z_s_1 = false; Val z_s_1 = 0; bool has_cc = false; Val commitment_coeff = 0, r_eval_coeff = 0; };
    std::vector<Coeff> coeffs;
    bool have_first = false;
    Val first_z_s = 0;
    for (const Set& s : sets) {
        Coeff c;
        const Val zz = powers_of_z[1], z_pow = powers_of_z[s.shifts.size() - 1];
        for (size_t j = 0; j < s.shifts.size(); ++j) {
            Fe ell = fe_from_u64(1);
            for (size_t i = 0; i < s.shifts.size(); ++i) if (i != j) ell = fe_mul(ell, fe_sub(s.shifts[j], s.shifts[i]));
            const Val m1 = b.mul(z_pow, z_prime);
            const Val c1 = b.constant(ell);
            const Val t1 = b.mul(m1, c1);
            const Val m2 = b.mul(z_pow, zz);
            const Val c2 = b.constant(fe_neg(fe_mul(ell, s.shifts[j])));
            const Val t2 = b.mul(m2, c2);
            c.bary.push_back(b.add(t1, t2));
        }
        bool have = false;
        Val z_s = 0;
        for (const Fe& x : s.shifts) { z_s = have ? b.mul(z_s, zp_of(x)) : zp_of(x); have = true; }
        c.z_s = z_s;
        if (have_first) { c.has_z_s_1 = true; c.z_s_1 = first_z_s; }
        else { have_first = true; first_z_s = z_s; }
        coeffs.push_back(c);
    }
    std::vector<Val> den1;
    for (const Coeff& c : coeffs) { den1.insert(den1.end(), c.bary.begin(), c.bary.end()); if (c.has_z_s_1) den1.push_back(c.z_s); }
    const std::vector<Val> inv1 = b.batch_invert(den1);
    size_t k = 0;
    std::vector<Val> den2;
    for (Coeff& c : coeffs) {
        c.eval_coeffs.assign(inv1.begin() + k, inv1.begin() + k + c.bary.size());
        k += c.bary.size();
        if (c.has_z_s_1) { c.has_cc = true; c.commitment_coeff = b.mul(c.z_s_1, inv1[k]); ++k; }
        Val wsum = c.eval_coeffs[0];
        for (size_t i = 1; i < c.eval_coeffs.size(); ++i) wsum = b.add(wsum, c.eval_coeffs[i]);
        den2.push_back(wsum);
    }
    const std::vector<Val> inv2 = b.batch_invert(den2);
    for (size_t i = 0; i < coeffs.size(); ++i) coeffs[i].r_eval_coeff = coeffs[i].has_cc ? b.mul(coeffs[i].commitment_coeff, inv2[i]) : inv2[i];
    // verify
    size_t max_polys = 0;
    for (const Set& s : sets) max_polys = std::max(max_polys, s.polys.size());
    const std::vector<Val> powers_of_mu = powers(b, mu, max_polys);
    const std::vector<Val> powers_of_gamma = powers(b, gamma, sets.size());
    std::vector<SymbolicMsm> commitments;
    for (size_t j = 0; j < num_polys; ++j) commitments.push_back(SymbolicMsm::base(b, {'c', (uint32_t)j}));
    SymbolicMsm f(b);
    for (size_t si = 0; si < sets.size(); ++si) {
        const Set& st = sets[si];
        const Coeff& co = coeffs[si];
        SymbolicMsm set_msm(b);
        for (size_t i = 0; i < st.polys.size(); ++i) {
            const SymbolicMsm commitment = co.has_cc ? commitments[st.polys[i]] * co.commitment_coeff : commitments[st.polys[i]];
            Val r_eval = b.mul(co.eval_coeffs[0], st.evals[i][0]);
            for (size_t j = 1; j < co.eval_coeffs.size(); ++j) { const Val t = b.mul(co.eval_coeffs[j], st.evals[i][j]); r_eval = b.add(r_eval, t); }
            r_eval = b.mul(r_eval, co.r_eval_coeff);
            const SymbolicMsm diff = commitment - SymbolicMsm::constant_(b, r_eval);
            const SymbolicMsm scaled = diff * powers_of_mu[i];
            set_msm = set_msm + scaled;
        }
        const SymbolicMsm sg = set_msm * powers_of_gamma[si];
        f = f + sg;
    }
    const SymbolicMsm w_base = SymbolicMsm::base(b, {'w', 0});
    const SymbolicMsm wz = w_base * coeffs[0].z_s;
    f = f - wz;
    const SymbolicMsm rhs = SymbolicMsm::base(b, {'w', 1});
    const SymbolicMsm rz = rhs * z_prime;
    const SymbolicMsm lhs = f + rz;
    return finish_msm_program(b, lhs, rhs, 4 + queries.size());
}

}  // namespace plonk
}  // namespace snarkv
