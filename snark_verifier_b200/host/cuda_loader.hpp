// cuda_loader.hpp — C++ host mirror of the reference's Loader / Msm / Decider surface over the C ABI (include/snarkv_cuda.h).
//
// The reference is Rust; this image has no Rust toolchain, so the host side above the C ABI is C++ with the reference's names
// and semantics (paths relative to /root/reference/snark-verifier/src):
//   CudaLoader::multi_scalar_multiplication  <->  EcPointLoader::multi_scalar_multiplication  loader.rs:108-113, native.rs:61-71
//   CudaLoader::ec_point_assert_eq           <->  native.rs:50-59
//   Msm                                      <->  util/msm.rs:20-128 (constant / scalars / bases, scale, push w/ dedupe, evaluate)
//   KzgDecidingKey, KzgAccumulator           <->  pcs/kzg/decider.rs:6-42, pcs/kzg/accumulator.rs:6-26
//   KzgAs::{decide, decide_all, verify}      <->  pcs/kzg/decider.rs:70-93, pcs/kzg/accumulation.rs:41-63
//   AssertionFailure                         <->  Error::AssertionFailure (lib.rs:18-28)
//   Query, Gwc19::verify                     <->  pcs.rs:20-49, pcs/kzg/multiopen/gwc19.rs:45-82 (SHPLONK: host/pcs.hpp)
//   LimbsEncoding<LIMBS, BITS>::from_repr    <->  pcs/kzg/accumulator.rs:57-81 (+ util/arithmetic.rs:270-282), also for a batch
//   FrProgram / CudaLoader::fr_program_eval  <->  verifier/plonk/protocol.rs:211-283, 336-392; proof.rs:298-349 for a batch of proofs
// Loaded values are plain host values exactly as NativeLoader keeps them (native.rs:44,75): Fr = 32 bytes, G1Affine = 64 bytes,
// canonical little-endian (SNARKV_CANONICAL).  There is no arithmetic in this header except the Fr bookkeeping of `Msm`, which
// the caller supplies through the tiny `FrOps` policy (the reference does that part with halo2curves' Fr on the host as well).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/snarkv_cuda.h"

namespace snarkv {

using Fr = std::array<uint8_t, 32>;        // PrimeField::to_repr()
using G1Affine = std::array<uint8_t, 64>;  // x || y, identity = (0,0)
using G2Affine = std::array<uint8_t, 128>;
using Gt = std::array<uint8_t, 384>;

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
struct AssertionFailure : Error { using Error::Error; };   // Error::AssertionFailure(String)
struct CudaError : Error { using Error::Error; };          // no device / runtime failure: never replaced by a CPU path

class CudaLoader {
  public:
    explicit CudaLoader(int device = 0) {
        if (snarkv_init(device, &ctx_) != SNARKV_OK || !ctx_) throw CudaError("snarkv_init failed: no sm_100 GPU (no CPU fallback)");
    }
    ~CudaLoader() { snarkv_destroy(ctx_); }
    CudaLoader(const CudaLoader&) = delete;
    CudaLoader& operator=(const CudaLoader&) = delete;

    G1Affine ec_point_load_const(const G1Affine& value) const { return value; }
    void ec_point_assert_eq(const std::string& annotation, const G1Affine& lhs, const G1Affine& rhs) const {
        if (lhs != rhs) throw AssertionFailure(annotation);
    }
    // pairs: (&scalar, &base) exactly like `&[(&LoadedScalar, &LoadedEcPoint)]`
    G1Affine multi_scalar_multiplication(const std::vector<std::pair<const Fr*, const G1Affine*>>& pairs) {
        std::vector<uint8_t> s(pairs.size() * 32), p(pairs.size() * 64);
        for (size_t i = 0; i < pairs.size(); ++i) {
            memcpy(&s[32 * i], pairs[i].first->data(), 32);
            memcpy(&p[64 * i], pairs[i].second->data(), 64);
        }
        return msm(s.data(), p.data(), pairs.size());
    }
    G1Affine msm(const uint8_t* scalars, const uint8_t* points, size_t n, int flags = 0) {
        G1Affine out;
        check(snarkv_g1_msm(ctx_, scalars, points, n, SNARKV_CANONICAL, flags, out.data()), "multi_scalar_multiplication");
        return out;
    }
    snarkv_ctx* raw() { return ctx_; }
    void check(int rc, const char* what) {
        if (rc == SNARKV_OK) return;
        std::string msg = std::string(what) + ": " + snarkv_last_error(ctx_) + " (rc=" + std::to_string(rc) + ")";
        if (rc == SNARKV_ERR_CUDA) throw CudaError(msg);
        throw Error(msg);
    }

  private:
    snarkv_ctx* ctx_ = nullptr;
};

// Fr bookkeeping policy for Msm: the caller provides add/mul on canonical 32-byte scalars (any bignum will do).
struct FrOps {
    Fr (*add)(const Fr&, const Fr&);
    Fr (*mul)(const Fr&, const Fr&);
    Fr one;
    Fr (*neg)(const Fr&) = nullptr;   // needed by Msm::operator- / the multi-open verifiers below
    Fr (*inv)(const Fr&) = nullptr;   // needed by Bdfg21 (host/pcs.hpp); must map 0 to 0 (batch_invert leaves zeros untouched)
};

// util/msm.rs:20-24.  `Loader` only needs `multi_scalar_multiplication(pairs)`: CudaLoader in production, a recording stand-in in the
// CPU tests of the host logic (tests/pcs_mirror_test.cpp).
template <class Loader>
class BasicMsm {
  public:
    using Msm = BasicMsm;
    BasicMsm(Loader& loader, const FrOps& ops) : loader_(&loader), ops_(&ops) {}
    static Msm constant(Loader& l, const FrOps& o, const Fr& c) { Msm m(l, o); m.constant_ = c; return m; }   // :46-52
    static Msm base(Loader& l, const FrOps& o, const G1Affine& b) {                                              // :54-61
        Msm m(l, o); m.scalars_.push_back(o.one); m.bases_.push_back(b); return m;
    }
    size_t size() const { return bases_.size(); }
    Msm& scale(const Fr& factor) {                                                                                   // :100-107
        if (constant_) constant_ = ops_->mul(*constant_, factor);
        for (auto& s : scalars_) s = ops_->mul(s, factor);
        return *this;
    }
    void push(const Fr& scalar, const G1Affine& base) {                                                              // :109-116
        for (size_t i = 0; i < bases_.size(); ++i)
            if (bases_[i] == base) { scalars_[i] = ops_->add(scalars_[i], scalar); return; }
        scalars_.push_back(scalar);
        bases_.push_back(base);
    }
    Msm& extend(const Msm& other) {                                                                                  // :118-128
        if (other.constant_) constant_ = constant_ ? ops_->add(*constant_, *other.constant_) : *other.constant_;
        for (size_t i = 0; i < other.bases_.size(); ++i) push(other.scalars_[i], other.bases_[i]);
        return *this;
    }
    Msm operator+(const Msm& o) const { Msm r = *this; r.extend(o); return r; }                                      // :130-154
    Msm operator*(const Fr& k) const { Msm r = *this; r.scale(k); return r; }                                        // :180-190
    Msm operator-() const {                                                                                          // :192-204
        if (!ops_->neg) throw Error("Msm: FrOps::neg is required for subtraction");
        Msm r = *this;
        if (r.constant_) r.constant_ = ops_->neg(*r.constant_);
        for (auto& s : r.scalars_) s = ops_->neg(s);
        return r;
    }
    Msm operator-(const Msm& o) const { return *this + (-o); }                                                       // :156-178
    // evaluate(gen): prepend (constant, gen) and call the loader's MSM                                               // :81-98
    G1Affine evaluate(const std::optional<G1Affine>& gen) const {
        std::vector<std::pair<const Fr*, const G1Affine*>> pairs;
        if (constant_) {
            if (!gen) throw Error("Msm::evaluate: constant term without a generator");  // the reference unwraps None
            pairs.emplace_back(&*constant_, &*gen);
        }
        for (size_t i = 0; i < bases_.size(); ++i) pairs.emplace_back(&scalars_[i], &bases_[i]);
        return loader_->multi_scalar_multiplication(pairs);
    }

  private:
    Loader* loader_;
    const FrOps* ops_;
    std::optional<Fr> constant_;
    std::vector<Fr> scalars_;
    std::vector<G1Affine> bases_;
};
using Msm = BasicMsm<CudaLoader>;

struct KzgAccumulator { G1Affine lhs, rhs; };            // pcs/kzg/accumulator.rs:6-26
struct KzgDecidingKey { G1Affine g; G2Affine g2, s_g2; };  // pcs/kzg/decider.rs:6-42 (svk.g, g2, s_g2)

// KzgAs<Bn256, MOS> restricted to the hot path
class KzgAs {
  public:
    static constexpr const char* ASSERTION = "e(lhs, g2)\xC2\xB7" "e(rhs, -s_g2) == O";   // decider.rs:81
    KzgAs(CudaLoader& loader, const KzgDecidingKey& dk) : loader_(&loader) {
        loader.check(snarkv_kzg_set_deciding_key(loader.raw(), dk.g.data(), dk.g2.data(), dk.s_g2.data()), "KzgDecidingKey");
    }
    void decide(const KzgAccumulator& acc) { decide_all({acc}); }                            // decider.rs:70-82
    void decide_all(const std::vector<KzgAccumulator>& accs) {                              // decider.rs:84-93
        if (accs.empty()) return;
        std::vector<uint8_t> lhs(accs.size() * 64), rhs(accs.size() * 64), accept(accs.size());
        for (size_t i = 0; i < accs.size(); ++i) {
            memcpy(&lhs[64 * i], accs[i].lhs.data(), 64);
            memcpy(&rhs[64 * i], accs[i].rhs.data(), 64);
        }
        loader_->check(snarkv_kzg_decide_batch(loader_->raw(), lhs.data(), rhs.data(), accs.size(), SNARKV_CANONICAL, accept.data(), nullptr),
                       "decide");
        for (uint8_t a : accept)
            if (a != 1) throw AssertionFailure(ASSERTION);
    }
    // AccumulationScheme::verify: (sum r^i lhs_i, sum r^i rhs_i)                              accumulation.rs:41-63
    KzgAccumulator verify(const std::vector<KzgAccumulator>& instances, const Fr& r) {
        std::vector<uint8_t> lhs(instances.size() * 64), rhs(instances.size() * 64);
        for (size_t i = 0; i < instances.size(); ++i) {
            memcpy(&lhs[64 * i], instances[i].lhs.data(), 64);
            memcpy(&rhs[64 * i], instances[i].rhs.data(), 64);
        }
        KzgAccumulator out;
        loader_->check(snarkv_kzg_accumulate(loader_->raw(), lhs.data(), rhs.data(), instances.size(), r.data(), SNARKV_CANONICAL,
                                             out.lhs.data(), out.rhs.data()), "KzgAs::verify");
        return out;
    }

  private:
    CudaLoader* loader_;
};

// pcs.rs:20-49
struct Query { size_t poly; Fr shift; Fr eval; };

// GWC19 multi-open verifier (pcs/kzg/multiopen/gwc19.rs:45-82): builds the (lhs, rhs) Msm pair of the accumulator from the
// commitments, the queries and the proof (v, W_i, u); the two `evaluate` calls are the loader's MSM.
struct Gwc19Proof { Fr v; std::vector<G1Affine> ws; Fr u; };
template <class Loader>
struct BasicGwc19 {
    using Msm = BasicMsm<Loader>;
    static std::vector<Fr> powers(const FrOps& o, const Fr& x, size_t n) {                                           // loader.rs:71-78
        std::vector<Fr> out;
        Fr acc = o.one;
        for (size_t i = 0; i < n; ++i) { out.push_back(acc); acc = o.mul(acc, x); }
        return out;
    }
    static KzgAccumulator verify(Loader& loader, const FrOps& o, const G1Affine& svk_g, const std::vector<Msm>& commitments, const Fr& z,
                                 const std::vector<Query>& queries, const Gwc19Proof& proof) {
        struct Set { Fr shift; std::vector<size_t> polys; std::vector<Fr> evals; };
        std::vector<Set> sets;                                                                                       // gwc19.rs:140-160
        for (const Query& q : queries) {
            Set* hit = nullptr;
            for (Set& s : sets) if (s.shift == q.shift) { hit = &s; break; }
            if (!hit) { sets.push_back({q.shift, {}, {}}); hit = &sets.back(); }
            hit->polys.push_back(q.poly);
            hit->evals.push_back(q.eval);
        }
        size_t max_polys = 0;
        for (const Set& s : sets) max_polys = s.polys.size() > max_polys ? s.polys.size() : max_polys;
        const std::vector<Fr> pu = powers(o, proof.u, sets.size()), pv = powers(o, proof.v, max_polys);
        Msm f(loader, o);
        for (size_t k = 0; k < sets.size(); ++k) {
            Msm set_msm(loader, o);                                                                                  // QuerySet::msm, gwc19.rs:120-137
            for (size_t i = 0; i < sets[k].polys.size(); ++i)
                set_msm = set_msm + (commitments[sets[k].polys[i]] - Msm::constant(loader, o, sets[k].evals[i])) * pv[i];
            f = f + set_msm * pu[k];
        }
        Msm lhs = f, rhs(loader, o);
        for (size_t k = 0; k < sets.size(); ++k) {
            const Msm uw = Msm::base(loader, o, proof.ws[k]) * pu[k];
            lhs = lhs + uw * o.mul(sets[k].shift, z);                                                                // z_omega = shift * z
            rhs = rhs + uw;
        }
        return KzgAccumulator{lhs.evaluate(svk_g), rhs.evaluate(svk_g)};
    }
};
using Gwc19 = BasicGwc19<CudaLoader>;

// `LimbsEncoding<LIMBS, BITS>` (pcs/kzg/accumulator.rs:28-82): an accumulator as 4 x LIMBS scalar-field limbs.  The reference
// panics when the limbs do not encode two curve points; here that is an `Error`.
template <uint32_t LIMBS, uint32_t BITS>
struct LimbsEncoding {
    static std::vector<KzgAccumulator> from_repr_batch(CudaLoader& loader, const std::vector<Fr>& limbs, std::vector<uint8_t>* valid_out = nullptr) {
        if (limbs.size() % (4 * LIMBS)) throw Error("LimbsEncoding::from_repr: limbs.len() must be a multiple of 4 * LIMBS");
        const size_t m = limbs.size() / (4 * LIMBS);
        std::vector<uint8_t> lhs(m * 64), rhs(m * 64), valid(m);
        loader.check(snarkv_kzg_accumulators_from_limbs(loader.raw(), limbs.empty() ? nullptr : limbs[0].data(), m, LIMBS, BITS, SNARKV_CANONICAL,
                                                        lhs.data(), rhs.data(), valid.data()), "LimbsEncoding::from_repr");
        std::vector<KzgAccumulator> out(m);
        for (size_t a = 0; a < m; ++a) {
            if (!valid[a] && !valid_out) throw Error("LimbsEncoding::from_repr: limbs do not encode two G1 points");   // from_xy(..).unwrap()
            memcpy(out[a].lhs.data(), &lhs[64 * a], 64);
            memcpy(out[a].rhs.data(), &rhs[64 * a], 64);
        }
        if (valid_out) *valid_out = valid;
        return out;
    }
    static KzgAccumulator from_repr(CudaLoader& loader, const std::vector<Fr>& limbs) {                           // accumulator.rs:57-81
        if (limbs.size() != 4 * LIMBS) throw Error("LimbsEncoding::from_repr: expected 4 * LIMBS limbs");        // assert_eq!(limbs.len(), 4 * LIMBS)
        return from_repr_batch(loader, limbs)[0];
    }
};

// A straight-line Fr register program (include/snarkv_cuda.h SNARKV_FR_OP_*): what `Expression::evaluate` +
// `CommonPolynomialEvaluation` flatten to for ONE protocol (the compiler lives in snark_verifier_b200/plonk_eval.py).
struct FrProgram {
    std::vector<snarkv_fr_instr> instrs;
    uint32_t n_regs = 0;
    std::vector<Fr> consts;
    size_t n_inputs = 0;
    std::vector<uint32_t> outputs;
    // inputs: m x n_inputs scalars -> m x outputs.size() scalars
    std::vector<Fr> eval_batch(CudaLoader& loader, const std::vector<Fr>& inputs, size_t m) const {
        if (inputs.size() != m * n_inputs) throw Error("FrProgram::eval_batch: inputs.len() != m * n_inputs");
        std::vector<Fr> out(m * outputs.size());
        loader.check(snarkv_fr_program_eval_batch(loader.raw(), instrs.data(), instrs.size(), n_regs, consts.empty() ? nullptr : consts[0].data(),
                                                  consts.size(), inputs.empty() ? nullptr : inputs[0].data(), n_inputs, m, outputs.data(),
                                                  outputs.size(), SNARKV_CANONICAL, out.empty() ? nullptr : out[0].data()), "FrProgram::eval_batch");
        return out;
    }
};

}  // namespace snarkv
