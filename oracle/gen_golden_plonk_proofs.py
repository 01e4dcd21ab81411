"""Writes tests/golden/plonk_proofs.json: genuine proofs of the small hand-built PLONK protocol (`snark_verifier_b200.plonk.
simple_plonk_protocol`) made by the from-scratch prover tests/plonk_toy.py under a fixed SRS — 8 different public inputs each for
GWC19 and SHPLONK over the Keccak EvmTranscript and for SHPLONK over the Poseidon transcript, plus one tampered proof per list.  bench.py replicates them to a 4096-proof batch for the "real proofs" leg of
BASELINE config 3 (it may not import the prover: the prover uses the CPU oracle).  TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import plonk_toy as T  # noqa: E402

K, SRS_SEED, CIRCUIT_SEED, INITIAL_STATE = 6, 2026, 77, 0xC0DE


def main():
    srs = T.Srs(SRS_SEED)
    out = {"k": K, "num_public": 2, "initial_state": INITIAL_STATE, "svk_g": T.GEN.hex(), "g2": srs.g2.hex(), "s_g2": srs.s_g2.hex(),
           "what": "tests/plonk_toy.py proofs for snark_verifier_b200.plonk.simple_plonk_protocol(k, preprocessed, 2, None, initial_state)",
           "preprocessed": None, "gwc19": [], "bdfg21": [], "bdfg21_poseidon": []}
    for j in range(8):
        circ = T.Circuit(K, CIRCUIT_SEED, [1000 + j, 77 * j + 5])      # same seed: same selector columns, other public inputs
        protocol = T.make_protocol(circ, srs, None, INITIAL_STATE)
        pre = [p.hex() for p in protocol.preprocessed]
        assert out["preprocessed"] in (None, pre)
        out["preprocessed"] = pre
        for scheme in ("gwc19", "bdfg21"):
            out[scheme].append({"instances": [[str(v) for v in circ.public]], "proof": T.prove(circ, protocol, srs, scheme).hex(), "valid": True})
        out["bdfg21_poseidon"].append({"instances": [[str(v) for v in circ.public]],
                                       "proof": T.prove(circ, protocol, srs, "bdfg21", transcript="poseidon").hex(), "valid": True})
    circ = T.Circuit(K, CIRCUIT_SEED, [1, 2])
    protocol = T.make_protocol(circ, srs, None, INITIAL_STATE)
    for scheme in ("gwc19", "bdfg21"):
        out[scheme].append({"instances": [[str(v) for v in circ.public]], "proof": T.prove(circ, protocol, srs, scheme, tamper="evaluation").hex(),
                            "valid": False})
    out["bdfg21_poseidon"].append({"instances": [[str(v) for v in circ.public]],
                                   "proof": T.prove(circ, protocol, srs, "bdfg21", tamper="evaluation", transcript="poseidon").hex(), "valid": False})
    path = os.path.join(ROOT, "tests", "golden", "plonk_proofs.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, {s: len(out[s]) for s in ("gwc19", "bdfg21", "bdfg21_poseidon")})


if __name__ == "__main__":
    main()
