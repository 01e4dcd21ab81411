"""CPU-only: csrc/fp_inv.cuh (binary extended-GCD inverse used on the device's serial critical paths) is plain
__host__ __device__ C, so the exact code is compiled with g++ here and checked against Python's pow(a, -1, m); and the
PTX generator's CPU emulation of the Montgomery multiply/add/sub blocks is re-run."""
import ctypes
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "snark_verifier_b200", "csrc")
P = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001


def test_binary_gcd_inverse_matches_python(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "fp_inv.cuh"\n#include <cstring>\nusing namespace snarkv;\n'
                   'extern "C" void inv_mod(const uint32_t* a, const uint32_t* p, uint32_t* out) {\n'
                   '  U256 x, m; memcpy(x.v, a, 32); memcpy(m.v, p, 32); U256 r = u256_inv_mod(x, m); memcpy(out, r.v, 32); }\n'
                   'extern "C" void inv_mod_fast(const uint32_t* a, const uint32_t* p, uint32_t* out) {\n'
                   '  U256 x, m; memcpy(x.v, a, 32); memcpy(m.v, p, 32); U256 r = u256_inv_mod_fast(x, m); memcpy(out, r.v, 32); }\n')
    so = tmp_path / "libt.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", CSRC, "-o", str(so), str(src)])
    lib = ctypes.CDLL(str(so))
    rnd = random.Random(11)
    for mod in (P, R):
        cases = [1, 2, 3, mod - 1, mod - 2, (mod + 1) // 2, 1 << 253] + [rnd.randrange(1, mod) for _ in range(500)]
        cases += [1 << k for k in range(1, 254)] + [mod - (1 << k) for k in range(1, 253)] + [rnd.randrange(1, 1 << 64) for _ in range(100)]
        for fn in (lib.inv_mod, lib.inv_mod_fast):      # bit-at-a-time loop and the 31-steps-at-a-time variant (Pornin)
            for a in cases:
                out = ctypes.create_string_buffer(32)
                fn(a.to_bytes(32, "little"), mod.to_bytes(32, "little"), out)
                assert int.from_bytes(out.raw, "little") == pow(a, -1, mod)
            out = ctypes.create_string_buffer(32)
            fn(bytes(32), mod.to_bytes(32, "little"), out)
            assert out.raw == bytes(32)          # inverse of zero is zero (halo2curves' CtOption::None is handled by callers)


def test_generated_ptx_blocks_pass_cpu_emulation():
    out = subprocess.run([sys.executable, os.path.join(CSRC, "gen_field_ptx.py"), "--selftest"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.count("ok") == 4, out.stdout + out.stderr   # BN254 Fq / Fr + Pallas base / scalar field


def test_committed_generated_headers_are_current():
    for gen, inc, args in (("gen_field_ptx.py", "fp_ptx.inc", []), ("gen_field_ptx.py", "fp_ptx_pallas.inc", ["--curve=pallas"]),
                           ("gen_pairing_consts.py", "pairing_consts.inc", []), ("gen_poseidon_consts.py", "poseidon_consts.inc", [])):
        out = subprocess.run([sys.executable, os.path.join(CSRC, gen)] + args, capture_output=True, text=True, check=True).stdout
        assert out == open(os.path.join(CSRC, inc)).read(), inc + " is stale: regenerate"
