"""CPU-only: csrc/glv.cuh (GLV scalar decomposition used by the device MSM for n < 2^22) is plain __host__ __device__ C; the exact
code is compiled with g++ and checked against the defining identities with Python big integers."""
import ctypes
import os
import random
import subprocess

from oracle import bn254_model as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "snark_verifier_b200", "csrc")
LAMBDA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23
BETA = 0x30644E72E131A0295E6DD9E7E0ACCCB0C28F069FBB966E3DE4BD44E5607CFD48


def test_endomorphism_constants():
    assert pow(LAMBDA, 3, m.R) == 1 and LAMBDA != 1 and pow(BETA, 3, m.P) == 1 and BETA != 1
    for k in (1, 2, 12345678901234567890):
        pt = m.g1_mul(m.G1_GEN, k)
        assert m.g1_mul(pt, LAMBDA) == (BETA * pt[0] % m.P, pt[1])          # phi(P) = (beta x, y) = [lambda] P


def test_decomposition_matches_identity_and_bounds(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "glv.cuh"\n#include <cstring>\n'
                   'extern "C" void dec(const uint32_t* k, uint32_t* out) {\n'
                   '  uint32_t m1[5], m2[5], n1, n2; snarkv::glv::decompose(k, m1, n1, m2, n2);\n'
                   '  memcpy(out, m1, 20); memcpy(out + 5, m2, 20); out[10] = n1; out[11] = n2; }\n')
    so = tmp_path / "libg.so"
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", CSRC, "-o", str(so), str(src)])
    lib = ctypes.CDLL(str(so))
    rnd = random.Random(9)
    cases = [0, 1, 2, 3, m.R - 1, m.R - 2, LAMBDA, m.R - LAMBDA, (m.R - 1) // 2, 1 << 253, (1 << 128) - 1, 1 << 128]
    cases += [rnd.randrange(m.R) for _ in range(20000)]
    # adversarial: scalars whose quotients k b2 / r, k |b1| / r sit next to a rounding boundary (largest residuals)
    B2, B1ABS = 0x6F4D8248EEB859FD0BE4E1541221250B, 0x89D3256894D213E3
    for den in (B2, B1ABS):
        for _ in range(2000):
            j = rnd.randrange(den)
            k0 = ((2 * j + 1) * m.R) // (2 * den)
            cases += [(k0 + d) % m.R for d in (-1, 0, 1)]
    worst = 0
    for k in cases:
        out = (ctypes.c_uint32 * 12)()
        lib.dec(k.to_bytes(32, "little"), out)
        m1 = sum(out[i] << (32 * i) for i in range(5)); m2 = sum(out[5 + i] << (32 * i) for i in range(5))
        k1 = -m1 if out[10] else m1
        k2 = -m2 if out[11] else m2
        assert (k1 + k2 * LAMBDA - k) % m.R == 0, hex(k)
        worst = max(worst, m1.bit_length(), m2.bit_length())
    assert worst <= 128, worst            # the device digit loop covers ceil(130 / c) windows
