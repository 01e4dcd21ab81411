"""Pallas side of the path: the IPA decider and the large MSM behind it, over the same C ABI.

Mirrored reference items (paths relative to snark-verifier/src):
  pcs/ipa/decider.rs:3-22     IpaDecidingKey { svk, g }              -> IpaDecidingKey (g is uploaded ONCE and stays in HBM)
  pcs/ipa/accumulator.rs      IpaAccumulator { xi, u }               -> IpaAccumulator
  pcs/ipa/decider.rs:47-70    IpaAs::{decide, decide_all}            -> IpaAs (error string "U == commit(G, h)")
  pcs/ipa.rs:401-417          h_coeffs                               -> PallasLoader.h_coeffs (parity entry; decide computes it on the device)
  util/msm.rs:308-343         multi_scalar_multiplication for Pallas -> PallasLoader.msm / msm_device
Values are plain bytes like everywhere in this package: scalars 32 B, affine points 64 B (x || y, identity (0, 0)), canonical LE."""
import ctypes
from dataclasses import dataclass
from typing import List, Sequence

from . import CANONICAL, AssertionFailure, CudaLoader, _addr

PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
PALLAS_Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
PALLAS_GENERATOR = (PALLAS_P - 1).to_bytes(32, "little") + (2).to_bytes(32, "little")


class PallasLoader:
    """Pallas entry points on a CudaLoader's context (one context serves both curves; its workspace is shared scratch)."""

    def __init__(self, loader: CudaLoader):
        self.loader, self.lib, self.h, self.fmt = loader, loader.lib, loader.h, loader.fmt

    def msm(self, scalars, points, n, flags=0):
        out = ctypes.create_string_buffer(64)
        self.loader._check(self.lib.snarkv_pallas_msm(self.h, _addr(scalars), _addr(points), n, self.fmt, flags, out), "pallas msm")
        return out.raw

    def msm_device(self, d_scalars, d_points, n, d_out_affine=None, d_out_jacobian=None, flags=0, d_status=None):
        self.loader._check(self.lib.snarkv_pallas_msm_device(self.h, _addr(d_scalars), _addr(d_points), n, self.fmt, flags, _addr(d_out_affine),
                                                             _addr(d_out_jacobian), _addr(d_status)), "pallas msm_device")

    def h_coeffs(self, xi: bytes, k: int, scalar: bytes = (1).to_bytes(32, "little")):
        out = ctypes.create_string_buffer(32 << k)
        self.loader._check(self.lib.snarkv_pallas_h_coeffs(self.h, _addr(xi), k, bytes(scalar), self.fmt, out), "h_coeffs")
        return out.raw

    def synth_scalars_device(self, seed, start, n, d_out):
        self.loader._check(self.lib.snarkv_pallas_synth_scalars_device(self.h, seed, start, n, self.fmt, _addr(d_out)), "pallas synth_scalars")

    def synth_points_device(self, seed, start, n, d_out):
        self.loader._check(self.lib.snarkv_pallas_synth_points_device(self.h, seed, start, n, self.fmt, _addr(d_out)), "pallas synth_points")

    def field_op(self, field, op, a, b, n):
        out = ctypes.create_string_buffer(32 * n)
        self.loader._check(self.lib.snarkv_pallas_debug_field_op(self.h, field, op, _addr(a), _addr(b), n, out), "pallas field_op")
        return out.raw


@dataclass
class IpaAccumulator:
    """pcs/ipa/accumulator.rs: the challenges xi (k scalars) and the claimed commitment u"""
    xi: List[bytes]
    u: bytes


class IpaDecidingKey:
    """pcs/ipa/decider.rs:3-22: the committing key g (2^k Pallas points)"""

    def __init__(self, g: bytes):
        self.g = bytes(g)
        self.n = len(self.g) // 64
        assert self.n > 0 and self.n & (self.n - 1) == 0 and len(self.g) == 64 * self.n
        self.k = self.n.bit_length() - 1


class IpaAs:
    """`IpaAs<pallas::Affine, MOS>` restricted to AccumulationDecider::{decide, decide_all} (pcs/ipa/decider.rs:47-70)."""

    ASSERTION = "U == commit(G, h)"   # decider.rs:57

    def __init__(self, loader: CudaLoader, dk: IpaDecidingKey, flags=1):
        assert loader.fmt == CANONICAL, "accumulators come off a transcript: canonical bytes"
        self.loader, self.dk = loader, dk
        loader._check(loader.lib.snarkv_ipa_set_deciding_key(loader.h, dk.g, dk.n, loader.fmt, flags), "IpaDecidingKey")

    def decide_batch(self, accumulators: Sequence[IpaAccumulator]) -> bytes:
        accs = list(accumulators)
        if not accs:
            return b""
        for a in accs:
            assert len(a.xi) == self.dk.k
        L = self.loader
        acc = ctypes.create_string_buffer(len(accs))
        u = b"".join(a.u for a in accs)
        xi = b"".join(b"".join(a.xi) for a in accs)
        L._check(L.lib.snarkv_ipa_decide_batch(L.h, u, xi, self.dk.k, len(accs), L.fmt, acc), "IpaAs::decide")
        return acc.raw

    def decide(self, accumulator: IpaAccumulator):
        if self.decide_batch([accumulator]) != b"\x01":
            raise AssertionFailure(self.ASSERTION)

    def decide_all(self, accumulators: Sequence[IpaAccumulator]):
        if any(b != 1 for b in self.decide_batch(accumulators)):       # the reference aborts at the first failure (try_collect)
            raise AssertionFailure(self.ASSERTION)
