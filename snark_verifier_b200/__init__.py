"""snark_verifier_b200 — B200-native BN254 G1 MSM + KZG accumulator decision behind snark-verifier's Loader / Decider surface.

The product is `libsnarkv_cuda.so` (hand-written sm_100a kernels, C ABI in include/snarkv_cuda.h).  This package is the
thin Python host mirror used by tests/ and bench.py; the C++ mirror of the reference traits lives in host/cuda_loader.hpp and
the Rust binding a maintainer would add is in INTEGRATION.md.

Names follow the reference (paths relative to /root/reference/snark-verifier/src):
  CudaLoader.multi_scalar_multiplication   <-> EcPointLoader::multi_scalar_multiplication   loader.rs:108-113, loader/native.rs:61-71
  Msm                                      <-> util::msm::Msm                                util/msm.rs:20-128
  KzgDecidingKey, KzgAs.decide/decide_all  <-> pcs/kzg/decider.rs:6-42, :70-93
  KzgAs.verify                             <-> pcs/kzg/accumulation.rs:41-63
  Error / AssertionFailure                 <-> lib.rs:18-28

There is no CPU fallback: importing works anywhere, but constructing a CudaLoader without the built library or without a
B200 raises immediately.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SNARKV_LIB_VARIANT=<name> (developer knob, tools/ A/B probes): load libsnarkv_cuda_<name>.so built by `make variant V=<name> DEFS=...`
LIB_PATH = os.path.join(_HERE, "libsnarkv_cuda%s.so" % ("_" + os.environ["SNARKV_LIB_VARIANT"] if os.environ.get("SNARKV_LIB_VARIANT") else ""))

CANONICAL, MONTGOMERY = 0, 1
CHECK_INPUTS = 1

OK, ERR_USAGE, ERR_EMPTY, ERR_CUDA, ERR_BAD_SCALAR, ERR_BAD_POINT, ERR_NO_KEY = 0, -1, -2, -3, -4, -5, -6

R_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


class Error(Exception):
    """Mirror of snark_verifier::Error (lib.rs:18-28)."""


class AssertionFailure(Error):
    """Error::AssertionFailure(String)"""


class CudaError(RuntimeError):
    """CUDA extension missing / no device / runtime failure.  Never swallowed, never replaced by a CPU path."""


class _FrInstr(ctypes.Structure):
    _fields_ = [("op", ctypes.c_uint32), ("dst", ctypes.c_uint32), ("a", ctypes.c_uint32), ("b", ctypes.c_uint32)]


class _PlonkPlanDesc(ctypes.Structure):
    """snarkv_plonk_plan_desc (include/snarkv_cuda.h)"""
    _fields_ = [("transcript", ctypes.c_uint32), ("stream_words", ctypes.c_uint32), ("seg_end", ctypes.c_void_p), ("n_challenges", ctypes.c_uint32),
                ("program", ctypes.c_void_p), ("n_instr", ctypes.c_size_t), ("n_regs", ctypes.c_uint32), ("consts", ctypes.c_void_p),
                ("n_consts", ctypes.c_size_t), ("n_inputs", ctypes.c_uint32), ("out_regs", ctypes.c_void_p), ("n_out", ctypes.c_uint32),
                ("row_src", ctypes.c_void_p), ("row_check", ctypes.c_void_p), ("n_lhs", ctypes.c_uint32), ("n_rhs", ctypes.c_uint32),
                ("lhs_src", ctypes.c_void_p), ("rhs_src", ctypes.c_void_p), ("const_points", ctypes.c_void_p), ("n_const_points", ctypes.c_uint32),
                ("n_pre", ctypes.c_uint32), ("n_items", ctypes.c_uint32), ("item_off", ctypes.c_void_p), ("item_pt", ctypes.c_void_p),
                ("n_points", ctypes.c_uint32)]


class _StageTime(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("ms", ctypes.c_float), ("launches", ctypes.c_int)]


_lib = None

# every symbol include/snarkv_cuda.h declares: (restype, argtypes)
_vp, _sz, _i, _u64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
C_ABI = {
    "snarkv_init": (_i, [_i, ctypes.POINTER(_vp)]),
    "snarkv_destroy": (None, [_vp]),
    "snarkv_last_error": (ctypes.c_char_p, [_vp]),
    "snarkv_version": (ctypes.c_char_p, []),
    "snarkv_set_stream": (_i, [_vp, _vp]),
    "snarkv_set_window_bits": (_i, [_vp, _i]),
    "snarkv_set_pairing_mode": (_i, [_vp, _i]),
    "snarkv_set_glv_mode": (_i, [_vp, _i]),
    "snarkv_set_accumulate_mode": (_i, [_vp, _i]),
    "snarkv_g1_msm": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp]),
    "snarkv_g1_msm_partial": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp]),
    "snarkv_g1_bases_upload": (_i, [_vp, _vp, _sz, _i, _i, ctypes.POINTER(_vp)]),
    "snarkv_g1_bases_free": (None, [_vp, _vp]),
    "snarkv_g1_msm_bases_resident": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp]),
    "snarkv_multi_init": (_i, [ctypes.POINTER(ctypes.c_int), _i, ctypes.POINTER(_vp)]),
    "snarkv_multi_destroy": (None, [_vp]),
    "snarkv_multi_device_count": (_i, [_vp]),
    "snarkv_multi_ctx": (_vp, [_vp, _i]),
    "snarkv_multi_last_error": (ctypes.c_char_p, [_vp]),
    "snarkv_multi_g1_msm": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp]),
    "snarkv_multi_g1_msm_batch_rlc": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _vp]),
    "snarkv_multi_kzg_set_deciding_key": (_i, [_vp, _vp, _vp, _vp]),
    "snarkv_multi_kzg_decide_batch": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "snarkv_g1_msm_device": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _vp]),
    "snarkv_g1_fold_partials_device": (_i, [_vp, _vp, _sz, _i, _vp]),
    "snarkv_g1_msm_batch": (_i, [_vp, _vp, _vp, _vp, _sz, _i, _i, _vp]),
    "snarkv_g1_msm_batch_rlc": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _i, _i, _vp]),
    "snarkv_fr_powers": (_i, [_vp, _vp, _sz, _i, _vp]),
    "snarkv_fr_batch_invert": (_i, [_vp, _vp, _sz, _vp, _i]),
    "snarkv_fr_mul_vec": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "snarkv_evm_transcript_challenges": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _i, _vp]),
    "snarkv_plonk_plan_create": (_i, [_vp, ctypes.POINTER(_PlonkPlanDesc), ctypes.POINTER(_vp)]),
    "snarkv_plonk_plan_free": (None, [_vp, _vp]),
    "snarkv_plonk_accumulate_batch": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, _vp]),
    "snarkv_poseidon_transcript_challenges": (_i, [_vp, _vp, _sz, _vp, _sz, _sz, _i, _vp]),
    "snarkv_g1_decompress_batch": (_i, [_vp, _vp, _sz, _i, _vp, _vp, _vp]),
    "snarkv_poseidon_permute": (_i, [_vp, _vp, _sz, _i, _vp]),
    "snarkv_kzg_accumulators_from_limbs": (_i, [_vp, _vp, _sz, ctypes.c_uint32, ctypes.c_uint32, _i, _vp, _vp, _vp]),
    "snarkv_fr_program_eval_batch": (_i, [_vp, _vp, _sz, ctypes.c_uint32, _vp, _sz, _vp, _sz, _sz, _vp, _sz, _i, _vp]),
    "snarkv_fr_program_eval_batch_device": (_i, [_vp, _vp, _sz, ctypes.c_uint32, _vp, _sz, _vp, _sz, _sz, _vp, _sz, _i, _vp]),
    "snarkv_kzg_accumulate": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp]),
    "snarkv_kzg_set_deciding_key": (_i, [_vp, _vp, _vp, _vp]),
    "snarkv_kzg_decide_batch": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "snarkv_kzg_decide_all_fused": (_i, [_vp, _vp, _vp, _sz, _vp, _i, _vp, _vp, _vp]),
    "snarkv_kzg_decide_batch_device": (_i, [_vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "snarkv_pallas_msm": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp]),
    "snarkv_pallas_msm_device": (_i, [_vp, _vp, _vp, _sz, _i, _i, _vp, _vp, _vp]),
    "snarkv_pallas_h_coeffs": (_i, [_vp, _vp, _sz, _vp, _i, _vp]),
    "snarkv_ipa_set_deciding_key": (_i, [_vp, _vp, _sz, _i, _i]),
    "snarkv_ipa_decide_batch": (_i, [_vp, _vp, _vp, _sz, _sz, _i, _vp]),
    "snarkv_pallas_synth_scalars_device": (_i, [_vp, _u64, _u64, _sz, _i, _vp]),
    "snarkv_pallas_synth_points_device": (_i, [_vp, _u64, _u64, _sz, _i, _vp]),
    "snarkv_pallas_debug_field_op": (_i, [_vp, _i, _i, _vp, _vp, _sz, _vp]),
    "snarkv_synth_scalars_device": (_i, [_vp, _u64, _u64, _sz, _i, _vp]),
    "snarkv_synth_points_device": (_i, [_vp, _u64, _u64, _sz, _i, _vp]),
    "snarkv_debug_field_op": (_i, [_vp, _i, _i, _vp, _vp, _sz, _vp]),
    "snarkv_g1_msm_plan": (_i, [_vp, _sz, ctypes.POINTER(ctypes.c_uint32)]),
    "snarkv_profile_enable": (_i, [_vp, _i]),
    "snarkv_profile_read": (_i, [_vp, ctypes.POINTER(_StageTime), _i]),
    "snarkv_launch_count": (_u64, [_vp]),
}


def load_library():
    """dlopen libsnarkv_cuda.so and bind every symbol of include/snarkv_cuda.h.  Raises CudaError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # not a fallback: build the sm_100a library in-tree if the toolchain is here, otherwise fail loudly
            import shutil
            import subprocess
            if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
                subprocess.call(["make", "-C", os.path.join(_HERE, "csrc"), "-j4", "-s"])
        if not os.path.exists(LIB_PATH):
            raise CudaError(f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "there is no CPU fallback")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in C_ABI.items():
            fn = getattr(lib, name)  # AttributeError here = header/library drift
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _addr(buf):
    """bytes-like / numpy array / int (device pointer) / None -> c_void_p"""
    if buf is None:
        return None
    if isinstance(buf, int):
        return ctypes.c_void_p(buf)
    if isinstance(buf, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(buf)) if isinstance(buf, bytearray) else ctypes.c_char_p(buf), ctypes.c_void_p)
    if hasattr(buf, "ctypes"):  # numpy
        return ctypes.c_void_p(buf.ctypes.data)
    if isinstance(buf, ctypes.Array):
        return ctypes.cast(buf, ctypes.c_void_p)
    raise TypeError(type(buf))


class CudaLoader:
    """The fourth interpretation of the verifier program (after Native / Evm / Halo2 loaders): values are computed on a B200.

    LoadedScalar = Fr and LoadedEcPoint = G1Affine as plain byte strings (32 B / 64 B) exactly like NativeLoader keeps
    them as plain host values (loader/native.rs:44,75); only the two hot operations are overridden.
    """

    def __init__(self, device=0, fmt=CANONICAL):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.snarkv_init(device, ctypes.byref(h))
        if rc != 0 or not h:
            raise CudaError(f"snarkv_init(device={device}) failed (rc={rc}): no sm_100 GPU visible; there is no CPU fallback")
        self.h = h
        self.device = device
        self.fmt = fmt

    # -- plumbing -------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.snarkv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc == 0:
            return
        msg = self.lib.snarkv_last_error(self.h).decode()
        if rc == ERR_CUDA:
            raise CudaError(f"{what}: {msg}")
        raise Error(f"{what}: rc={rc} {msg}")

    def set_stream(self, cuda_stream):
        self._check(self.lib.snarkv_set_stream(self.h, ctypes.c_void_p(cuda_stream) if cuda_stream else None), "set_stream")

    def set_glv_mode(self, mode):
        self._check(self.lib.snarkv_set_glv_mode(self.h, mode), "set_glv_mode")

    def set_accumulate_mode(self, mode):
        self._check(self.lib.snarkv_set_accumulate_mode(self.h, mode), "set_accumulate_mode")

    def set_pairing_mode(self, mode):
        self._check(self.lib.snarkv_set_pairing_mode(self.h, mode), "set_pairing_mode")

    def set_window_bits(self, c):
        self._check(self.lib.snarkv_set_window_bits(self.h, c), "set_window_bits")

    def msm_plan(self, n):
        out = (ctypes.c_uint32 * 4)()
        self._check(self.lib.snarkv_g1_msm_plan(self.h, n, out), "msm_plan")
        return {"window_bits": out[0], "windows": out[1], "buckets_per_window": out[2], "task_len": out[3]}

    def profile(self, on=True):
        self._check(self.lib.snarkv_profile_enable(self.h, 1 if on else 0), "profile_enable")

    def stage_times(self):
        arr = (_StageTime * 32)()
        k = self.lib.snarkv_profile_read(self.h, arr, 32)
        if k < 0:
            self._check(k, "profile_read")
        return [(arr[i].name.decode(), arr[i].ms, arr[i].launches) for i in range(k)]

    @property
    def launch_count(self):
        return int(self.lib.snarkv_launch_count(self.h))

    # -- EcPointLoader --------------------------------------------------------------------------------------------
    def ec_point_load_const(self, value):
        return bytes(value)

    def ec_point_assert_eq(self, annotation, lhs, rhs):
        if bytes(lhs) != bytes(rhs):
            raise AssertionFailure(annotation)  # loader/native.rs:50-59

    def multi_scalar_multiplication(self, pairs):
        """pairs: sequence of (scalar 32 B, point 64 B) — `&[(&LoadedScalar, &LoadedEcPoint)]` (loader.rs:108-113)."""
        pairs = list(pairs)
        scalars = b"".join(bytes(s) for s, _ in pairs)
        points = b"".join(bytes(p) for _, p in pairs)
        return self.msm(scalars, points, len(pairs))

    def msm(self, scalars, points, n, flags=0):
        """Contiguous form: n x 32 B scalars, n x 64 B points (bytes or numpy uint8) -> 64 B affine."""
        out = ctypes.create_string_buffer(64)
        self._check(self.lib.snarkv_g1_msm(self.h, _addr(scalars), _addr(points), n, self.fmt, flags, out), "multi_scalar_multiplication")
        return out.raw

    def bases_upload(self, points, n, flags=CHECK_INPUTS):
        """Upload a fixed base set once (validated by default); returns an opaque handle for msm_bases_resident / bases_free."""
        h = ctypes.c_void_p()
        self._check(self.lib.snarkv_g1_bases_upload(self.h, _addr(points), n, self.fmt, flags, ctypes.byref(h)), "bases_upload")
        return h

    def bases_free(self, handle):
        self.lib.snarkv_g1_bases_free(self.h, handle)

    def msm_bases_resident(self, handle, scalars, n, flags=0):
        """MSM against a resident base set: only the n x 32 B scalars cross PCIe."""
        out = ctypes.create_string_buffer(64)
        self._check(self.lib.snarkv_g1_msm_bases_resident(self.h, handle, _addr(scalars), n, self.fmt, flags, out), "msm_bases_resident")
        return out.raw

    def msm_partial(self, scalars, points, n, d_out_jacobian, flags=0):
        """Host slices in, 96-byte Jacobian partial left on the device (one rank's chunk of util/msm.rs:322-336)."""
        self._check(self.lib.snarkv_g1_msm_partial(self.h, _addr(scalars), _addr(points), n, self.fmt, flags, _addr(d_out_jacobian)),
                    "msm_partial")

    def msm_device(self, d_scalars, d_points, n, d_out_affine=None, d_out_jacobian=None, d_status=None, flags=0):
        self._check(self.lib.snarkv_g1_msm_device(self.h, _addr(d_scalars), _addr(d_points), n, self.fmt, flags,
                                                  _addr(d_out_affine), _addr(d_out_jacobian), _addr(d_status)), "msm_device")

    def fold_partials_device(self, d_partials, k, d_out_affine):
        self._check(self.lib.snarkv_g1_fold_partials_device(self.h, _addr(d_partials), k, self.fmt, _addr(d_out_affine)), "fold_partials")

    def msm_batch(self, scalars, points, offsets, flags=0):
        """m independent MSMs; offsets = m+1 cumulative term counts."""
        m = len(offsets) - 1
        off = (ctypes.c_uint64 * (m + 1))(*offsets)
        out = ctypes.create_string_buffer(64 * m)
        self._check(self.lib.snarkv_g1_msm_batch(self.h, _addr(scalars), _addr(points), ctypes.cast(off, ctypes.c_void_p), m,
                                                 self.fmt, flags, out), "msm_batch")
        return [out.raw[64 * j:64 * j + 64] for j in range(m)]

    def field_op(self, field, op, a, b, n):
        out = ctypes.create_string_buffer(32 * n)
        self._check(self.lib.snarkv_debug_field_op(self.h, field, op, _addr(a), _addr(b), n, out), "field_op")
        return out.raw

    def msm_batch_rlc(self, scalars, points, offsets, rho, flags=0):
        """sum_j rho^j * MSM_j as one MSM (scalars scaled by rho^j on the device)."""
        m = len(offsets) - 1
        off = offsets if hasattr(offsets, "ctypes") else (ctypes.c_uint64 * (m + 1))(*offsets)   # numpy uint64 array or sequence
        out = ctypes.create_string_buffer(64)
        self._check(self.lib.snarkv_g1_msm_batch_rlc(self.h, _addr(scalars), _addr(points), _addr(off), m, bytes(rho),
                                                     self.fmt, flags, out), "msm_batch_rlc")
        return out.raw

    # -- ScalarLoader helpers (Fr vectors) --------------------------------------------------------------------------------
    def powers(self, r, n):
        """LoadedScalar::powers (loader.rs:71-78): [1, r, ..., r^(n-1)] as n x 32 B."""
        out = ctypes.create_string_buffer(32 * n)
        self._check(self.lib.snarkv_fr_powers(self.h, bytes(r), n, self.fmt, out), "powers")
        return out.raw

    def batch_invert(self, values, n, coeff=None):
        """ScalarLoader::batch_invert / batch_invert_and_mul: non-zero v -> coeff / v, zeros untouched.  Returns new bytes."""
        buf = ctypes.create_string_buffer(bytes(values), 32 * n)
        self._check(self.lib.snarkv_fr_batch_invert(self.h, buf, n, bytes(coeff) if coeff is not None else None, self.fmt), "batch_invert")
        return buf.raw

    def fr_mul_vec(self, a, b, n):
        out = ctypes.create_string_buffer(32 * n)
        self._check(self.lib.snarkv_fr_mul_vec(self.h, _addr(a), _addr(b), n, self.fmt, out), "fr_mul_vec")
        return out.raw

    def evm_transcript_challenges(self, streams, stream_len, seg_end, m):
        """Keccak EvmTranscript challenges (transcript/evm.rs:184-222) for m proofs sharing one transcript shape -> m*k*32 bytes."""
        k = len(seg_end)
        se = (ctypes.c_uint32 * k)(*seg_end)
        out = ctypes.create_string_buffer(32 * k * m)
        self._check(self.lib.snarkv_evm_transcript_challenges(self.h, _addr(streams) if stream_len else None, stream_len,
                                                              ctypes.cast(se, ctypes.c_void_p), k, m, self.fmt, out), "evm_transcript")
        return out.raw

    def plonk_plan_create(self, stream_words, seg_end, program, row_src, row_check, lhs_src, rhs_src, const_points, poseidon=None):
        """snarkv_plonk_plan_create: the device-resident batch pipeline for one protocol (see include/snarkv_cuda.h) -> opaque plan"""
        import numpy as np
        from .plonk_eval import pack_program
        assert self.fmt == CANONICAL
        ins, consts, outs = pack_program(program)
        keep = [np.ascontiguousarray(ins), np.asarray(seg_end, dtype=np.uint32), np.asarray(outs, dtype=np.uint32),
                np.asarray(row_src, dtype=np.int32), np.asarray(row_check, dtype=np.uint8), np.asarray(lhs_src, dtype=np.int32),
                np.asarray(rhs_src, dtype=np.int32), np.frombuffer(b"".join(const_points) or bytes(64), dtype=np.uint8),
                np.frombuffer(consts or bytes(32), dtype=np.uint8)]
        d = _PlonkPlanDesc(0, stream_words, keep[1].ctypes.data, len(seg_end), keep[0].ctypes.data, ins.shape[0], program.n_regs,
                           keep[8].ctypes.data, len(program.consts), program.n_inputs, keep[2].ctypes.data, len(program.outputs),
                           keep[3].ctypes.data, keep[4].ctypes.data, len(lhs_src), len(rhs_src), keep[5].ctypes.data, keep[6].ctypes.data,
                           keep[7].ctypes.data, len(const_points), 0, 0, None, None, 0)
        if poseidon is not None:                                   # (n_pre, item_off, item_pt, n_points): the Poseidon transcript's item table
            n_pre, item_off, item_pt, n_points = poseidon
            keep += [np.asarray(item_off, dtype=np.int32), np.asarray(item_pt, dtype=np.int32)]
            d.transcript, d.n_pre, d.n_items, d.n_points = 1, n_pre, len(item_off), n_points
            d.item_off, d.item_pt = keep[-2].ctypes.data, keep[-1].ctypes.data
        plan = ctypes.c_void_p()
        self._check(self.lib.snarkv_plonk_plan_create(self.h, ctypes.byref(d), ctypes.byref(plan)), "plonk_plan_create")
        return plan

    def plonk_plan_free(self, plan):
        self.lib.snarkv_plonk_plan_free(self.h, plan)

    def plonk_accumulate_batch(self, plan, streams, m, rho, decide=False):
        """-> (lhs 64 B, rhs 64 B, accept | None)"""
        lhs, rhs, acc = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64), ctypes.create_string_buffer(1)
        self._check(self.lib.snarkv_plonk_accumulate_batch(self.h, plan, _addr(streams), m, bytes(rho), 1 if decide else 0, lhs, rhs, acc),
                    "plonk_accumulate_batch")
        return lhs.raw, rhs.raw, (acc.raw == b"\x01" if decide else None)

    def poseidon_transcript_challenges(self, elements, stream_len, seg_end, m):
        """Poseidon transcript challenges (util/hash/poseidon.rs:117-203, transcript/halo2.rs:201-242) for m proofs sharing one
        transcript shape: `elements` = m x stream_len x 32 B absorbed scalar-field elements, `seg_end` = element offsets -> m*k*32 bytes."""
        k = len(seg_end)
        se = (ctypes.c_uint32 * k)(*seg_end)
        out = ctypes.create_string_buffer(32 * k * m)
        self._check(self.lib.snarkv_poseidon_transcript_challenges(self.h, _addr(elements) if stream_len else None, stream_len,
                                                                   ctypes.cast(se, ctypes.c_void_p), k, m, self.fmt, out), "poseidon_transcript")
        return out.raw

    def g1_decompress(self, compressed, n, want_elements=True):
        """`C::from_bytes` for n compressed G1 points (transcript/halo2.rs:258-272) -> (points n*64 B, elements n*64 B | None, valid n bytes)"""
        pts, valid = ctypes.create_string_buffer(64 * n), ctypes.create_string_buffer(max(n, 1))
        el = ctypes.create_string_buffer(64 * n) if want_elements else None
        self._check(self.lib.snarkv_g1_decompress_batch(self.h, _addr(compressed), n, self.fmt, pts, el, valid), "g1_decompress")
        return pts.raw, (el.raw if want_elements else None), valid.raw[:n]

    def poseidon_permute(self, states, m):
        out = ctypes.create_string_buffer(32 * 5 * m)
        self._check(self.lib.snarkv_poseidon_permute(self.h, _addr(states), m, self.fmt, out), "poseidon_permute")
        return out.raw

    def accumulators_from_limbs(self, limbs, m, num_limbs=4, limb_bits=68):
        """LimbsEncoding::from_repr (pcs/kzg/accumulator.rs:57-81) for m accumulators -> (lhs m*64 B, rhs m*64 B, valid m bytes)."""
        lhs, rhs, valid = ctypes.create_string_buffer(64 * m), ctypes.create_string_buffer(64 * m), ctypes.create_string_buffer(max(m, 1))
        self._check(self.lib.snarkv_kzg_accumulators_from_limbs(self.h, _addr(limbs), m, num_limbs, limb_bits, self.fmt, lhs, rhs, valid),
                    "accumulators_from_limbs")
        return lhs.raw, rhs.raw, valid.raw[:m]

    def fr_program_eval(self, program, inputs, m, d_inputs=None, d_outputs=None):
        """Run a plonk_eval.Program for m proofs (protocol.rs:211-283, 333-392; proof.rs:298-349 for a batch).  `inputs`:
        m * n_inputs * 32 bytes -> m * n_out * 32 bytes; with d_inputs / d_outputs (device pointers) nothing crosses PCIe but
        the program."""
        from .plonk_eval import pack_program
        ins, consts, outs = pack_program(program)
        if self.fmt == MONTGOMERY:   # constants travel in the loader's format like every other scalar
            consts = b"".join(((c << 256) % R_MODULUS).to_bytes(32, "little") for c in program.consts)
        n_out = len(program.outputs)
        args = (ins.ctypes.data, ins.shape[0], program.n_regs, consts if consts else None, len(program.consts))
        if d_inputs is not None:
            self._check(self.lib.snarkv_fr_program_eval_batch_device(self.h, *args, _addr(d_inputs), program.n_inputs, m, outs.ctypes.data, n_out,
                                                                     self.fmt, _addr(d_outputs)), "fr_program_eval")
            return None
        out = ctypes.create_string_buffer(max(32 * n_out * m, 1))
        self._check(self.lib.snarkv_fr_program_eval_batch(self.h, *args, _addr(inputs) if program.n_inputs and m else None, program.n_inputs, m,
                                                          outs.ctypes.data, n_out, self.fmt, out), "fr_program_eval")
        return out.raw[: 32 * n_out * m]

    # -- synthetic workload ---------------------------------------------------------------------------------------
    def synth_scalars_device(self, seed, start, n, d_out):
        self._check(self.lib.snarkv_synth_scalars_device(self.h, seed, start, n, self.fmt, _addr(d_out)), "synth_scalars")

    def synth_points_device(self, seed, start, n, d_out):
        self._check(self.lib.snarkv_synth_points_device(self.h, seed, start, n, self.fmt, _addr(d_out)), "synth_points")


class MultiCudaLoader:
    """One host call, all the GPUs of the box (snarkv_multi_*): the chunk partition of util/msm.rs:322-336 with devices for rayon
    threads, folded on device 0 over NVLink peer memory; pairing checks and RLC batches shard as independent units."""

    def __init__(self, devices=None, fmt=CANONICAL):
        self.lib = load_library()
        h = ctypes.c_void_p()
        if devices is None:
            rc = self.lib.snarkv_multi_init(None, 0, ctypes.byref(h))
        else:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self.lib.snarkv_multi_init(arr, len(devices), ctypes.byref(h))
        if rc != 0 or not h:
            raise CudaError(f"snarkv_multi_init failed (rc={rc}): needs sm_100 GPUs with peer access; there is no CPU fallback")
        self.h, self.fmt = h, fmt
        self.n_devices = int(self.lib.snarkv_multi_device_count(h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.snarkv_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc == 0:
            return
        msg = self.lib.snarkv_multi_last_error(self.h).decode()
        if rc == ERR_CUDA:
            raise CudaError(f"{what}: {msg}")
        raise Error(f"{what}: rc={rc} {msg}")

    def device_ctx(self, i):
        """Per-device context handle (for the tuning setters of the C ABI)."""
        return ctypes.c_void_p(self.lib.snarkv_multi_ctx(self.h, i))

    def set_accumulate_mode(self, mode):
        for i in range(self.n_devices):
            self.lib.snarkv_set_accumulate_mode(self.device_ctx(i), mode)

    def multi_scalar_multiplication(self, pairs):
        pairs = list(pairs)
        return self.msm(b"".join(bytes(s) for s, _ in pairs), b"".join(bytes(p) for _, p in pairs), len(pairs))

    def msm(self, scalars, points, n, flags=0):
        out = ctypes.create_string_buffer(64)
        self._check(self.lib.snarkv_multi_g1_msm(self.h, _addr(scalars), _addr(points), n, self.fmt, flags, out), "multi msm")
        return out.raw

    def msm_batch_rlc(self, scalars, points, offsets, rho, flags=0):
        m = len(offsets) - 1
        off = offsets if hasattr(offsets, "ctypes") else (ctypes.c_uint64 * (m + 1))(*offsets)
        out = ctypes.create_string_buffer(64)
        self._check(self.lib.snarkv_multi_g1_msm_batch_rlc(self.h, _addr(scalars), _addr(points), _addr(off), m, bytes(rho), self.fmt, flags, out),
                    "multi msm_batch_rlc")
        return out.raw

    def set_deciding_key(self, dk):
        self._check(self.lib.snarkv_multi_kzg_set_deciding_key(self.h, dk.g1, dk.g2, dk.s_g2), "multi KzgDecidingKey")

    def decide_batch(self, lhs, rhs, n, want_gt=False):
        acc = ctypes.create_string_buffer(max(n, 1))
        gt = ctypes.create_string_buffer(384 * n) if want_gt else None
        self._check(self.lib.snarkv_multi_kzg_decide_batch(self.h, _addr(lhs), _addr(rhs), n, self.fmt, acc, gt), "multi decide")
        return acc.raw[:n], (gt.raw if want_gt else None)


class Msm:
    """Host mirror of `util::msm::Msm` (util/msm.rs:20-128): constant + scalars + bases, evaluated by the loader's MSM.
    Scalars are Python ints mod r here (the deferred algebra is microseconds of Fr work and stays on the host)."""

    def __init__(self, loader, constant=None, scalars=None, bases=None):
        self.loader, self.constant = loader, constant
        self.scalars, self.bases = list(scalars or []), list(bases or [])

    @classmethod
    def constant_(cls, loader, c):
        return cls(loader, constant=c % R_MODULUS)

    @classmethod
    def base(cls, loader, b):
        return cls(loader, scalars=[1], bases=[bytes(b)])  # util/msm.rs:54-61

    def scale(self, k):  # util/msm.rs:100-107
        if self.constant is not None:
            self.constant = self.constant * k % R_MODULUS
        self.scalars = [s * k % R_MODULUS for s in self.scalars]
        return self

    def push(self, scalar, base):  # util/msm.rs:109-116 (dedupe by equality)
        base = bytes(base)
        if base in self.bases:
            i = self.bases.index(base)
            self.scalars[i] = (self.scalars[i] + scalar) % R_MODULUS
        else:
            self.scalars.append(scalar % R_MODULUS)
            self.bases.append(base)

    def extend(self, other):  # util/msm.rs:118-128
        if other.constant is not None:
            self.constant = other.constant if self.constant is None else (self.constant + other.constant) % R_MODULUS
        for s, b in zip(other.scalars, other.bases):
            self.push(s, b)
        return self

    def __add__(self, other):
        r = Msm(self.loader, self.constant, self.scalars, self.bases)
        return r.extend(other)

    def __mul__(self, k):
        return Msm(self.loader, self.constant, self.scalars, self.bases).scale(k)

    def __neg__(self):  # util/msm.rs:192-204
        return Msm(self.loader, None if self.constant is None else (-self.constant) % R_MODULUS,
                   [(-s) % R_MODULUS for s in self.scalars], self.bases)

    def __sub__(self, other):  # util/msm.rs:156-178
        return self + (-other)

    def size(self):  # util/msm.rs:63-65
        return len(self.bases)

    @staticmethod
    def sum(loader, msms):  # util/msm.rs:218-226 (an empty sum is the default, constant-free Msm)
        acc = Msm(loader)
        for m in msms:
            acc = acc + m
        return acc

    def evaluate(self, gen=None):  # util/msm.rs:81-98
        pairs = []
        if self.constant is not None:
            if gen is None:
                raise ValueError("constant term needs a generator")  # the reference panics (unwrap on None)
            pairs.append((self.constant.to_bytes(32, "little"), bytes(gen)))
        pairs += [(s.to_bytes(32, "little"), b) for s, b in zip(self.scalars, self.bases)]
        assert self.loader.fmt == CANONICAL, "Msm host mirror keeps canonical scalars"
        return self.loader.multi_scalar_multiplication(pairs)


class KzgAccumulator:
    """pcs/kzg/accumulator.rs:6-26"""

    def __init__(self, lhs, rhs):
        self.lhs, self.rhs = bytes(lhs), bytes(rhs)


class KzgDecidingKey:
    """pcs/kzg/decider.rs:6-42: (svk.g, g2, s_g2).  Installing it on a loader precomputes both G2Prepared on the device."""

    def __init__(self, g1, g2, s_g2):
        self.g1, self.g2, self.s_g2 = bytes(g1), bytes(g2), bytes(s_g2)


class KzgAs:
    """`KzgAs<Bn256, MOS>` restricted to the hot path: AccumulationDecider::{decide, decide_all} and AccumulationScheme::verify."""

    ASSERTION = "e(lhs, g2)·e(rhs, -s_g2) == O"  # decider.rs:81

    def __init__(self, loader, dk):
        self.loader, self.dk = loader, dk
        L = loader
        L._check(L.lib.snarkv_kzg_set_deciding_key(L.h, dk.g1, dk.g2, dk.s_g2), "KzgDecidingKey")

    def decide_batch(self, lhs, rhs, n, want_gt=False):
        """n accumulators as contiguous n x 64 B arrays -> (accept bytes, gt bytes | None)."""
        L = self.loader
        acc = ctypes.create_string_buffer(n)
        gt = ctypes.create_string_buffer(384 * n) if want_gt else None
        L._check(L.lib.snarkv_kzg_decide_batch(L.h, _addr(lhs), _addr(rhs), n, L.fmt, acc, gt), "decide")
        return acc.raw, (gt.raw if want_gt else None)

    def decide_batch_device(self, d_lhs, d_rhs, n, d_accept, d_gt=None):
        L = self.loader
        L._check(L.lib.snarkv_kzg_decide_batch_device(L.h, _addr(d_lhs), _addr(d_rhs), n, L.fmt, _addr(d_accept), _addr(d_gt)), "decide_device")

    def decide(self, accumulator):  # decider.rs:70-82
        acc, _ = self.decide_batch(accumulator.lhs, accumulator.rhs, 1)
        if acc[0] != 1:
            raise AssertionFailure(self.ASSERTION)

    def decide_all(self, accumulators):  # decider.rs:84-93
        accumulators = list(accumulators)
        if not accumulators:
            return
        lhs = b"".join(a.lhs for a in accumulators)
        rhs = b"".join(a.rhs for a in accumulators)
        acc, _ = self.decide_batch(lhs, rhs, len(accumulators))
        if any(b != 1 for b in acc):
            raise AssertionFailure(self.ASSERTION)

    def decide_all_fused(self, lhs, rhs, n, rho):
        """RLC batching of decider.rs:146-185: one accumulate (two MSMs with powers of rho) + ONE pairing.  -> (accept, KzgAccumulator)"""
        L = self.loader
        acc = ctypes.create_string_buffer(1)
        ol, orr = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        L._check(L.lib.snarkv_kzg_decide_all_fused(L.h, _addr(lhs), _addr(rhs), n, bytes(rho), L.fmt, acc, ol, orr), "decide_all_fused")
        return acc.raw[0] == 1, KzgAccumulator(ol.raw, orr.raw)

    def verify(self, instances, r, blind=None):  # accumulation.rs:41-63
        accs = list(instances) + ([blind] if blind is not None else [])
        L = self.loader
        lhs = b"".join(a.lhs for a in accs)
        rhs = b"".join(a.rhs for a in accs)
        ol, orr = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        L._check(L.lib.snarkv_kzg_accumulate(L.h, lhs, rhs, len(accs), bytes(r), L.fmt, ol, orr), "KzgAs::verify")
        return KzgAccumulator(ol.raw, orr.raw)
