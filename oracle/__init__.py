"""CPU oracle package (TEST INFRASTRUCTURE ONLY — see oracle/bn254.hpp).

`oracle.lib()` returns a ctypes handle on liboracle_bn254.so, building it with `make` when it is missing.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_bn254.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO) for f in ("oracle.cpp", "bn254.hpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        u8p, sz, i32, u64 = ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint64
        vp = ctypes.c_void_p
        sigs = {
            "oracle_fp_mul": [i32, u8p, u8p, u8p],
            "oracle_fp_inv": [i32, u8p, u8p],
            "oracle_fp_to_mont": [i32, u8p, u8p],
            "oracle_g1_is_on_curve": [u8p],
            "oracle_g1_add": [u8p, u8p, u8p],
            "oracle_g1_mul": [u8p, u8p, u8p],
            "oracle_g2_generator": [u8p],
            "oracle_g2_mul": [u8p, u8p, u8p],
            "oracle_g2_is_on_curve": [u8p],
            "oracle_msm_native": [vp, vp, sz, u8p],
            "oracle_msm_pippenger": [vp, vp, sz, i32, u8p],
            "oracle_msm_pippenger_raw": [vp, vp, sz, i32, u8p],
            "oracle_to_mont_batch": [i32, vp, sz, i32, vp],
            "oracle_kzg_decide": [u8p, u8p, u8p, u8p, u8p, u8p],
            "oracle_kzg_decide_batch": [vp, vp, sz, u8p, u8p, i32, i32, vp, vp],
            "oracle_kzg_accumulate": [vp, vp, sz, u8p, u8p, u8p],
            "oracle_synth_scalars": [u64, u64, sz, vp],
            "oracle_synth_point_scalars": [u64, u64, sz, vp],
            "oracle_synth_points": [u64, u64, sz, i32, vp],
            "oracle_synth_points_slow": [u64, u64, sz, vp],
            "oracle_msm_expected_from_dlogs": [vp, vp, sz, u8p],
        }
        for name, args in sigs.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = None if name.startswith("oracle_synth") else i32
        L.oracle_mulmod_count_reset.restype = u64
        _lib = L
    return _lib


def _buf(n):
    return ctypes.create_string_buffer(n)


def _ptr(b):
    """bytes / bytearray / numpy array -> void* usable for the vp-typed arguments"""
    if b is None:
        return None
    if isinstance(b, (bytes, bytearray)):
        return ctypes.cast(ctypes.c_char_p(bytes(b)), ctypes.c_void_p)
    return ctypes.c_void_p(b.ctypes.data)  # numpy


def _chk(rc, what):
    if rc != 0:
        raise ValueError(f"oracle {what} failed: rc={rc}")


# ---- thin pythonic wrappers (bytes in / bytes out) -----------------------------------------------------------------
def fp_mul(field, a, b):
    o = _buf(32); _chk(lib().oracle_fp_mul(field, a, b, o), "fp_mul"); return o.raw
def fp_inv(field, a):
    o = _buf(32); _chk(lib().oracle_fp_inv(field, a, o), "fp_inv"); return o.raw
def fp_to_mont(field, a):
    o = _buf(32); _chk(lib().oracle_fp_to_mont(field, a, o), "fp_to_mont"); return o.raw
def g1_is_on_curve(p): return bool(lib().oracle_g1_is_on_curve(p))
def g1_add(a, b):
    o = _buf(64); _chk(lib().oracle_g1_add(a, b, o), "g1_add"); return o.raw
def g1_mul(p, s):
    o = _buf(64); _chk(lib().oracle_g1_mul(p, s, o), "g1_mul"); return o.raw
def g2_generator():
    o = _buf(128); lib().oracle_g2_generator(o); return o.raw
def g2_mul(q, s):
    o = _buf(128); _chk(lib().oracle_g2_mul(q, s, o), "g2_mul"); return o.raw
def msm_native(scalars, points, n):
    o = _buf(64); _chk(lib().oracle_msm_native(_ptr(scalars), _ptr(points), n, o), "msm_native"); return o.raw
def msm_pippenger(scalars, points, n, threads=1):
    o = _buf(64); _chk(lib().oracle_msm_pippenger(_ptr(scalars), _ptr(points), n, threads, o), "msm_pippenger"); return o.raw
def msm_pippenger_raw(scalars_mont, points_mont, n, threads=1):
    """halo2curves in-memory layout in (what the Rust reference holds), canonical affine bytes out; no parsing inside."""
    o = _buf(64); _chk(lib().oracle_msm_pippenger_raw(_ptr(scalars_mont), _ptr(points_mont), n, threads, o), "msm_pippenger_raw"); return o.raw
def to_mont_batch(field, data, n, threads=1):
    o = _buf(32 * n); _chk(lib().oracle_to_mont_batch(field, _ptr(data), n, threads, ctypes.cast(o, ctypes.c_void_p)), "to_mont_batch"); return o.raw
def kzg_decide(lhs, rhs, g2, s_g2, want_gt=True):
    acc = _buf(1); gt = _buf(384) if want_gt else None
    _chk(lib().oracle_kzg_decide(lhs, rhs, g2, s_g2, acc, gt), "kzg_decide")
    return acc.raw[0] == 1, (gt.raw if want_gt else None)
def kzg_decide_batch(lhs, rhs, n, g2, s_g2, threads=1, hoist=0, want_gt=False):
    acc = _buf(n); gt = _buf(384 * n) if want_gt else None
    _chk(lib().oracle_kzg_decide_batch(_ptr(lhs), _ptr(rhs), n, g2, s_g2, threads, hoist,
                                       ctypes.cast(acc, ctypes.c_void_p), ctypes.cast(gt, ctypes.c_void_p) if want_gt else None),
         "kzg_decide_batch")
    return acc.raw, (gt.raw if want_gt else None)
def kzg_accumulate(lhs, rhs, n, r):
    a = _buf(64); b = _buf(64)
    _chk(lib().oracle_kzg_accumulate(_ptr(lhs), _ptr(rhs), n, r, a, b), "kzg_accumulate"); return a.raw, b.raw
def synth_scalars(seed, start, n):
    o = _buf(32 * n); lib().oracle_synth_scalars(seed, start, n, ctypes.cast(o, ctypes.c_void_p)); return o.raw
def synth_point_scalars(seed, start, n):
    o = (ctypes.c_uint64 * n)(); lib().oracle_synth_point_scalars(seed, start, n, ctypes.cast(o, ctypes.c_void_p)); return o
def synth_points(seed, start, n, threads=1):
    o = _buf(64 * n); lib().oracle_synth_points(seed, start, n, threads, ctypes.cast(o, ctypes.c_void_p)); return o.raw
def synth_points_slow(seed, start, n):
    o = _buf(64 * n); lib().oracle_synth_points_slow(seed, start, n, ctypes.cast(o, ctypes.c_void_p)); return o.raw
def msm_expected_from_dlogs(scalars, t, n):
    o = _buf(64)
    _chk(lib().oracle_msm_expected_from_dlogs(_ptr(scalars), ctypes.cast(t, ctypes.c_void_p), n, o), "expected_from_dlogs"); return o.raw
def mulmod_count_reset(): return lib().oracle_mulmod_count_reset()
