"""ctypes wrapper of the C oracle (oracle/c/bp_oracle.c -> oracle/liboracle_c.so).  Test infrastructure."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle_c.so")
    if force or not os.path.exists(so):
        r = subprocess.run(["make", "-C", os.path.join(_HERE, "c")], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError("building the C oracle failed:\n" + r.stdout + r.stderr)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, sz, i = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.orc_msm.argtypes = [i, vp, vp, sz, i, vp]
        L.orc_scalar_mul.argtypes = [i, vp, vp, vp]
        L.orc_binary_scalar_mul.argtypes = [i, vp, vp, vp, vp, vp]
        L.orc_multiples.argtypes = [i, vp, sz, vp]
        L.orc_fr_op.argtypes = [i, i, vp, vp, vp]
        L.orc_keccak_f1600.argtypes = [vp]
        L.orc_keccak_f1600.restype = None
        _LIB = L
    return _LIB


def _p(b):
    return ctypes.cast(ctypes.c_char_p(bytes(b)), ctypes.c_void_p)


def msm(curve_id, points_xy, scalars_be, n, threads=1):
    """Straus/wNAF-5 MSM (the reference's var-time algorithm); returns X||Y bytes."""
    mb = 48 if curve_id == 0 else 32
    out = ctypes.create_string_buffer(2 * mb)
    lib().orc_msm(curve_id, _p(points_xy), _p(scalars_be), n, threads, out)
    return out.raw


def scalar_mul(curve_id, pt_xy, scalar_be):
    mb = 48 if curve_id == 0 else 32
    out = ctypes.create_string_buffer(2 * mb)
    lib().orc_scalar_mul(curve_id, _p(pt_xy), _p(scalar_be), out)
    return out.raw


def binary_scalar_mul(curve_id, g_xy, h_xy, r1_be, r2_be):
    mb = 48 if curve_id == 0 else 32
    out = ctypes.create_string_buffer(2 * mb)
    lib().orc_binary_scalar_mul(curve_id, _p(g_xy), _p(h_xy), _p(r1_be), _p(r2_be), out)
    return out.raw


def multiples(curve_id, base_xy, n):
    """[(i+1)*B for i < n] as concatenated X||Y bytes (structured synthetic inputs)."""
    mb = 48 if curve_id == 0 else 32
    out = ctypes.create_string_buffer(max(1, n * 2 * mb))
    lib().orc_multiples(curve_id, _p(base_xy), n, out)
    return out.raw[:n * 2 * mb]


def fr_op(curve_id, op, a_be, b_be):
    mb = 48 if curve_id == 0 else 32
    out = ctypes.create_string_buffer(mb)
    lib().orc_fr_op(curve_id, op, _p(a_be), _p(b_be), out)
    return out.raw


def keccak_f1600(state):
    """in-place Keccak-f[1600] on a 200-byte bytearray (same contract as oracle.merlin.keccak_f1600)."""
    buf = (ctypes.c_uint8 * 200).from_buffer(state)
    lib().orc_keccak_f1600(ctypes.addressof(buf))
