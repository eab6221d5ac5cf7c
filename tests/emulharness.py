"""TEST INFRASTRUCTURE: ctypes front end of tests/emul/libemul.so (host replay of the per-thread
device functions in csrc/gr_math.cuh and csrc/fft_radix.cuh)."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(ROOT, "tests", "emul", "libemul.so")
        src = os.path.join(ROOT, "tests", "emul", "emul.cu")
        deps = [src] + [os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200", "csrc", f) for f in ("gr_math.cuh", "fft_radix.cuh")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call([os.path.join(ROOT, "tools", "gen_tables.py")])
            subprocess.check_call(["nvcc", "-O2", "-shared", "-Wno-deprecated-gpu-targets", "-Xcompiler",
                                   "-fPIC,-ffp-contract=off", "-I" + os.path.join(ROOT, "build", "generated"),
                                   "-o", so, src])
        _LIB = C.CDLL(so)
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def fft(x, radices, dirn):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    r = (C.c_int * len(radices))(*radices)
    assert lib().emul_fft(len(x), len(radices), r, int(dirn), _p(x), _p(out)) == 0
    return out


def fast_atan2f(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    o = np.zeros_like(y)
    lib().emul_fast_atan2f(_p(y), _p(x), _p(o), C.c_long(len(y)))
    return o


def quad_demod(gain, x):
    xin = np.concatenate([np.zeros(1, np.complex64), np.ascontiguousarray(x, np.complex64)])
    o = np.zeros(len(x), np.float32)
    lib().emul_quad_demod(C.c_float(gain), _p(xin), C.c_long(len(x)), _p(o))
    return o


def fir_fff(taps, x, order, nchan=1, chan=0):
    """x: new items of one stream; embeds it in a [time][nchan] matrix to exercise the stride."""
    t = np.ascontiguousarray(taps, np.float32)
    h = max(len(t) - 1, 0)
    col = np.concatenate([np.zeros(h, np.float32), np.ascontiguousarray(x, np.float32)])
    mat = np.zeros((len(col), nchan), np.float32)
    mat[:, chan] = col
    out = np.zeros((len(x), nchan), np.float32)
    lib().emul_fir_fff(_p(t), len(t), C.c_void_p(mat.ctypes.data + 4 * chan), C.c_long(nchan), C.c_long(len(x)),
                       C.c_void_p(out.ctypes.data + 4 * chan), int(order), C.c_long(-h))
    return out[:, chan].copy()


def mm(args, x, order, abs0=0, noutput=None):
    x = np.ascontiguousarray(x, np.float32)
    nout = len(x) if noutput is None else noutput
    out = np.zeros(max(nout, 1), np.float32)
    consumed = C.c_int(0)
    st = np.zeros(3, np.float32)
    a = [C.c_float(v) for v in args]
    r = lib().emul_mm(*a, _p(x), len(x), _p(out), int(nout), C.byref(consumed), int(order), C.c_long(abs0), _p(st))
    return out[:r], consumed.value, st


def slice4(alpha, x):
    x = np.ascontiguousarray(x, np.float32)
    o = np.zeros(len(x), np.uint8)
    lib().emul_slice4(C.c_float(alpha), _p(x), C.c_long(len(x)), _p(o))
    return o


def corr(code_bits, threshold, bits):
    n = len(code_bits)
    code = 0
    for i in range(64):
        code = (code << 1) | (code_bits[i] if i < n else 0)
    mask = ((1 << n) - 1) << (64 - n)
    flag = 1 << (64 - n)
    b = np.ascontiguousarray(bits, np.uint8)
    o = np.zeros(len(b), np.uint8)
    lib().emul_corr(C.c_ulonglong(code), C.c_ulonglong(mask), C.c_ulonglong(flag), C.c_uint(threshold), _p(b),
                    C.c_long(len(b)), _p(o))
    return o
