"""oracle -- TEST INFRASTRUCTURE ONLY: loaders for the CPU checkers.

  liboracle.so          our plain-C double restatement (oracle/ssim_oracle.c)
  _ref/libref_f32.so    the unmodified reference, float build  (CPU baseline "B")
  _ref/libref_f64.so    the unmodified reference, double build (gating oracle: "O1" AUTO, "O2" GENERIC)
  _ref/libnaive.so      the reference's own test oracle tests/ssim_naive.h, <double, uint8_t> and <double, uint16_t>

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

from ssim_b200._abi import Params, ThreadPool, bind_reference_api, make_params

HERE = os.path.dirname(os.path.abspath(__file__))
TAPS_RUNTIME, TAPS_TABLE = 0, 1
IMPL_AUTO, IMPL_GENERIC, IMPL_SSE, IMPL_SSE2, IMPL_AVX, IMPL_FMA, IMPL_AVX512 = range(7)

_cache = {}


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, oracle/_ref/."""
    subprocess.run(["make", "-C", HERE, "-j8"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    if path not in _cache:
        _cache[path] = C.CDLL(path, mode=os.RTLD_LOCAL)
    return _cache[path]


def have_ref():
    return all(os.path.exists(os.path.join(HERE, "_ref", n)) for n in ("libref_f32.so", "libref_f64.so"))


def have_naive():
    return os.path.exists(os.path.join(HERE, "_ref", "libnaive.so"))


def naive_ssim(a, b, want_map=False):
    """The reference's naive::compute_ssim<double, T> (tests/ssim_naive.h:230-339) on contiguous uint8 or uint16 arrays:
    T follows the dtype, so uint16 input means L = 65535.  Returns (double mean, float64 map or None)."""
    assert a.dtype == b.dtype and a.dtype in (np.uint8, np.uint16) and a.flags.c_contiguous and b.flags.c_contiguous
    lib = _load(os.path.join(HERE, "_ref", "libnaive.so"))
    fn = lib.naive_ssim_u16 if a.dtype == np.uint16 else lib.naive_ssim_u8
    fn.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_ssize_t, C.c_ssize_t, C.c_void_p, C.c_ssize_t, C.c_ssize_t,
                   C.c_void_p, C.c_ssize_t, C.c_ssize_t]
    fn.restype = C.c_double
    h, w = a.shape
    m = np.empty((h, w), dtype=np.float64) if want_map else None
    s = fn(w, h, a.ctypes.data, 1, w, b.ctypes.data, 1, w, m.ctypes.data if want_map else None, 1, w)
    return float(s), m


def oracle_lib():
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = _load(path)
    lib.ssim_oracle_compute.argtypes = [C.c_uint32, C.c_uint32,
                                        C.c_void_p, C.c_ssize_t, C.c_ssize_t,
                                        C.c_void_p, C.c_ssize_t, C.c_ssize_t,
                                        C.c_void_p, C.c_ssize_t, C.c_ssize_t,
                                        C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double)]
    lib.ssim_oracle_compute.restype = C.c_int
    lib.ssim_oracle_compute_u16.argtypes = lib.ssim_oracle_compute.argtypes
    lib.ssim_oracle_compute_u16.restype = C.c_int
    lib.ssim_oracle_taps.argtypes = [C.POINTER(C.c_double), C.c_int]
    lib.ssim_oracle_taps.restype = None
    lib.ssim_oracle_num_threads.restype = C.c_int
    return lib


def ref_lib(precision):
    """precision: 'f32' or 'f64'."""
    lib = _load(os.path.join(HERE, "_ref", "libref_%s.so" % precision))
    bind_reference_api(lib)
    lib.ref_select_impl.argtypes = [C.c_int]
    lib.ref_select_impl.restype = C.c_uint
    lib.ref_uses_double.restype = C.c_int
    return lib


def oracle_taps(mode=TAPS_RUNTIME):
    t = (C.c_double * 121)()
    oracle_lib().ssim_oracle_taps(t, mode)
    return np.array(t, dtype=np.float64).reshape(11, 11)


def oracle_ssim(a, b, want_map=False, taps=TAPS_TABLE, step_a=1, step_b=1, stride_a=None, stride_b=None,
                width=None, height=None, a_off=0, b_off=0):
    """Run the restatement on uint8 arrays.  Returns (ssim_float32, sum_double, map or None)."""
    if width is None:
        height, width = a.shape[:2]
    stride_a = stride_a if stride_a is not None else width * step_a
    stride_b = stride_b if stride_b is not None else width * step_b
    m = np.empty((height, width), dtype=np.float32) if want_map else None
    s = C.c_float()
    d = C.c_double()
    rc = oracle_lib().ssim_oracle_compute(width, height, a.ctypes.data + a_off, step_a, stride_a,
                                          b.ctypes.data + b_off, step_b, stride_b,
                                          m.ctypes.data if want_map else None, 1, width,
                                          taps, C.byref(s), C.byref(d))
    if rc != 0:
        raise OSError(rc, os.strerror(rc))
    return np.float32(s.value), d.value, m


def oracle_ssim_u16(a, b, want_map=False, taps=TAPS_TABLE):
    """16-bit restatement (L = 65535) on contiguous uint16 arrays.  Returns (ssim_float32, sum_double, map or None)."""
    assert a.dtype == np.uint16 and b.dtype == np.uint16 and a.flags.c_contiguous and b.flags.c_contiguous
    height, width = a.shape
    m = np.empty((height, width), dtype=np.float32) if want_map else None
    s = C.c_float()
    d = C.c_double()
    rc = oracle_lib().ssim_oracle_compute_u16(width, height, a.ctypes.data, 1, width, b.ctypes.data, 1, width,
                                              m.ctypes.data if want_map else None, 1, width, taps, C.byref(s), C.byref(d))
    if rc != 0:
        raise OSError(rc, os.strerror(rc))
    return np.float32(s.value), d.value, m


def ref_ssim(precision, a, b, want_map=False, impl=IMPL_AUTO, openmp=False, step_a=1, step_b=1,
             stride_a=None, stride_b=None, width=None, height=None, a_off=0, b_off=0, heap=False):
    """Run the unmodified reference build.  Returns (ssim_float32, map or None)."""
    lib = ref_lib(precision)
    lib.ref_select_impl(impl)
    if width is None:
        height, width = a.shape[:2]
    m = np.empty((height, width), dtype=np.float32) if want_map else None
    p = make_params(a, b, width, height, step_a, stride_a, step_b, stride_b, m, a_off=a_off, b_off=b_off)
    if heap:
        lib.rmgr_ssim_use_default_allocator(C.byref(p))
    s = C.c_float()
    if openmp:
        rc = lib.rmgr_ssim_compute_ssim_openmp(C.byref(s), C.byref(p))
    else:
        rc = lib.rmgr_ssim_compute_ssim(C.byref(s), C.byref(p), None)
    lib.ref_select_impl(IMPL_AUTO)
    if rc != 0:
        raise OSError(rc, os.strerror(rc))
    return np.float32(s.value), m
