"""ctypes binding of the product libraries (no torch needed):

  librmgr-ssim.so   the reference's C API  -- rmgr_ssim_compute_ssim() & friends (include/rmgr/ssim.h)
  libssim_cuda.so   the C-ABI CUDA engine  -- ssim_cuda_*()                      (include/ssim_cuda.h)
  libssim_imgio.so  the front end's JPEG reader -- ssim_imgio_decode_jpeg()      (include/ssim_imgio.h; host code)

There is no CPU fallback: importing works anywhere, but every compute call needs the built libraries and a
B200; a missing library raises immediately (RuntimeError) instead of degrading."""
import ctypes as C
import os

import numpy as np

from ._abi import Params, ThreadPool, Version, bind_reference_api, make_params

LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
_libs = {}


def _load(name):
    if name not in _libs:
        path = os.path.join(LIB_DIR, name)
        if not os.path.exists(path):
            raise RuntimeError("%s is not built (run `make` or __graft_entry__.build()); ssim_b200 has no CPU fallback" % path)
        _libs[name] = C.CDLL(path, mode=os.RTLD_LOCAL)
    return _libs[name]


def cuda_lib():
    """libssim_cuda.so with argtypes declared for every symbol of include/ssim_cuda.h."""
    lib = _load("libssim_cuda.so")
    if getattr(lib, "_bound", False):
        return lib
    u8p, f32p, f64p, vp = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.ssim_cuda_abi_version.restype = C.c_int
    lib.ssim_cuda_device_count.restype = C.c_int
    lib.ssim_cuda_init.argtypes = [C.c_int]
    lib.ssim_cuda_init.restype = C.c_int
    lib.ssim_cuda_shutdown.restype = None
    lib.ssim_cuda_last_error_string.restype = C.c_char_p
    lib.ssim_cuda_host_alloc.argtypes = [C.c_size_t]
    lib.ssim_cuda_host_alloc.restype = C.c_void_p
    lib.ssim_cuda_host_free.argtypes = [C.c_void_p]
    lib.ssim_cuda_host_free.restype = None
    lib.ssim_cuda_compute.argtypes = [C.c_int, C.c_uint32, C.c_uint32, u8p, C.c_ssize_t, C.c_ssize_t, u8p, C.c_ssize_t, C.c_ssize_t,
                                      f32p, C.c_ssize_t, C.c_ssize_t, C.POINTER(C.c_float)]
    lib.ssim_cuda_compute.restype = C.c_int
    lib.ssim_cuda_compute_luma.argtypes = lib.ssim_cuda_compute.argtypes
    lib.ssim_cuda_compute_luma.restype = C.c_int
    lib.ssim_cuda_compute_channels.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, u8p, C.c_ssize_t, u8p, C.c_ssize_t,
                                               f32p, C.c_ssize_t, C.POINTER(C.c_float)]
    lib.ssim_cuda_compute_channels.restype = C.c_int
    lib.ssim_cuda_compute_device.argtypes = [C.c_int, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                             u8p, C.c_size_t, C.c_size_t, u8p, C.c_size_t, C.c_size_t,
                                             f32p, C.c_size_t, C.c_size_t, f64p, f32p]
    lib.ssim_cuda_compute_device.restype = C.c_int
    lib.ssim_cuda_compute_u16.argtypes = lib.ssim_cuda_compute.argtypes
    lib.ssim_cuda_compute_u16.restype = C.c_int
    lib.ssim_cuda_compute_device_u16.argtypes = lib.ssim_cuda_compute_device.argtypes
    lib.ssim_cuda_compute_device_u16.restype = C.c_int
    lib.ssim_cuda_exchange_create.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_void_p]
    lib.ssim_cuda_exchange_create.restype = C.c_int
    lib.ssim_cuda_exchange_open.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.ssim_cuda_exchange_open.restype = C.c_int
    lib.ssim_cuda_exchange_close.argtypes = [C.c_int, C.c_void_p]
    lib.ssim_cuda_exchange_close.restype = C.c_int
    lib.ssim_cuda_exchange_destroy.argtypes = [C.c_int, C.c_void_p]
    lib.ssim_cuda_exchange_destroy.restype = C.c_int
    lib.ssim_cuda_exchange_enable_peer.argtypes = [C.c_int, C.c_int]
    lib.ssim_cuda_exchange_enable_peer.restype = C.c_int
    lib.ssim_cuda_compute_strip_allreduce.argtypes = [C.c_int, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                      u8p, C.c_size_t, u8p, C.c_size_t, f32p, C.c_size_t,
                                                      C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_uint64, f64p, f32p, vp]
    lib.ssim_cuda_compute_strip_allreduce.restype = C.c_int
    lib.ssim_cuda_last_launch_count.restype = C.c_int
    lib.ssim_cuda_compute_strips.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_uint32, C.c_uint32, u8p, C.c_ssize_t, C.c_ssize_t,
                                             u8p, C.c_ssize_t, C.c_ssize_t, f32p, C.c_ssize_t, C.c_ssize_t, C.POINTER(C.c_float)]
    lib.ssim_cuda_compute_strips.restype = C.c_int
    lib.ssim_cuda_synth_fill.argtypes = [C.c_int, vp, u8p, C.c_size_t, u8p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.c_uint64]
    lib.ssim_cuda_synth_fill.restype = C.c_int
    lib.ssim_cuda_set_tuning.argtypes = [C.c_int, C.c_int]
    lib.ssim_cuda_set_tuning.restype = None
    lib.ssim_cuda_debug_slot_times.argtypes = [C.c_void_p]
    lib.ssim_cuda_debug_slot_times.restype = None
    lib.ssim_cuda_debug_map_cache_entry.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_size_t, C.c_int]
    lib.ssim_cuda_debug_map_cache_entry.restype = C.c_int
    lib._bound = True
    return lib


def rmgr_lib():
    """librmgr-ssim.so with the reference API's argtypes."""
    cuda_lib()  # dependency, loaded first so the $ORIGIN rpath is not even needed
    lib = _load("librmgr-ssim.so")
    if not getattr(lib, "_bound", False):
        bind_reference_api(lib)
        lib._bound = True
    return lib


class SsimError(OSError):
    pass


def _check(rc):
    if rc != 0:
        raise SsimError(rc, "%s (%s)" % (os.strerror(rc), cuda_lib().ssim_cuda_last_error_string().decode()))


def get_version():
    v = Version()
    _check(rmgr_lib().rmgr_ssim_get_version(C.byref(v)))
    return v.major, v.minor, v.patch, v.string.decode()


def compute_ssim(a, b, want_map=False, want_ssim=True, width=None, height=None, step_a=1, step_b=1, stride_a=None,
                 stride_b=None, a_off=0, b_off=0, ssim_map=None, map_step=1, map_stride=None, map_off=0, openmp=False,
                 thread_pool=None):
    """rmgr_ssim_compute_ssim() on numpy uint8 buffers (host memory).  Returns (ssim or None, map or None).

    Mirrors the reference call: Params{width,height,imgA,imgB,ssimMap,ssimStep,ssimStride} (include/rmgr/ssim.h)."""
    if width is None:
        height, width = a.shape[:2]
    if want_map and ssim_map is None:
        ssim_map = np.empty((height, width), dtype=np.float32)
    p = make_params(a, b, width, height, step_a, stride_a, step_b, stride_b, ssim_map, map_step, map_stride, a_off, b_off, map_off)
    out = C.c_float()
    lib = rmgr_lib()
    if openmp:
        rc = lib.rmgr_ssim_compute_ssim_openmp(C.byref(out) if want_ssim else None, C.byref(p))
    else:
        rc = lib.rmgr_ssim_compute_ssim(C.byref(out) if want_ssim else None, C.byref(p), thread_pool)
    _check(rc)
    return (np.float32(out.value) if want_ssim else None), ssim_map


def compute_device(device, stream, width, src_rows, out_y0, out_rows, frames, d_a, pitch_a, fstride_a, d_b, pitch_b, fstride_b,
                   d_map=None, map_pitch=0, map_fstride=0, d_sums=None, d_ssim=None):
    """ssim_cuda_compute_device(): raw device addresses (ints), asynchronous on `stream` (int handle or None)."""
    _check(cuda_lib().ssim_cuda_compute_device(device, stream, width, src_rows, out_y0, out_rows, frames, d_a, pitch_a, fstride_a,
                                              d_b, pitch_b, fstride_b, d_map, map_pitch, map_fstride, d_sums, d_ssim))


def compute_device_u16(device, stream, width, src_rows, out_y0, out_rows, frames, d_a, pitch_a, fstride_a, d_b, pitch_b, fstride_b,
                       d_map=None, map_pitch=0, map_fstride=0, d_sums=None, d_ssim=None):
    """ssim_cuda_compute_device_u16(): 16-bit planes, pitches and frame strides in BYTES."""
    _check(cuda_lib().ssim_cuda_compute_device_u16(device, stream, width, src_rows, out_y0, out_rows, frames, d_a, pitch_a, fstride_a,
                                                  d_b, pitch_b, fstride_b, d_map, map_pitch, map_fstride, d_sums, d_ssim))


def compute_u16(a, b, want_map=False, want_ssim=True, width=None, height=None, step_a=1, step_b=1, stride_a=None, stride_b=None,
                a_off=0, b_off=0, device=0):
    """ssim_cuda_compute_u16() on numpy uint16 buffers; steps/strides/offsets in uint16 ELEMENTS.  Returns (ssim or None, map or None)."""
    if width is None:
        height, width = a.shape[:2]
    stride_a = stride_a if stride_a is not None else width * step_a
    stride_b = stride_b if stride_b is not None else width * step_b
    m = np.empty((height, width), dtype=np.float32) if want_map else None
    out = C.c_float()
    _check(cuda_lib().ssim_cuda_compute_u16(device, width, height, a.ctypes.data + 2 * a_off, step_a, stride_a,
                                           b.ctypes.data + 2 * b_off, step_b, stride_b,
                                           m.ctypes.data if want_map else None, 1, width, C.byref(out) if want_ssim else None))
    return (np.float32(out.value) if want_ssim else None), m


def exchange_create(device):
    """Allocates this rank's exchange buffer; returns (device pointer, 64-byte IPC handle)."""
    buf = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    _check(cuda_lib().ssim_cuda_exchange_create(device, C.byref(buf), handle))
    return buf.value, bytes(handle)


def exchange_open(device, handle):
    """Maps another process's exchange buffer (IPC handle from exchange_create) into this process."""
    buf = C.c_void_p()
    raw = (C.c_ubyte * 64).from_buffer_copy(handle)
    _check(cuda_lib().ssim_cuda_exchange_open(device, raw, C.byref(buf)))
    return buf.value


def compute_strip_allreduce(device, stream, width, src_rows, out_y0, out_rows, image_rows, d_a, pitch_a, d_b, pitch_b, d_map, map_pitch,
                            peer_bufs, rank, epoch, d_sum_all, d_ssim_all=None, d_status=None):
    """ssim_cuda_compute_strip_allreduce(): one strip, sum over ranks exchanged through peer memory inside the (single) kernel launch."""
    arr = (C.c_void_p * len(peer_bufs))(*peer_bufs)
    _check(cuda_lib().ssim_cuda_compute_strip_allreduce(device, stream, width, src_rows, out_y0, out_rows, image_rows, d_a, pitch_a, d_b, pitch_b,
                                                       d_map, map_pitch, arr, len(peer_bufs), rank, epoch, d_sum_all, d_ssim_all, d_status))


def compute_strips(devices, a, b, want_map=False):
    """ssim_cuda_compute_strips() on dense numpy uint8 images; returns (ssim, map or None)."""
    h, w = a.shape
    m = np.empty((h, w), dtype=np.float32) if want_map else None
    out = C.c_float()
    devs = (C.c_int * len(devices))(*devices)
    _check(cuda_lib().ssim_cuda_compute_strips(len(devices), devs, w, h, a.ctypes.data, 1, w, b.ctypes.data, 1, w,
                                              m.ctypes.data if want_map else None, 1, w, C.byref(out)))
    return np.float32(out.value), m


def compute_channels(a, b, want_map=False, device=0):
    """ssim_cuda_compute_channels() on interleaved uint8 arrays of shape (H, W, C); returns (ssim[C], map (H, W, C) or None)."""
    h, w, c = a.shape
    m = np.empty((h, w, c), dtype=np.float32) if want_map else None
    out = (C.c_float * c)()
    _check(cuda_lib().ssim_cuda_compute_channels(device, w, h, c, a.ctypes.data, w * c, b.ctypes.data, w * c,
                                                m.ctypes.data if want_map else None, w * c, out))
    return np.array(out, dtype=np.float32), m


def synth_fill(device, stream, d_a, pitch_a, d_b, pitch_b, width, rows, y0=0, frame=0, seed=0x5517):
    _check(cuda_lib().ssim_cuda_synth_fill(device, stream, d_a, pitch_a, d_b, pitch_b, width, rows, y0, frame, seed))


def imgio_lib():
    """libssim_imgio.so with argtypes declared for every symbol of include/ssim_imgio.h."""
    lib = _load("libssim_imgio.so")
    if not getattr(lib, "_bound", False):
        lib.ssim_imgio_decode_jpeg.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.ssim_imgio_decode_jpeg.restype = C.c_int
        lib.ssim_imgio_last_error.restype = C.c_char_p
        lib._bound = True
    return lib


def decode_jpeg(data):
    """JPEG bytes (bytes or a uint8 array) -> uint8 array (h, w) or (h, w, 3), decoded like the reference's image loader
    (stbi_load, src/ssim-cli.cpp:143).  Raises ValueError on a corrupt or unsupported file."""
    lib = imgio_lib()
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8)
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    rc = lib.ssim_imgio_decode_jpeg(buf.ctypes.data, buf.size, None, 0, C.byref(w), C.byref(h), C.byref(c))
    if rc:
        raise ValueError("JPEG: " + lib.ssim_imgio_last_error().decode())
    out = np.empty((h.value, w.value, c.value), dtype=np.uint8)
    rc = lib.ssim_imgio_decode_jpeg(buf.ctypes.data, buf.size, out.ctypes.data, out.size, C.byref(w), C.byref(h), C.byref(c))
    if rc:
        raise ValueError("JPEG: " + lib.ssim_imgio_last_error().decode())
    return out[:, :, 0].copy() if c.value == 1 else out
