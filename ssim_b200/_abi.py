"""ctypes mirror of the reference's C ABI (include/rmgr/ssim.h:428-533 in the reference; our
include/rmgr/ssim.h re-declares it identically).  Shared by the product bindings (ssim_b200.api) and by
the test-only loaders of the oracle / reference builds, because the layouts are the contract."""
import ctypes as C

AllocFct = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_size_t)
DeallocFct = C.CFUNCTYPE(None, C.c_void_p)
ThreadFct = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32)
ThreadPoolFct = C.CFUNCTYPE(C.c_int32, C.c_void_p, ThreadFct, C.POINTER(C.c_void_p), C.c_uint32, C.c_uint32)


class Version(C.Structure):
    _fields_ = [("major", C.c_uint32), ("minor", C.c_uint32), ("patch", C.c_uint32), ("string", C.c_char_p)]


class ImgParams(C.Structure):  # 24 bytes on LP64
    _fields_ = [("topLeft", C.c_void_p), ("step", C.c_ssize_t), ("stride", C.c_ssize_t)]


class Params(C.Structure):  # rmgr_ssim_Params, 96 bytes on LP64
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("imgA", ImgParams),
        ("imgB", ImgParams),
        ("ssimMap", C.c_void_p),
        ("ssimStep", C.c_ssize_t),
        ("ssimStride", C.c_ssize_t),
        ("alloc", C.c_void_p),
        ("dealloc", C.c_void_p),
    ]


class ThreadPool(C.Structure):  # 24 bytes on LP64
    _fields_ = [("dispatch", C.c_void_p), ("context", C.c_void_p), ("threadCount", C.c_uint32)]


assert C.sizeof(ImgParams) == 24 and C.sizeof(Params) == 96 and C.sizeof(ThreadPool) == 24


def bind_reference_api(lib):
    """Declare the argtypes of the reference C API on a loaded library (ours or the reference's)."""
    lib.rmgr_ssim_get_version.argtypes = [C.POINTER(Version)]
    lib.rmgr_ssim_get_version.restype = C.c_int32
    lib.rmgr_ssim_init_interleaved.argtypes = [C.POINTER(ImgParams), C.c_void_p, C.c_ssize_t, C.c_uint32, C.c_uint32]
    lib.rmgr_ssim_init_interleaved.restype = C.c_int32
    lib.rmgr_ssim_init_planar.argtypes = [C.POINTER(ImgParams), C.POINTER(C.c_void_p), C.POINTER(C.c_ssize_t), C.c_uint32]
    lib.rmgr_ssim_init_planar.restype = C.c_int32
    lib.rmgr_ssim_use_default_allocator.argtypes = [C.POINTER(Params)]
    lib.rmgr_ssim_use_default_allocator.restype = C.c_int32
    lib.rmgr_ssim_compute_ssim.argtypes = [C.POINTER(C.c_float), C.POINTER(Params), C.POINTER(ThreadPool)]
    lib.rmgr_ssim_compute_ssim.restype = C.c_int32
    lib.rmgr_ssim_compute_ssim_openmp.argtypes = [C.POINTER(C.c_float), C.POINTER(Params)]
    lib.rmgr_ssim_compute_ssim_openmp.restype = C.c_int32
    return lib


def make_params(a, b, width, height, step_a=1, stride_a=None, step_b=1, stride_b=None,
                ssim_map=None, map_step=1, map_stride=None, a_off=0, b_off=0, map_off=0):
    """Build a Params for numpy uint8 buffers `a`, `b` (any shape; addresses are base + offset) and an
    optional float32 buffer `ssim_map`.  Offsets are in bytes for images and in floats for the map."""
    p = Params()
    p.width, p.height = width, height
    p.imgA.topLeft = a.ctypes.data + a_off
    p.imgA.step = step_a
    p.imgA.stride = stride_a if stride_a is not None else width * step_a
    p.imgB.topLeft = b.ctypes.data + b_off
    p.imgB.step = step_b
    p.imgB.stride = stride_b if stride_b is not None else width * step_b
    if ssim_map is not None:
        p.ssimMap = ssim_map.ctypes.data + 4 * map_off
        p.ssimStep = map_step
        p.ssimStride = map_stride if map_stride is not None else width * map_step
    else:
        p.ssimMap = None
    p.alloc = None
    p.dealloc = None
    return p
