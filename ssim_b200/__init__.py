"""ssim_b200 -- B200-native drop-in for romigrou/ssim's rmgr::ssim::compute_ssim() hot path.

The product is the pair of shared libraries built from ssim_b200/csrc (libssim_cuda.so: C-ABI shim +
sm_100a kernels; librmgr-ssim.so: the reference's C/C++ API on top of it).  This Python package is a thin
ctypes binding used by tests, bench.py and __graft_entry__.py; see ssim_b200.api."""
__version__ = "0.1.0"
