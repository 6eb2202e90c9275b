"""ctypes binding of include/agxntt.h -- the Python host side over the C ABI.

The CUDA library is REQUIRED: importing this module without lib/libagxntt.so raises, and every call goes to the
GPU kernels.  There is no CPU fallback anywhere in this package (oracle/ is test infrastructure and is never
imported from here).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_u32p = C.POINTER(C.c_uint32)
_u64p = C.POINTER(C.c_uint64)


class AgxError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: {error_string(code)} (code {code})")
        self.code = code


class _Parms(C.Structure):
    _fields_ = [("n", C.c_uint32), ("logn", C.c_uint32), ("nlimbs", C.c_uint32), ("q", _u32p)]


def _load() -> C.CDLL:
    path = os.environ.get("AGX_LIB", _build.LIB)   # AGX_LIB: experiment builds of the same ABI (kernel variants)
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(the CUDA extension is mandatory; there is no CPU fallback)")
    if "AGX_LIB" not in os.environ and _build.recorded_hash() not in (None, _build.source_hash()):
        import warnings
        warnings.warn(f"{path} was built from other sources than the ones in csrc/: rebuild with "
                      "`python -c 'import __graft_entry__ as g; g.build()'`", RuntimeWarning, stacklevel=3)
    L = C.CDLL(path)
    vp = C.c_void_p
    sigs = {
        "agx_create": [C.POINTER(vp), C.POINTER(_Parms), C.c_int],
        "agx_create_tables": [C.POINTER(vp), C.POINTER(_Parms), _u32p, C.c_int],
        "agx_set_tables": [vp, C.c_uint32, C.c_int, _u32p, _u32p],
        "agx_destroy": [vp],
        "agx_get_psi": [vp, C.c_uint32, _u32p],
        "agx_get_tables": [vp, C.c_uint32, C.c_int, _u32p, _u32p],
        "agx_ntt_fwd": [vp, vp, C.c_size_t, vp],
        "agx_ntt_inv": [vp, vp, C.c_size_t, vp],
        "agx_polymul": [vp, vp, vp, vp, C.c_size_t, vp],
        "agx_polymul_by_spectrum": [vp, vp, vp, vp, C.c_size_t, vp],
        "agx_elementwise": [vp, C.c_int, vp, vp, vp, C.c_size_t, vp],
        "agx_bitrev": [vp, vp, C.c_size_t, vp],
        "agx_ntt_fwd_host": [vp, vp, vp, C.c_size_t],
        "agx_ntt_inv_host": [vp, vp, vp, C.c_size_t],
        "agx_polymul_host": [vp, vp, vp, vp, C.c_size_t],
        "agx_polymul_by_spectrum_host": [vp, vp, vp, vp, C.c_size_t],
        "agx_host_alloc": [C.POINTER(vp), C.c_size_t],
        "agx_host_free": [vp],
        "agx_fill_synthetic": [vp, vp, C.c_size_t, C.c_uint64, C.c_size_t, vp],
        "agx_checksum": [vp, vp, C.c_size_t, C.c_size_t, _u64p, vp],
        "agx_ref_input": [vp, C.c_uint32, _u64p, _u64p, _u64p, _u64p, _u64p, C.c_uint32],
        "agx_ref_fwd": [vp, C.c_uint32],
        "agx_ref_output": [vp, _u64p, C.c_int32],
        "agx_wait": [vp],
        "agx_ref_fwd_dev": [vp, C.c_uint32, vp, vp, vp, C.c_uint64, vp, vp, C.c_uint32, vp],
        "agx_measure_butterfly_peak": [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)],
        "agx_launch_count": [vp, _u64p],
        "agx_variant": [vp, C.c_char_p, C.c_size_t],
    }
    for name, args in sigs.items():
        fn = getattr(L, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.argtypes = args
        fn.restype = C.c_int
    L.agx_error_string.argtypes = [C.c_int]
    L.agx_error_string.restype = C.c_char_p
    return L


EXPORTS = ("agx_create agx_create_tables agx_set_tables agx_ref_fwd_dev agx_measure_butterfly_peak agx_destroy agx_get_psi agx_get_tables agx_ntt_fwd agx_ntt_inv agx_polymul agx_polymul_by_spectrum agx_polymul_by_spectrum_host agx_elementwise agx_bitrev "
           "agx_ntt_fwd_host agx_ntt_inv_host agx_polymul_host agx_host_alloc agx_host_free agx_fill_synthetic "
           "agx_checksum agx_ref_input agx_ref_fwd agx_ref_output agx_wait agx_error_string agx_launch_count "
           "agx_variant").split()

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def error_string(code: int) -> str:
    return lib().agx_error_string(code).decode()


def _ck(code: int, what: str) -> None:
    if code != 0:
        raise AgxError(code, what)


def _dev_ptr(t) -> int:
    """Device pointer of a torch CUDA tensor (any 4-byte integer dtype) or a raw int address."""
    if isinstance(t, int):
        return t
    if not t.is_cuda or not t.is_contiguous() or t.element_size() != 4:
        raise ValueError("need a contiguous CUDA tensor of a 4-byte integer dtype")
    return t.data_ptr()


def _stream_ptr(stream) -> int:
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    return stream if isinstance(stream, int) else stream.cuda_stream


class Context:
    """One (n, primes) parameter set bound to one GPU: owns the device twiddle tables and the host pipeline."""

    def __init__(self, n: int, primes, device: int = 0, psi=None):
        """psi: the caller's primitive 2n-th root per limb (agx_create_tables); None = the minimal root (agx_create)."""
        self._h = C.c_void_p()
        self.n = int(n)
        self.logn = self.n.bit_length() - 1
        self.primes = [int(q) for q in primes]
        self.L = len(self.primes)
        self.device = device
        self._q = (C.c_uint32 * self.L)(*self.primes)
        p = _Parms(self.n, self.logn, self.L, C.cast(self._q, _u32p))
        if psi is None:
            _ck(lib().agx_create(C.byref(self._h), C.byref(p), device), "agx_create")
        else:
            psi = [int(v) for v in psi]
            if len(psi) != self.L:
                raise ValueError("need one psi per limb")
            arr = (C.c_uint32 * self.L)(*psi)
            _ck(lib().agx_create_tables(C.byref(self._h), C.byref(p), C.cast(arr, _u32p), device), "agx_create_tables")

    def set_tables(self, limb: int, roots, precons=None, inverse: bool = False):
        """Replace one limb's forward / inverse table by the caller's, in the reference's order (ntt.cpp:298-300)."""
        r = np.ascontiguousarray(roots, dtype=np.uint32)
        if r.size != self.n:
            raise ValueError("need n table entries")
        pp = None
        if precons is not None:
            pc = np.ascontiguousarray(precons, dtype=np.uint32)
            if pc.size != self.n:
                raise ValueError("need n table entries")
            pp = pc.ctypes.data_as(_u32p)
        _ck(lib().agx_set_tables(self._h, limb, int(inverse), r.ctypes.data_as(_u32p), pp), "agx_set_tables")

    def measure_butterfly_peak(self, kind: int = 0, threads_per_sm: int = 1024):
        """(butterflies per clock per SM, implied SM MHz) of the butterfly instruction stream alone on this GPU."""
        v, mhz = C.c_double(), C.c_double()
        _ck(lib().agx_measure_butterfly_peak(self._h, kind, threads_per_sm, C.byref(v), C.byref(mhz)),
            "agx_measure_butterfly_peak")
        return v.value, mhz.value

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().agx_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    # ---- introspection
    def psi(self, limb: int = 0) -> int:
        v = C.c_uint32()
        _ck(lib().agx_get_psi(self._h, limb, C.byref(v)), "agx_get_psi")
        return v.value

    def tables(self, limb: int = 0, inverse: bool = False):
        r = np.empty(self.n, dtype=np.uint32)
        p = np.empty(self.n, dtype=np.uint32)
        _ck(lib().agx_get_tables(self._h, limb, int(inverse), r.ctypes.data_as(_u32p), p.ctypes.data_as(_u32p)),
            "agx_get_tables")
        return r, p

    def variant(self) -> str:
        b = C.create_string_buffer(64)
        _ck(lib().agx_variant(self._h, b, 64), "agx_variant")
        return b.value.decode()

    def launch_count(self) -> int:
        v = C.c_uint64()
        _ck(lib().agx_launch_count(self._h, C.byref(v)), "agx_launch_count")
        return v.value

    def _batch(self, t) -> int:
        words = t.numel()
        if words % (self.L * self.n):
            raise ValueError("tensor is not a whole number of [L][n] polynomials")
        return words // (self.L * self.n)

    # ---- device-pointer entry points (torch CUDA tensors, in place)
    def fwd(self, t, stream=None):
        _ck(lib().agx_ntt_fwd(self._h, _dev_ptr(t), self._batch(t), _stream_ptr(stream)), "agx_ntt_fwd")
        return t

    def inv(self, t, stream=None):
        _ck(lib().agx_ntt_inv(self._h, _dev_ptr(t), self._batch(t), _stream_ptr(stream)), "agx_ntt_inv")
        return t

    def polymul(self, c, a, b, stream=None):
        B = self._batch(a)
        if self._batch(b) != B or self._batch(c) != B:
            raise ValueError("shape mismatch")
        _ck(lib().agx_polymul(self._h, _dev_ptr(c), _dev_ptr(a), _dev_ptr(b), B, _stream_ptr(stream)), "agx_polymul")
        return c

    def polymul_by_spectrum(self, c, a, b_hat, stream=None):
        """c = INTT(NTT(a) .* b_hat) with b_hat = fwd(b): the product when one operand is kept in evaluation form."""
        B = self._batch(a)
        if self._batch(b_hat) != B or self._batch(c) != B:
            raise ValueError("shape mismatch")
        _ck(lib().agx_polymul_by_spectrum(self._h, _dev_ptr(c), _dev_ptr(a), _dev_ptr(b_hat), B, _stream_ptr(stream)),
            "agx_polymul_by_spectrum")
        return c

    EW = {"add": 0, "sub": 1, "mul": 2, "mac": 3}

    def elementwise(self, op: str, c, a, b, stream=None):
        """c = a + b | a - b | a * b | c + a * b, modulo each limb's prime (operands reduced, device tensors)."""
        B = self._batch(a)
        if self._batch(b) != B or self._batch(c) != B:
            raise ValueError("shape mismatch")
        _ck(lib().agx_elementwise(self._h, self.EW[op], _dev_ptr(c), _dev_ptr(a), _dev_ptr(b), B, _stream_ptr(stream)),
            "agx_elementwise")
        return c

    def bitrev(self, t, stream=None):
        """Permute every polynomial of t by bit reversal, in place (natural <-> bit-reversed coefficient order)."""
        _ck(lib().agx_bitrev(self._h, _dev_ptr(t), self._batch(t), _stream_ptr(stream)), "agx_bitrev")
        return t

    def fill_synthetic(self, t, seed: int = 42, first_poly: int = 0, stream=None):
        _ck(lib().agx_fill_synthetic(self._h, _dev_ptr(t), self._batch(t), seed, first_poly, _stream_ptr(stream)),
            "agx_fill_synthetic")
        return t

    def checksum(self, t, first_index: int = 0, stream=None) -> int:
        v = C.c_uint64()
        _ck(lib().agx_checksum(self._h, _dev_ptr(t), t.numel(), first_index, C.byref(v), _stream_ptr(stream)),
            "agx_checksum")
        return v.value

    # ---- host-pointer entry points (numpy uint32 arrays or raw addresses)
    @staticmethod
    def _host_ptr(a) -> int:
        if isinstance(a, int):
            return a
        if isinstance(a, np.ndarray):
            if a.dtype != np.uint32 or not a.flags.c_contiguous:
                raise ValueError("need a C-contiguous uint32 array")
            return a.ctypes.data
        if hasattr(a, "data_ptr"):   # pinned torch CPU tensor
            if a.is_cuda or not a.is_contiguous() or a.element_size() != 4:
                raise ValueError("need a contiguous 4-byte CPU tensor")
            return a.data_ptr()
        raise TypeError(type(a))

    def _host_batch(self, a, B):
        if B is not None:
            return B
        words = a.size if isinstance(a, np.ndarray) else a.numel()
        if words % (self.L * self.n):
            raise ValueError("array is not a whole number of [L][n] polynomials")
        return words // (self.L * self.n)

    def fwd_host(self, src, dst=None, B=None):
        dst = src if dst is None else dst
        _ck(lib().agx_ntt_fwd_host(self._h, self._host_ptr(src), self._host_ptr(dst), self._host_batch(src, B)),
            "agx_ntt_fwd_host")
        return dst

    def inv_host(self, src, dst=None, B=None):
        dst = src if dst is None else dst
        _ck(lib().agx_ntt_inv_host(self._h, self._host_ptr(src), self._host_ptr(dst), self._host_batch(src, B)),
            "agx_ntt_inv_host")
        return dst

    def polymul_host(self, c, a, b, B=None):
        _ck(lib().agx_polymul_host(self._h, self._host_ptr(c), self._host_ptr(a), self._host_ptr(b),
                                   self._host_batch(a, B)), "agx_polymul_host")
        return c

    def polymul_by_spectrum_host(self, c, a, b_hat, B=None):
        _ck(lib().agx_polymul_by_spectrum_host(self._h, self._host_ptr(c), self._host_ptr(a), self._host_ptr(b_hat),
                                               self._host_batch(a, B)), "agx_polymul_by_spectrum_host")
        return c


class RefPipeline:
    """Reference-shaped u64 forward pipeline: ntt_input_kernel / fwd_ntt_kernel<0> / ntt_output_kernel / q.wait()
    (include/kernel/ntt.h:32-45, src/main.cpp:60-74) over agx_ref_*."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _ck(lib().agx_create(C.byref(self._h), None, device), "agx_create")
        self._keep = []

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().agx_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    @staticmethod
    def _p(a):
        assert a.dtype == np.uint64 and a.flags.c_contiguous
        return a.ctypes.data_as(_u64p)

    def ntt_input_kernel(self, in1, in2, modulus, twiddles, precons, num_frames: int):
        N = len(twiddles)
        arrs = [np.ascontiguousarray(x, dtype=np.uint64) for x in (in1, in2, modulus, twiddles, precons)]
        # agx_ref_input takes bare pointers: the buffer-size contract of ntt_input_kernel (main.cpp:26-37) is checked here
        if len(arrs[4]) != N or len(arrs[2]) < 1 or len(arrs[0]) < N * num_frames or len(arrs[1]) < N * num_frames:
            raise ValueError("buffer sizes do not describe numFrames x N")
        self._n = N
        self._check_out()
        self._keep += arrs
        _ck(lib().agx_ref_input(self._h, N, *[self._p(a) for a in arrs], num_frames), "agx_ref_input")

    def fwd_ntt_kernel(self, compute_unit: int = 0):
        _ck(lib().agx_ref_fwd(self._h, compute_unit), "agx_ref_fwd")

    def _check_out(self):
        if getattr(self, "_n", 0) and getattr(self, "_out", None) is not None:
            if self._out[0] < self._out[1] * self._n:
                self._out = None
                raise ValueError("out is smaller than numFrames x N")

    def ntt_output_kernel(self, out: np.ndarray, num_frames: int):
        if num_frames < 0:
            raise ValueError("negative numFrames")
        self._out = (out.size, num_frames)
        self._check_out()
        self._keep.append(out)
        _ck(lib().agx_ref_output(self._h, self._p(out), num_frames), "agx_ref_output")

    def wait(self):
        try:
            _ck(lib().agx_wait(self._h), "agx_wait")
        finally:
            self._keep.clear()
            self._n, self._out = 0, None

    def fwd_dev(self, N: int, d_in, d_in2, d_out, modulus: int, d_tw, d_pre, frames: int, stream=None):
        """agx_ref_fwd_dev on torch CUDA int64 tensors (or raw device addresses)."""
        def ptr(t):
            return t if isinstance(t, int) else t.data_ptr()
        _ck(lib().agx_ref_fwd_dev(self._h, N, ptr(d_in), ptr(d_in2), ptr(d_out), modulus, ptr(d_tw), ptr(d_pre), frames,
                                  _stream_ptr(stream)), "agx_ref_fwd_dev")

    def launch_count(self) -> int:
        v = C.c_uint64()
        _ck(lib().agx_launch_count(self._h, C.byref(v)), "agx_launch_count")
        return v.value
