"""A form-factor CSR slab that stays in device memory.

At the sizes BASELINE.json names the matrix is too large to move (160 GB at
200k faces; PCIe gives 11-50 GB/s per GPU), so downstream consumers -- the
Jacobi radiosity iteration of src/flux/solve.py:25-45 and the products of
src/flux/model.py:8-24 -- run on the slab where it was assembled.  Vectors are
float64 torch CUDA tensors (torch = device memory and NCCL plumbing); the
product kernel is ``csr_jacobi_kernel`` in ``csrc/spmv.cuh``.
"""
import ctypes

import numpy as np

from . import _lib, config


class DeviceCsrSlab:
    """Rows ``[row_start, row_stop)`` of an ``(m_global, n)`` matrix, resident on ``device``."""

    def __init__(self, handle, device, row_start=0, m_global=None):
        self._h, self.device = handle, device
        m, n, nnz = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        dt, iw = ctypes.c_int(), ctypes.c_int()
        _lib.check(_lib.lib().fluxb200_csr_info(handle, ctypes.byref(m), ctypes.byref(n), ctypes.byref(nnz),
                                                ctypes.byref(dt), ctypes.byref(iw), None))
        self.shape = (m.value, n.value)
        self.nnz = nnz.value
        self.dtype = np.dtype(np.float64 if dt.value == _lib.F64 else np.float32)
        self.index_dtype = np.dtype(np.int32 if iw.value == 4 else np.int64)
        self.row_start = int(row_start)
        self.row_stop = self.row_start + m.value
        self.m_global = m.value if m_global is None else int(m_global)

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib._lib is not None:
            _lib._lib.fluxb200_csr_destroy(h)

    @property
    def nbytes(self):
        return self.nnz*(self.dtype.itemsize + self.index_dtype.itemsize) + 8*(self.shape[0] + 1)

    def last_ms(self):
        ms = ctypes.c_float()
        _lib.check(_lib.lib().fluxb200_csr_info(self._h, None, None, None, None, None, ctypes.byref(ms)))
        return ms.value

    def _torch(self):
        import torch
        return torch, torch.device('cuda', self.device)

    def step(self, x, E=None, rho=1.0, want_diff=False, out=None):
        """``y = E + FF @ (rho * x)`` on this slab (x: float64 CUDA tensor of
        length n; E: length m or None; rho: scalar or length-n tensor).
        Returns ``y`` (length m) and, if asked, ``max |y - x[rows]|``."""
        torch, dev = self._torch()
        m, n = self.shape
        assert x.dtype == torch.float64 and x.is_cuda and x.numel() == n and x.is_contiguous()
        y = torch.empty(m, dtype=torch.float64, device=dev) if out is None else out
        rho_t = rho if isinstance(rho, torch.Tensor) else None
        if rho_t is not None:
            assert rho_t.dtype == torch.float64 and rho_t.numel() == n and rho_t.is_contiguous()
        if E is not None:
            assert E.dtype == torch.float64 and E.numel() == m and E.is_contiguous()
        diff = ctypes.c_double(0.0)
        torch.cuda.current_stream(dev).synchronize()        # inputs were produced on torch's stream
        _lib.check(_lib.lib().fluxb200_csr_jacobi_step(
            self._h, None if E is None else E.data_ptr(), None if rho_t is None else rho_t.data_ptr(),
            1.0 if rho_t is not None else float(rho), x.data_ptr(), y.data_ptr(),
            ctypes.byref(diff) if want_diff else None, self.row_start if want_diff else 0))
        return (y, diff.value) if want_diff else y

    def matvec(self, x):
        """``FF @ x`` for a NumPy vector or a float64 CUDA tensor (same type back)."""
        torch, dev = self._torch()
        if isinstance(x, torch.Tensor):
            return self.step(x.to(torch.float64).contiguous())
        xt = torch.as_tensor(np.ascontiguousarray(x, np.float64), device=dev)
        return self.step(xt).cpu().numpy().astype(np.result_type(self.dtype, np.asarray(x).dtype), copy=False)

    __matmul__ = matvec

    def extract(self, rows, cols):
        """``self[rows, :][:, cols]`` as a new device-resident slab
        (src/flux/compressed_form_factors.py:562).  ``rows``: local row numbers,
        ``cols``: column positions without repeats."""
        rows = np.ascontiguousarray(np.asarray(rows).astype(np.int64))
        cols = np.ascontiguousarray(np.asarray(cols).astype(np.int64))
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().fluxb200_csr_extract(self._h, _lib.ptr(rows), len(rows), _lib.ptr(cols), len(cols),
                                                   ctypes.byref(h)))
        return DeviceCsrSlab(h, self.device)

    def _thin_product(self, X, transpose):
        torch, dev = self._torch()
        m, n = self.shape
        rows_in, rows_out = (m, n) if transpose else (n, m)
        X = X.to(torch.float64)
        assert X.is_cuda and X.dim() == 2 and X.shape[0] == rows_in
        Y = torch.empty(rows_out, X.shape[1], dtype=torch.float64, device=dev)
        torch.cuda.current_stream(dev).synchronize()
        for c0 in range(0, X.shape[1], 32):            # the kernel takes up to 32 right-hand sides
            Xc = X[:, c0:c0 + 32].contiguous()
            Yc = torch.empty(rows_out, Xc.shape[1], dtype=torch.float64, device=dev)
            torch.cuda.current_stream(dev).synchronize()
            _lib.check(_lib.lib().fluxb200_csr_matmat(self._h, Xc.data_ptr(), Xc.shape[1], Yc.data_ptr(),
                                                      1 if transpose else 0))
            Y[:, c0:c0 + 32] = Yc
        return Y

    def matmat(self, X):
        """``A @ X`` for a thin dense float64 CUDA tensor ``X`` (n x k)."""
        return self._thin_product(X, False)

    def rmatmat(self, X):
        """``A.T @ X`` for a thin dense float64 CUDA tensor ``X`` (m x k)."""
        return self._thin_product(X, True)

    def to_scipy(self):
        """Download as ``scipy.sparse.csr_matrix`` (tests / small matrices)."""
        import scipy.sparse
        m, n = self.shape
        indptr = np.empty(m + 1, self.index_dtype)
        indices = np.empty(self.nnz, self.index_dtype)
        data = np.empty(self.nnz, self.dtype)
        _lib.check(_lib.lib().fluxb200_csr_to_host(self._h, _lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(data)))
        return scipy.sparse.csr_matrix((data, indices, indptr), shape=(m, n))


def get_form_factor_matrix_device(shape_model, I=None, J=None, eps=None, row_start=0, m_global=None):
    """``get_form_factor_matrix`` (src/flux/form_factors.py:11-72) with the result
    left in device memory as a :class:`DeviceCsrSlab`."""
    if eps is None:
        eps = config.DEFAULT_EPS
    # int32 column positions always suffice (n < 2^31); the slab's indptr is int64 on the device, so the
    # entry count is not limited by the index width
    m, n, _, st = shape_model._ff_assemble_device(I, J, eps, 4)
    h = ctypes.c_void_p()
    _lib.check(_lib.lib().fluxb200_ff_detach_csr(shape_model._handle, ctypes.byref(h)))
    return DeviceCsrSlab(h, shape_model.device, row_start, m_global)
