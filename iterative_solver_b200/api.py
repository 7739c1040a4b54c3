"""Python face of the C ABI (include/itsolv_b200.h): a Context bound to one GPU and thin methods that pass device
pointers of torch tensors to the CUDA kernels. torch is used for device memory and streams only; every operation below
runs in libitsolv_b200.so, and raises if that library or a Blackwell GPU is missing."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _native as N


class BackendError(RuntimeError):
    pass


def _ptr(t) -> int:
    """device pointer of a torch tensor (float64, contiguous, 1-D) or a raw integer address"""
    if isinstance(t, int):
        return t
    if t.dtype.__str__() != "torch.float64" or not t.is_cuda or not t.is_contiguous():
        raise TypeError("expected a contiguous CUDA float64 tensor")
    return t.data_ptr()


def _ptr_array(ts: Sequence) -> C.Array:
    arr = (C.c_void_p * max(1, len(ts)))()
    for i, t in enumerate(ts):
        arr[i] = None if t is None else _ptr(t)
    return arr


def _dbl(a: np.ndarray):
    return a.ctypes.data_as(N.c_double_p)


class Context:
    """One per process and GPU: stream, workspaces, optional NCCL communicator (row-sharded vectors)."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = N.kernels()
        h = C.c_void_p()
        rc = (self.lib.itsolv_ctx_create(device, C.byref(h)) if stream is None else
              self.lib.itsolv_ctx_create_on_stream(device, C.c_void_p(stream), C.byref(h)))
        if rc:
            raise BackendError(self.lib.itsolv_last_error().decode())
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, "handle", None):
            self.lib.itsolv_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc:
            raise BackendError(self.lib.itsolv_last_error().decode())

    # ---- plumbing
    @property
    def stream(self) -> int:
        return int(self.lib.itsolv_ctx_stream(self.handle) or 0)

    def synchronize(self):
        self._check(self.lib.itsolv_ctx_synchronize(self.handle))

    def set_option(self, name: str, value: int) -> int:
        return self.lib.itsolv_ctx_set_option(self.handle, name.encode(), int(value))

    def mem_usage(self, reset_peak: bool = False):
        live, peak = C.c_size_t(), C.c_size_t()
        self._check(self.lib.itsolv_mem_usage(self.handle, C.byref(live), C.byref(peak), int(reset_peak)))
        return live.value, peak.value

    def alloc(self, n: int) -> int:
        """n doubles from the context's stream-ordered pool (where the solver's vectors live); returns the device address"""
        p = C.c_void_p()
        self._check(self.lib.itsolv_alloc(self.handle, int(n), C.byref(p)))
        return int(p.value)

    def free(self, address: int):
        self._check(self.lib.itsolv_free(self.handle, C.c_void_p(address)))

    def mem_info(self):
        """(free, total) bytes of the device as the driver sees them"""
        free, total = C.c_size_t(), C.c_size_t()
        self._check(self.lib.itsolv_mem_info(self.handle, C.byref(free), C.byref(total)))
        return free.value, total.value

    def mem_trim(self):
        """give the pool's unused memory back to the driver"""
        self._check(self.lib.itsolv_mem_trim(self.handle))

    def counters(self) -> N.Counters:
        c = N.Counters()
        self.lib.itsolv_ctx_counters(self.handle, C.byref(c))
        return c

    def reset_counters(self):
        self.lib.itsolv_ctx_reset_counters(self.handle)

    def set_profiling(self, on: bool):
        self.lib.itsolv_ctx_set_profiling(self.handle, 1 if on else 0)

    def timer_start(self, timer: int = 1):
        """CUDA-event stopwatch on the context's stream (timer 0 belongs to the solve harness)"""
        self._check(self.lib.itsolv_ctx_timer_start(self.handle, timer))

    def timer_stop(self, timer: int = 1) -> float:
        ms = C.c_double()
        self._check(self.lib.itsolv_ctx_timer_stop(self.handle, timer, C.byref(ms)))
        return ms.value

    def init_comm(self, rank: int, nranks: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, N.UNIQUE_ID_BYTES)
        self._check(self.lib.itsolv_comm_init(self.handle, rank, nranks, buf))

    def unique_id(self) -> bytes:
        buf = C.create_string_buffer(N.UNIQUE_ID_BYTES)
        self._check(self.lib.itsolv_comm_unique_id(buf))
        return buf.raw

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(N.IPC_HANDLE_BYTES)
        self._check(self.lib.itsolv_comm_p2p_export(self.handle, buf))
        return buf.raw

    def p2p_import(self, handles: bytes) -> bool:
        """map the other ranks' exchange buffers; False (nothing left mapped) when a peer cannot be reached"""
        buf = C.create_string_buffer(handles, len(handles))
        return self.lib.itsolv_comm_p2p_import(self.handle, buf) == 0

    def p2p_disable(self):
        self._check(self.lib.itsolv_comm_p2p_disable(self.handle))

    @property
    def rank(self) -> int:
        return self.lib.itsolv_comm_rank(self.handle)

    @property
    def nranks(self) -> int:
        return self.lib.itsolv_comm_size(self.handle)

    def allreduce_host(self, values: np.ndarray, op_max: bool = False) -> np.ndarray:
        v = np.ascontiguousarray(values, dtype=np.float64).copy()
        self._check(self.lib.itsolv_comm_allreduce_host(self.handle, _dbl(v), v.size, 1 if op_max else 0))
        return v

    # ---- the contract
    def fill(self, alpha: float, x):
        self._check(self.lib.itsolv_fill_f64(self.handle, alpha, _ptr(x), x.numel()))

    def scal(self, alpha: float, x):
        self._check(self.lib.itsolv_scal_f64(self.handle, alpha, _ptr(x), x.numel()))

    def copy(self, dst, src):
        self._check(self.lib.itsolv_copy_f64(self.handle, _ptr(dst), _ptr(src), src.numel()))

    def axpy(self, alpha: float, x, y):
        self._check(self.lib.itsolv_axpy_f64(self.handle, alpha, _ptr(x), _ptr(y), x.numel()))

    def scal_batch(self, alpha: Sequence[float], xs: Sequence):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        self._check(self.lib.itsolv_scal_batch_f64(self.handle, _dbl(a), _ptr_array(xs), len(xs), xs[0].numel()))

    def fill_batch(self, alpha: Sequence[float], xs: Sequence):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        self._check(self.lib.itsolv_fill_batch_f64(self.handle, _dbl(a), _ptr_array(xs), len(xs), xs[0].numel()))

    def axpy_batch(self, alpha: Sequence[float], xs: Sequence, ys: Sequence):
        a = np.ascontiguousarray(alpha, dtype=np.float64)
        self._check(self.lib.itsolv_axpy_batch_f64(self.handle, _dbl(a), _ptr_array(xs), _ptr_array(ys), len(ys),
                                                   ys[0].numel()))

    def mgs_step(self, inv_norm: float, ri, ov: Sequence[float], rjs: Sequence):
        o = np.ascontiguousarray(ov, dtype=np.float64)
        self._check(self.lib.itsolv_mgs_step_f64(self.handle, inv_norm, _ptr(ri), _dbl(o) if len(rjs) else None,
                                                 _ptr_array(rjs), len(rjs), ri.numel()))

    def mgs_step_dots(self, inv_norm: float, ri, ov: Sequence[float], rjs: Sequence) -> np.ndarray:
        """mgs_step that also returns [<ri', ri'>, <rjs[0]', rjs[t]'> for t in range(len(rjs))]"""
        o = np.ascontiguousarray(ov, dtype=np.float64)
        dots = np.zeros(len(rjs) + 1)
        self._check(self.lib.itsolv_mgs_step_dots_f64(self.handle, inv_norm, _ptr(ri), _dbl(o) if len(rjs) else None,
                                                      _ptr_array(rjs), len(rjs), ri.numel(), _dbl(dots)))
        return dots

    def mgs_chain(self, rs: Sequence, thresh: float = 1e-10) -> np.ndarray:
        """R-R modified Gram-Schmidt of the vectors as one chain of launches (itsolv_mgs_chain_f64); returns the rows of
        inner products described in include/itsolv_b200.h"""
        w = len(rs)
        rows = np.zeros(w + w * (w + 1) // 2)
        self._check(self.lib.itsolv_mgs_chain_f64(self.handle, _ptr_array(rs), w, rs[0].numel(), thresh, _dbl(rows)))
        return rows

    def project_mgs_chain(self, alpha: np.ndarray, xx: Sequence, yy: Sequence, yscale, keep: Sequence[int],
                          thresh: float = 1e-10) -> np.ndarray:
        """projection of all yy against xx (scaled by yscale when given) and the Gram-Schmidt chain of yy[keep] as one chain
        of launches (itsolv_project_mgs_chain_f64); returns the rows of inner products as mgs_chain does"""
        k, m, w = len(xx), len(yy), len(keep)
        a = np.ascontiguousarray(alpha, dtype=np.float64).reshape(k, m)
        s = np.ascontiguousarray(yscale, dtype=np.float64) if yscale is not None else None
        kp = (C.c_int * w)(*[int(i) for i in keep])
        rows = np.zeros(w + w * (w + 1) // 2)
        self._check(self.lib.itsolv_project_mgs_chain_f64(self.handle, _dbl(a), k, m, _ptr_array(xx), _ptr_array(yy),
                                                          _dbl(s) if s is not None else None, kp, w, yy[0].numel(), thresh,
                                                          _dbl(rows)))
        return rows

    def dot(self, x, y) -> float:
        r = C.c_double()
        self._check(self.lib.itsolv_dot_f64(self.handle, _ptr(x), _ptr(y), x.numel(), C.byref(r)))
        return r.value

    def gemm_inner(self, xx: Sequence, yy: Sequence, n: int | None = None) -> np.ndarray:
        k, m = len(xx), len(yy)
        out = np.zeros((k, m))
        if k == 0 or m == 0:
            return out
        n = xx[0].numel() if n is None else n
        self._check(self.lib.itsolv_gemm_inner_f64(self.handle, _ptr_array(xx), k, _ptr_array(yy), m, n, _dbl(out)))
        return out

    def gemm_outer(self, alpha: np.ndarray, xx: Sequence, yy: Sequence, beta_zero: bool = False, n: int | None = None):
        k, m = len(xx), len(yy)
        a = np.ascontiguousarray(alpha, dtype=np.float64).reshape(k, m)
        if m == 0:
            return
        n = yy[0].numel() if n is None else n
        self._check(self.lib.itsolv_gemm_outer_f64(self.handle, _dbl(a), k, m, _ptr_array(xx), _ptr_array(yy), n,
                                                   1 if beta_zero else 0))

    def gemm_outer_scaled(self, alpha: np.ndarray, xx: Sequence, yy: Sequence, yscale: Sequence[float]):
        """yy[j] = yscale[j]*yy[j] + sum_i alpha[i,j] xx[i]  (scal_batch followed by gemm_outer, in one pass)"""
        k, m = len(xx), len(yy)
        a = np.ascontiguousarray(alpha, dtype=np.float64).reshape(k, m)
        s = np.ascontiguousarray(yscale, dtype=np.float64)
        if m == 0:
            return
        self._check(self.lib.itsolv_gemm_outer_scaled_f64(self.handle, _dbl(a), k, m, _ptr_array(xx), _ptr_array(yy),
                                                          yy[0].numel(), _dbl(s)))

    def precondition(self, residuals: Sequence, diag, shift: Sequence[float]):
        s = np.ascontiguousarray(shift, dtype=np.float64)
        self._check(self.lib.itsolv_precondition_f64(self.handle, _ptr_array(residuals), len(residuals), _ptr(diag),
                                                     _dbl(s), diag.numel()))

    def davidson_residual(self, coef: np.ndarray, q: Sequence, a: Sequence, lam: Sequence[float], out_r: Sequence,
                          diag=None, out_x: Sequence | None = None):
        """Fused solution/residual/norm/preconditioner pass (itsolv_davidson_residual_f64); returns
        (<r_j, r_j> before preconditioning, <out_r_j, out_r_j>)."""
        k, m = len(q), len(out_r)
        c = np.ascontiguousarray(coef, dtype=np.float64).reshape(k, m)
        l = np.ascontiguousarray(lam, dtype=np.float64)
        n2, n2w = np.zeros(m), np.zeros(m)
        self._check(self.lib.itsolv_davidson_residual_f64(
            self.handle, _dbl(c), k, m, _ptr_array(q), _ptr_array(a), _dbl(l), _ptr(diag) if diag is not None else None,
            _dbl(l), _ptr_array(out_x) if out_x is not None else None, _ptr_array(out_r), out_r[0].numel(), _dbl(n2),
            _dbl(n2w)))
        return n2, n2w

    def subspace_residual(self, coef: np.ndarray, q: Sequence, a: Sequence, out_r: Sequence, lam=None, rhs=None,
                          rscale=None, diag=None, shift=None, out_x: Sequence | None = None, accumulate: bool = False,
                          mode: int | None = None):
        """itsolv_subspace_residual_f64. Eigenproblem form with `lam` (r = sum c a - lam x), linear-equations form with
        `rhs` and `rscale` (r = (sum c a - rhs) * rscale); `accumulate`: continue from the contents of out_x / out_r;
        `mode` 2: r = sum c a as it is, 3: the same and out_x = x - out_r (the DIIS step).
        Returns (<r_j, r_j> before preconditioning, <out_r_j, out_r_j>)."""
        k, m = len(q), len(out_r)
        if mode is None:
            mode = 0 if rhs is None else 1
        c = np.ascontiguousarray(coef, dtype=np.float64).reshape(k, m)
        l = np.ascontiguousarray(lam, dtype=np.float64) if lam is not None else None
        s = np.ascontiguousarray(rscale, dtype=np.float64) if rscale is not None else None
        sh = np.ascontiguousarray(shift if shift is not None else (lam if lam is not None else np.zeros(m)), dtype=np.float64)
        n2, n2w = np.zeros(m), np.zeros(m)
        self._check(self.lib.itsolv_subspace_residual_f64(
            self.handle, mode, int(accumulate), _dbl(c), k, m, _ptr_array(q), _ptr_array(a), _dbl(l) if l is not None else None,
            _ptr_array(rhs) if rhs is not None else None, _dbl(s) if s is not None else None,
            _ptr(diag) if diag is not None else None, _dbl(sh), _ptr_array(out_x) if out_x is not None else None,
            _ptr_array(out_r), out_r[0].numel(), _dbl(n2), _dbl(n2w)))
        return n2, n2w

    def select(self, x, nsel: int, max: bool = False, ignore_sign: bool = False, y=None, global_offset: int = 0):
        idx = np.zeros(nsel, dtype=np.int64)
        val = np.zeros(nsel)
        found = C.c_int()
        self._check(self.lib.itsolv_select_f64(self.handle, _ptr(x), _ptr(y) if y is not None else None, x.numel(),
                                               global_offset, nsel, int(max), int(ignore_sign),
                                               idx.ctypes.data_as(N.c_int64_p), _dbl(val), C.byref(found)))
        return idx[:found.value].copy(), val[:found.value].copy()

    def banded_apply(self, x, y, n_global: int, row_offset: int, b: int, eps: float, x_lo=None, x_hi=None):
        self._check(self.lib.itsolv_banded_apply_f64(self.handle, n_global, row_offset, x.numel(), b, eps, _ptr(x),
                                                     _ptr(x_lo) if x_lo is not None else None,
                                                     _ptr(x_hi) if x_hi is not None else None, _ptr(y)))


    def csr_apply_multi(self, row_ptr, col, val, xs: Sequence, ys: Sequence, n_global: int, row_offset: int, b: int,
                        x_lo: Sequence | None = None, x_hi: Sequence | None = None):
        """ys[k] = A xs[k] for the rows of this shard (itsolv_csr_apply_multi_f64); row_ptr int64, col int32, val
        float64 device tensors; x_lo / x_hi: per vector the b halo rows below / above the shard (or None)."""
        self._check(self.lib.itsolv_csr_apply_multi_f64(
            self.handle, n_global, row_offset, ys[0].numel(), b, _ptr(row_ptr), _ptr(col), _ptr(val), len(xs),
            _ptr_array(xs), _ptr_array(x_lo) if x_lo is not None else None,
            _ptr_array(x_hi) if x_hi is not None else None, _ptr_array(ys)))


def distribution(n: int, nranks: int) -> np.ndarray:
    """chunk borders of the reference's make_distribution_spread_remainder (array/util/Distribution.h:99-110)"""
    b = np.zeros(nranks + 1, dtype=np.int64)
    N.kernels().itsolv_distribution(n, nranks, b.ctypes.data_as(N.c_int64_p))
    return b


def select_merge(idx: np.ndarray, val: np.ndarray, nsel: int, max: bool = False, ignore_sign: bool = False):
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    val = np.ascontiguousarray(val, dtype=np.float64)
    oi = np.zeros(nsel, dtype=np.int64)
    ov = np.zeros(nsel)
    c = N.kernels().itsolv_select_merge(idx.ctypes.data_as(N.c_int64_p), _dbl(val), idx.size, nsel, int(max),
                                        int(ignore_sign), oi.ctypes.data_as(N.c_int64_p), _dbl(ov))
    return oi[:c], ov[:c]
