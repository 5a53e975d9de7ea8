"""ctypes binding of ``libstair_b200.so`` (the C-ABI boundary, include/stair_b200.h).

There is no fallback: if the library is missing, or a call returns a non-zero status, this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstair_b200.so')

BF16, F32 = 0, 1
ACT_NONE, ACT_RELU = 0, 1

_ERRORS = {-1: 'bad argument', -2: 'CUDA error', -3: 'workspace / arena capacity exceeded',
           -4: 'invalid program layout', -5: 'unsupported configuration'}

_lib = None


class StairError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise StairError('%s not found: run `python -m stair_b200.build` (sm_100a CUDA build); '
                             'stair_b200 has no CPU or PyTorch fallback' % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def ptr(t):
    """Raw device/host pointer of a tensor (or None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(int(t))


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return ctypes.c_void_p(s.cuda_stream)


def check(rc, what):
    if rc != 0:
        extra = ''
        if rc == -2:
            try:
                extra = ' (gemm error flag %d)' % lib().stair_gemm_error_flag()
            except Exception:
                pass
        raise StairError('%s failed: %s%s' % (what, _ERRORS.get(rc, 'status %d' % rc), extra))


def dtype_code(dt):
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float32:
        return F32
    raise StairError('unsupported dtype %s' % dt)


i32, i64 = ctypes.c_int, ctypes.c_longlong


def gemm(A, W, bias=None, out=None, out_dtype=torch.bfloat16, act=ACT_NONE, row_scale=None, accumulate=False,
         M=None, N=None, K=None, nplanes=1, a_plane_rows=0, w_plane_rows=0, lda=None, ldw=None, ldc=None, stream=None):
    """out[M,N] = act(row_scale * (A[M,K] @ W[N,K]^T) + bias) via ``stair_gemm_bf16`` (tcgen05/TMA kernel)."""
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = W.shape[0] if N is None else N
    lda = A.stride(0) if lda is None else lda
    ldw = W.stride(0) if ldw is None else ldw
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=out_dtype)
    ldc = out.stride(0) if ldc is None else ldc
    rc = lib().stair_gemm_bf16(ptr(A), i64(lda), i32(a_plane_rows), ptr(W), i64(ldw), i32(w_plane_rows), i32(nplanes),
                               ptr(bias), ptr(row_scale), ptr(out), i64(ldc), i32(dtype_code(out.dtype)),
                               i32(M), i32(N), i32(K), i32(act), i32(1 if accumulate else 0), stream_ptr(stream))
    check(rc, 'stair_gemm_bf16')
    return out
