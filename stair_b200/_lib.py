"""ctypes binding of ``libstair_b200.so`` (the C-ABI boundary, include/stair_b200.h).

There is no fallback: if the library is missing, or a call returns a non-zero status, this raises.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstair_b200.so')
HEADER_PATH = os.path.join(os.path.dirname(_HERE), 'include', 'stair_b200.h')

BF16, F32 = 0, 1
ACT_NONE, ACT_RELU = 0, 1
FWD_ENCODE_VIDEO, FWD_ENCODE_TEXT, FWD_GROUP, FWD_MODULES, FWD_DECODE, FWD_ALL = 1, 2, 4, 8, 16, 31
MAX_GROUP_DEPS = 8          # STAIR_MAX_GROUP_DEPS

_ERRORS = {-1: 'bad argument', -2: 'CUDA error', -3: 'workspace / arena capacity exceeded',
           -4: 'invalid program layout', -5: 'unsupported configuration'}

i32, i64 = ctypes.c_int, ctypes.c_longlong
vp = ctypes.c_void_p


def _parse_enum(name):
    """Read ``enum <name> {...}`` from the public header so Python and C can never disagree on the numbering."""
    src = open(HEADER_PATH).read()
    body = re.search(r'enum\s+%s\s*\{(.*?)\};' % name, src, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    out, nxt = {}, 0
    for item in body.split(','):
        item = item.strip()
        if not item:
            continue
        if '=' in item:
            key, expr = (s.strip() for s in item.split('=', 1))
            nxt = int(eval(expr, {}, dict(out)))      # expressions only reference earlier enumerators
        else:
            key = item
        out[key] = nxt
        nxt += 1
    return out


OP = {k[len('STAIR_OP_'):]: v for k, v in _parse_enum('StairOp').items()}
W = {k[len('STAIR_W_'):]: v for k, v in _parse_enum('StairWeight').items()}
W_COUNT = W['COUNT']


class StairModel(ctypes.Structure):
    _fields_ = [('T_max', i32), ('V', i32), ('V_ld', i32), ('H', i32), ('text_size', i32), ('text_ld', i32), ('A', i32),
                ('O', i32), ('conv_k', i32), ('precision', i32), ('w', vp * W_COUNT), ('wt', vp * W_COUNT)]


class StairGroup(ctypes.Structure):
    _fields_ = [('op', i32), ('variant', i32), ('level', i32), ('count', i32), ('node_off', i32), ('out_base', i32),
                ('out_mult', i32), ('aux_base', i32), ('head', i32)]


class StairBatch(ctypes.Structure):
    _fields_ = [('B', i32), ('T', i32), ('n_tok', i32), ('L_max', i32), ('n_nodes', i32), ('n_groups', i32),
                ('video_dtype', i32), ('question_dtype', i32), ('video', vp), ('question', vp), ('q_off', vp),
                ('node_gid', vp), ('node_q', vp), ('node_arg', vp), ('node_span', vp), ('root_node', vp),
                ('groups', ctypes.POINTER(StairGroup)), ('group_tab', vp), ('group_deps', vp),
                ('q_order', vp), ('q_soff', vp), ('tok_src', vp)]


class StairBuffers(ctypes.Structure):
    _fields_ = [('vid', vp), ('vid_slots', i64), ('vec', vp), ('vec_rows', i64), ('att', vp), ('att_rows', i64),
                ('tokfeat', vp), ('qfeat', vp), ('logits', vp), ('answers', vp), ('head_small', vp), ('head_vec', vp),
                ('head_ff', vp), ('itab', vp), ('itab_ints', i64), ('workspace', vp), ('workspace_bytes', i64),
                ('status', vp)]


class StairItabLayout(ctypes.Structure):
    _fields_ = [('perm', i64), ('out_slot', i64), ('aux_slot', i64), ('arg_slot', i64), ('pos_q', i64), ('pos_span', i64),
                ('group_off', i64), ('total', i64)]


class StairTrain(ctypes.Structure):
    _fields_ = [('grad', vp * W_COUNT),
                ('n_att', i32), ('att_node', vp), ('att_kind', vp), ('att_slot', vp), ('att_gold', vp), ('att_w', vp),
                ('n_bin', i32), ('bin_node', vp), ('bin_which', vp), ('bin_label', vp), ('bin_w', vp),
                ('n_con', i32), ('con_node', vp), ('con_pos', vp), ('con_w', vp), ('n_cls', i32), ('cls_rep', vp),
                ('answer', vp), ('dec_w', ctypes.c_float), ('loss', vp),
                ('dvid', vp), ('dvec', vp), ('datt', vp), ('dtokfeat', vp), ('dqfeat', vp), ('dlogits', vp),
                ('saved', vp), ('saved_bytes', i64), ('workspace', vp), ('workspace_bytes', i64),
                ('dropout_p', ctypes.c_float), ('dropout_seed', ctypes.c_uint64), ('act_saved', vp), ('act_saved_bytes', i64),
                ('n_ff', i32), ('ff_node', vp), ('ff_gold', vp), ('ff_w', vp), ('dhead_ff', vp), ('dhead_ff_elems', i64),
                ('ext_dlogits', vp), ('ext_datt', vp), ('ext_dhead_small', vp), ('ext_dhead_vec', vp), ('ext_dhead_ff', vp)]


class StairAdamSeg(ctypes.Structure):
    _fields_ = [('p', vp), ('g', vp), ('m', vp), ('v', vp), ('p2', vp), ('g2', vp), ('m2', vp), ('v2', vp),
                ('packed', vp), ('packed_ld', i64), ('packed_plane', i64),
                ('packed_t', vp), ('packed_t_ld', i64), ('packed_t_plane', i64), ('packed_perm', vp),
                ('rows', i32), ('cols', i32), ('kind', i32), ('nplanes', i32), ('perm_hh', i32), ('tile0', i32),
                ('bc1', ctypes.c_float), ('bc2', ctypes.c_float)]


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
        _lib.stair_itab_ints.restype = i64
        _lib.stair_nmn_workspace_bytes.restype = i64
        _lib.stair_last_launch_count.restype = i64
        _lib.stair_train_saved_bytes.restype = i64
        _lib.stair_train_workspace_bytes.restype = i64
        _lib.stair_train_act_bytes.restype = i64
    return _lib


def ptr(t):
    """Raw device/host pointer of a tensor (or None -> NULL)."""
    if t is None:
        return vp(0)
    if isinstance(t, torch.Tensor):
        return vp(t.data_ptr())
    return vp(int(t))


def addr(t):
    return 0 if t is None else t.data_ptr()


def stream_ptr(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return vp(s.cuda_stream)


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


def require_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise StairError('%s must be a CUDA tensor: stair_b200 runs only on sm_100a devices (no CPU fallback)' % what)


def gemm(A, Wt, bias=None, out=None, out_dtype=torch.bfloat16, act=ACT_NONE, row_scale=None, accumulate=False,
         M=None, N=None, K=None, nplanes=1, a_plane_rows=0, w_plane_rows=0, lda=None, ldw=None, ldc=None, stream=None):
    """out[M,N] = act(row_scale * (A[M,K] @ W[N,K]^T) + bias) via ``stair_gemm_bf16`` (tcgen05/TMA kernel)."""
    require_cuda(A, 'A')
    M = A.shape[0] if M is None else M
    K = A.shape[1] if K is None else K
    N = Wt.shape[0] if N is None else N
    lda = A.stride(0) if lda is None else lda
    ldw = Wt.stride(0) if ldw is None else ldw
    if out is None:
        out = torch.empty((M, N), device=A.device, dtype=out_dtype)
    ldc = out.stride(0) if ldc is None else ldc
    rc = lib().stair_gemm_bf16(ptr(A), i64(lda), i32(a_plane_rows), ptr(Wt), i64(ldw), i32(w_plane_rows), i32(nplanes),
                               ptr(bias), ptr(row_scale), ptr(out), i64(ldc), i32(dtype_code(out.dtype)),
                               i32(M), i32(N), i32(K), i32(act), i32(1 if accumulate else 0), stream_ptr(stream))
    check(rc, 'stair_gemm_bf16')
    return out


def gemm_gather(arena, slots, slot_rows, Wt, bias=None, out_dtype=torch.bfloat16, act=ACT_NONE, row_scale=None, stream=None):
    """A rows gathered from ``arena`` [slots, slot_rows, K] by ``slots`` (int32) — TMA 3-D gather, no staging copy."""
    require_cuda(arena, 'arena')
    n = slots.numel()
    K = arena.shape[-1]
    M, N = n * slot_rows, Wt.shape[0]
    out = torch.empty((M, N), device=arena.device, dtype=out_dtype)
    rc = lib().stair_gemm_bf16_gather(ptr(arena), i64(K), i64(arena.shape[0]), ptr(slots), i32(slot_rows), ptr(Wt), i64(Wt.stride(0)),
                                      ptr(bias), ptr(row_scale), ptr(out), i64(N), i32(dtype_code(out_dtype)), i32(M), i32(N),
                                      i32(K), i32(act), stream_ptr(stream))
    check(rc, 'stair_gemm_bf16_gather')
    return out
