"""ctypes binding of libvog_b200.so (C ABI declared in include/vog_b200.h).

The library is built in-tree by ``build()`` (called from ``__graft_entry__.build()``) so that the
``.so`` travels with the repo snapshot to the GPU box.  There is NO fallback: if the library is
missing or a call fails the caller gets an exception.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
SO_PATH = os.path.join(_HERE, 'libvog_b200.so')
SOURCES = ['vog_abi.cu', 'fp32_path.cu', 'tc_gemm.cu', 'tc_attn.cu', 'lstm_rec.cu', 'fused_glue.cu', 'loss_fwd.cu',
           'relayout.cu', 'optim.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

_lock = threading.Lock()
_lib = None

c_int, c_float, c_void_p, c_i64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64
P = c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPE)
_SIGNATURES = {
    'vog_last_error': [],
    'vog_abi_version': [],
    'vog_launch_count': [],
    'vog_device_is_sm100': [],
    'vog_sgemm_nt': [P, c_int, P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_int, P],
    'vog_attn_fwd_f32': [P, P, P, c_int, P, c_int, c_int, c_int, c_int, P, P, c_float, c_int, P,
                         c_int, P, P, P],
    'vog_add_layernorm': [P, c_int, P, c_int, P, P, P, c_int, P, c_int, c_int, c_int, c_int,
                          c_float, P],
    'vog_pe_project': [P, c_int, P, P, c_int, c_int, c_float, c_float, c_float, c_float, P],
    'vog_select_fwd': [P, P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P],
    'vog_select_sep_fwd': [P, P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P],
    'vog_sep_fin_scores': [P, P, P, P, P, P, P, c_int, c_int, c_int, P],
    'vog_adam_step': [P, P, P, P, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, c_i64,
                      ctypes.c_double, P],
    'vog_verb_loss_fwd': [P, P, P, c_int, c_int, c_float, P, P],
    'vog_concat_videos': [P, c_int, P, c_int, P, c_int, P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P],
    'vog_cast_lp': [P, c_i64, P, c_i64, c_i64, c_int, c_int, P],
    'vog_tc_gemm_workspace_bytes': [c_int, c_int, c_int, c_int, c_int],
    'vog_tc_gemm': [P, c_i64, P, c_i64, c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_i64,
                    P, c_i64, P, c_i64, c_int, c_int, P, c_i64, P],
    'vog_build_xmul': [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P],
    'vog_lin2_tail': [P, c_int, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P],
    'vog_lang_embed': [P, c_int, P, c_int, P, c_int, c_i64, c_int, P, c_int, P],
    'vog_lang_gather': [P, c_int, P, c_int, c_int, c_int, P, c_int, P],
    'vog_mask_rows': [P, P, c_int, c_int, P, P, c_int, P],
    'vog_loss_workspace_bytes': [c_int, c_int, c_int],
    'vog_loss_bwd': [P, P, P, P, P, c_int, c_int, c_int, P],
    'vog_loss_fwd': [P, P, c_int, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                     c_float, P, P, P, P],
    'vog_lstm_set_max_ctas': [c_int],
    'vog_set_reserved_sms': [c_int],
    'vog_lstm_workspace_bytes': [c_int, c_int],
    'vog_debug_lstm_force_streaming': [c_int],
    'vog_debug_gemm_trace': [P],
    'vog_debug_lstm_trace': [P],
    'vog_debug_lstm_exchange': [c_int],
    'vog_debug_attn_prof': [P],
    'vog_debug_attn_impl': [c_int],
    'vog_debug_attn_cluster': [c_int],
    'vog_lstm_layer_fwd': [P, c_i64, P, P, c_int, c_int, c_int, P, c_i64, c_int, P, P],
    'vog_tc_attn_workspace_bytes': [c_int, c_int, c_int],
    'vog_tc_attn_fwd': [P, P, P, c_int, c_int, c_int, c_int, P, c_float, c_int, P, c_int, P, P, P,
                        c_i64, c_int, P, c_i64, P],
    'vog_tc_gemm_qkv_factored': [P, c_i64, P, c_i64, c_int, c_int, c_int, c_int, c_int, P, c_i64, c_int, c_int,
                                 c_int, P, P, P, P],
    'vog_tc_gemm_lin2': [P, c_i64, P, c_i64, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, c_int, c_int, c_int,
                         c_int, c_int, c_int, c_int, c_int, P],
    'vog_tc_gemm_gres': [P, c_i64, P, c_i64, c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_i64, P, c_i64,
                         c_int, c_int, c_int, c_int, P, c_i64, P, c_i64, c_int, P],
    'vog_tc_gemm_qkv': [P, c_i64, P, c_i64, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P],
}
_RESTYPE = {'vog_last_error': ctypes.c_char_p, 'vog_launch_count': ctypes.c_longlong,
            'vog_tc_gemm_workspace_bytes': ctypes.c_int64, 'vog_lstm_workspace_bytes': ctypes.c_int64,
            'vog_tc_attn_workspace_bytes': ctypes.c_int64, 'vog_loss_workspace_bytes': ctypes.c_int64}


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    deps.append(os.path.join(os.path.dirname(_HERE), 'include', 'vog_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> vognet_pytorch_b200/libvog_b200.so"""
    if not force and not needs_build():
        return SO_PATH
    extra = os.environ.get('VOG_NVCC_EXTRA', '').split()       # e.g. -DVOG_ATTN_PROFILE
    cmd = ['nvcc'] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + ['-o', SO_PATH] + sources()
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO_PATH


def lib():
    """The loaded library; raises if it has not been built (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(SO_PATH):
                raise RuntimeError(
                    f'{SO_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                    '(nvcc, sm_100a).  vognet_pytorch_b200 has no fallback path.')
            l = ctypes.CDLL(SO_PATH)
            for name, argtypes in _SIGNATURES.items():
                fn = getattr(l, name)          # AttributeError if the symbol is not exported
                fn.argtypes = argtypes
                fn.restype = _RESTYPE.get(name, c_int)
            _lib = l
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        msg = lib().vog_last_error()
        raise RuntimeError(f'{what} failed: {msg.decode() if msg else "unknown error"}')
