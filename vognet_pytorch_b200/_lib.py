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
           'relayout.cu', 'optim.cu', 'train_f32.cu', 'tc_gemm_tn.cu', 'tc_attn_bwd.cu', 'lstm_bwd.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

_lock = threading.Lock()
_lib = None

c_int, c_float, c_void_p, c_i64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64
P = c_void_p

HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'vog_b200.h')
_CTYPE = {'int': c_int, 'int64_t': c_i64, 'uint64_t': ctypes.c_uint64, 'float': c_float, 'double': ctypes.c_double, 'void': None,
          'long long': ctypes.c_longlong}


def _parse_header(path=HEADER):
    """The C ABI is declared ONCE, in include/vog_b200.h; the ctypes prototypes are derived from those declarations
    (pointer -> void*, int / int64_t / float / double by value), so binding and header cannot drift apart."""
    import re
    hdr = re.sub(r'/\*.*?\*/', ' ', open(path).read(), flags=re.S)
    sigs, res = {}, {}
    for ret, name, args in re.findall(r'\b(const char\s*\*|long long|int64_t|int|void)\s+(vog_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;',
                                      hdr, flags=re.S):
        args = ' '.join(args.split())
        params = [] if args in ('', 'void') else [a.strip() for a in args.split(',')]
        at = []
        for prm in params:
            if '*' in prm:
                at.append(c_void_p)
            else:
                base = prm.replace('const ', '').replace('unsigned ', '')
                base = 'long long' if base.startswith('long long') else base.split()[0]
                at.append(_CTYPE[base])
        sigs[name] = at
        ret = ' '.join(ret.split())
        res[name] = ctypes.c_char_p if '*' in ret else (_CTYPE[ret] if ret != 'int' else c_int)
    return sigs, res


# name -> argtypes / restype
_SIGNATURES, _RESTYPE = _parse_header()


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
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> vognet_pytorch_b200/libvog_b200.so.  One object per source
    (csrc/_obj/, rebuilt only when the source or a header is newer), compiled in parallel, then linked."""
    if not force and not needs_build():
        return SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    extra = os.environ.get('VOG_NVCC_EXTRA', '').split()       # e.g. -DVOG_ATTN_PROFILE
    objdir = os.path.join(CSRC, '_obj')
    os.makedirs(objdir, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    hdrs.append(HEADER)
    t_hdr = max(os.path.getmtime(h) for h in hdrs if os.path.exists(h))
    tag = os.path.join(objdir, 'flags.txt')
    flags_now = ' '.join(NVCC_FLAGS + extra)
    if force or not os.path.exists(tag) or open(tag).read() != flags_now:
        for f in os.listdir(objdir):
            os.remove(os.path.join(objdir, f))
        open(tag, 'w').write(flags_now)
    cflags = [f for f in NVCC_FLAGS if f != '-shared']

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + '.o')
        if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), t_hdr):
            return obj, ''
        cmd = ['nvcc'] + cflags + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, src]
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n' + r.stdout + r.stderr)
        return obj, r.stderr
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        res = list(ex.map(compile_one, sources()))
    if verbose:
        print(''.join(e for _, e in res))
    cmd = ['nvcc', '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', SO_PATH] + [o for o, _ in res]
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
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
            # VOG_B200_SO: an instrumented build of the SAME sources (profiles/: -DVOG_LSTM_TRACE, -DVOG_ATTN_PROFILE)
            l = ctypes.CDLL(os.environ.get('VOG_B200_SO') or SO_PATH)
            for name, argtypes in _SIGNATURES.items():
                fn = getattr(l, name)          # AttributeError if the symbol is not exported
                fn.argtypes = argtypes
                fn.restype = _RESTYPE[name]
            if os.environ.get('VOG_PDL', '0') == '1':       # programmatic dependent launches everywhere (default: mdl_vog decides)
                l.vog_debug_pdl(1)
            if os.environ.get('VOG_LSTM_XMODE'):            # A/B: h_t exchange protocol of the recurrence kernel
                l.vog_debug_lstm_exchange(int(os.environ['VOG_LSTM_XMODE'], 0))
            _lib = l
    return _lib


def exported_symbols():
    return list(_SIGNATURES)


def check(rc, what):
    if rc != 0:
        msg = lib().vog_last_error()
        raise RuntimeError(f'{what} failed: {msg.decode() if msg else "unknown error"}')
