"""ctypes binding of libdeephumor_sm100.so, generated from include/deephumor_b200.h.

The prototypes are parsed from the header so the Python side can never drift from the C ABI.  There is no
fallback: if the shared library is missing (or a symbol is), loading raises.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), 'include', 'deephumor_b200.h')
LIB_PATH = os.path.join(HERE, 'libdeephumor_sm100.so')

_SCALARS = {
    'int': ctypes.c_int, 'float': ctypes.c_float, 'long long': ctypes.c_longlong,
    'unsigned long long': ctypes.c_ulonglong, 'cudaStream_t': ctypes.c_void_p,
}


class BeamState(ctypes.Structure):
    """struct dh_beam_state (include/deephumor_b200.h)."""
    _fields_ = [('seq', ctypes.c_void_p), ('seq_ld', ctypes.c_longlong), ('val', ctypes.c_void_p),
                ('ended', ctypes.c_void_p), ('done', ctypes.c_void_p), ('final_len', ctypes.c_void_p),
                ('last_tok', ctypes.c_void_p), ('parent_state', ctypes.c_void_p), ('src', ctypes.c_void_p),
                ('S_alloc', ctypes.c_int)]


class VocabSparse(ctypes.Structure):
    """struct dh_vocab_sparse (include/deephumor_b200.h)."""
    _fields_ = [('thresh', ctypes.c_void_p), ('hitmap', ctypes.c_void_p), ('hit_ld', ctypes.c_longlong),
                ('logits', ctypes.c_void_p), ('ld', ctypes.c_longlong), ('n_cols', ctypes.c_int)]


class LstmOperands(ctypes.Structure):
    """struct dh_lstm_operands (include/deephumor_b200.h)."""
    _fields_ = [('table', ctypes.c_void_p), ('ldt', ctypes.c_longlong), ('n_tok_rows', ctypes.c_longlong),
                ('E', ctypes.c_int), ('H', ctypes.c_int), ('L', ctypes.c_int),
                ('hs', ctypes.c_void_p * 8), ('A', ctypes.c_void_p * 8), ('lda', ctypes.c_longlong * 8),
                ('in_off', ctypes.c_int * 8)]


class Resnet50Weights(ctypes.Structure):
    """struct dh_resnet50_weights (include/deephumor_b200.h)."""
    _fields_ = [('stem_w', ctypes.c_void_p), ('stem_b', ctypes.c_void_p),
                ('conv_w', (ctypes.c_void_p * 3) * 16), ('conv_b', (ctypes.c_void_p * 3) * 16),
                ('dual_w', ctypes.c_void_p * 4), ('dual_b', ctypes.c_void_p * 4),
                ('mean', ctypes.c_float * 3), ('std', ctypes.c_float * 3), ('dtype', ctypes.c_int)]


class XfmrLayer(ctypes.Structure):
    """struct dh_xfmr_layer (include/deephumor_b200.h)."""
    _fields_ = [('qkv_w', ctypes.c_void_p), ('qkv_b', ctypes.c_void_p),
                ('so_w', ctypes.c_void_p), ('so_b', ctypes.c_void_p), ('sln_g', ctypes.c_void_p), ('sln_b', ctypes.c_void_p),
                ('s_scale', ctypes.c_float),
                ('cq_w', ctypes.c_void_p), ('cq_b', ctypes.c_void_p),
                ('co_w', ctypes.c_void_p), ('co_b', ctypes.c_void_p), ('cln_g', ctypes.c_void_p), ('cln_b', ctypes.c_void_p),
                ('c_scale', ctypes.c_float),
                ('f1_w', ctypes.c_void_p), ('f1_b', ctypes.c_void_p), ('f2_w', ctypes.c_void_p), ('f2_b', ctypes.c_void_p),
                ('fln_g', ctypes.c_void_p), ('fln_b', ctypes.c_void_p)]


XFMR_MAX_LAYERS = 8


class XfmrWeights(ctypes.Structure):
    """struct dh_xfmr_weights (include/deephumor_b200.h)."""
    _fields_ = [('n_layers', ctypes.c_int), ('D', ctypes.c_int), ('n_heads', ctypes.c_int), ('pf', ctypes.c_int),
                ('cross', ctypes.c_int), ('dtype', ctypes.c_int), ('pad', ctypes.c_int), ('scale', ctypes.c_float),
                ('tok', ctypes.c_void_p), ('pos', ctypes.c_void_p), ('ld_tok', ctypes.c_longlong),
                ('layer', XfmrLayer * XFMR_MAX_LAYERS)]


class XfmrBuffers(ctypes.Structure):
    """struct dh_xfmr_buffers (include/deephumor_b200.h)."""
    _fields_ = [('x', ctypes.c_void_p), ('qb', ctypes.c_void_p), ('attn', ctypes.c_void_p), ('tmp', ctypes.c_void_p),
                ('h1', ctypes.c_void_p),
                ('Kc', ctypes.c_void_p * XFMR_MAX_LAYERS), ('Vc', ctypes.c_void_p * XFMR_MAX_LAYERS),
                ('xK', ctypes.c_void_p * XFMR_MAX_LAYERS), ('xV', ctypes.c_void_p * XFMR_MAX_LAYERS),
                ('enc_mask', ctypes.c_void_p), ('start', ctypes.c_void_p), ('ld_start', ctypes.c_longlong),
                ('seq', ctypes.c_void_p), ('seq_ld', ctypes.c_longlong), ('src', ctypes.c_void_p),
                ('slots', ctypes.c_int), ('S', ctypes.c_int)]


class Ctx:
    """Opaque dh_ctx* (path-level entries): created lazily, destroyed with the Python object."""

    def __init__(self):
        self.handle = ctypes.c_void_p()
        LIB.load()
        LIB.call('dh_ctx_create', ctypes.byref(self.handle), launches=0)
        self.keep = []                    # tensors whose device pointers the context holds

    def __del__(self):
        try:
            if self.handle:
                LIB.load().dh_ctx_destroy(self.handle)
        except Exception:
            pass


def parse_header(path=HEADER):
    """Returns {name: (restype, [argtypes], [argnames])} for every function the header declares."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    protos = {}
    for m in re.finditer(r'\b(int|const char\s*\*)\s+(dh_\w+)\s*\(([^;{}]*?)\)\s*;', text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), ' '.join(m.group(3).split())
        restype = ctypes.c_int if ret == 'int' else ctypes.c_char_p
        argtypes, argnames = [], []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                pm = re.match(r'^(.*?)(\w+)$', a)
                ty, nm = pm.group(1).strip(), pm.group(2)
                if '*' in ty:
                    argtypes.append(ctypes.c_void_p)
                else:
                    ty = ty.replace('const ', '').strip()
                    argtypes.append(_SCALARS[ty])
                argnames.append(nm)
        protos[name] = (restype, argtypes, argnames)
    return protos


class DeepHumorLibError(RuntimeError):
    pass


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()
        self.launches = 0          # number of kernel-launching C-ABI calls made (bench reports it)

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise DeepHumorLibError(
                f'{LIB_PATH} is missing: build it with `python -m deephumor_b200.build` '
                '(there is no CPU or PyTorch fallback for the caption path)')
        dll = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes, _) in self.protos.items():
            try:
                fn = getattr(dll, name)
            except AttributeError as e:
                raise DeepHumorLibError(f'{LIB_PATH} does not export {name} declared in {HEADER}') from e
            fn.restype, fn.argtypes = restype, argtypes
        if dll.dh_version() != self._header_version():
            raise DeepHumorLibError('libdeephumor_sm100.so is stale (DH_VERSION mismatch); rebuild')
        self._dll = dll
        return dll

    def _header_version(self):
        return int(re.search(r'#define DH_VERSION (\d+)', open(HEADER).read()).group(1))

    def call(self, name, *args, launches=1):
        """launches: kernels the entry launches (path-level entries launch many; pure host entries none)."""
        dll = self.load()
        rc = getattr(dll, name)(*args)
        self.launches += launches
        if rc != 0:
            msg = dll.dh_last_error().decode(errors='replace')
            kind = 'argument error' if rc < 0 else f'CUDA error {rc}'
            raise DeepHumorLibError(f'{name} failed ({kind}): {msg}')


LIB = _Lib()


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
