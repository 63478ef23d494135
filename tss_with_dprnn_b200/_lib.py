"""ctypes binding of libdprnn_b200.so.

The prototypes are read from ``include/dprnn_b200.h`` (the single source of truth for the C ABI), so
the Python side cannot drift from the header.  There is no fallback: if the library is missing or a
call fails, this raises.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, '..', 'include', 'dprnn_b200.h')
LIB_PATH = os.path.join(_HERE, 'csrc', 'libdprnn_b200.so')

_CTYPES = {
    'int': ctypes.c_int, 'long': ctypes.c_long, 'float': ctypes.c_float, 'size_t': ctypes.c_size_t,
    'double': ctypes.c_double,
}


def _ctype(decl: str):
    decl = decl.replace('const', '').strip()
    if '*' in decl:
        return ctypes.c_char_p if decl.replace(' ', '') == 'char*' else ctypes.c_void_p
    return _CTYPES[decl.split()[0]]


def parse_header(path: str = HEADER):
    """-> {name: (restype, [argtypes])} for every function the header declares."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    text = re.sub(r'//[^\n]*', '', text)
    protos = {}
    for m in re.finditer(r'((?:const\s+)?(?:char|int|long|float|void|size_t)\s*\*?)\s*(dprnn_\w+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != 'void':
            for a in args.split(','):
                a = a.strip()
                a = re.sub(r'\b\w+$', '', a).strip() if not a.endswith('*') else a   # drop the parameter name
                argtypes.append(_ctype(a))
        protos[name] = (_ctype(ret), argtypes)
    return protos


#: kernels launched per C-ABI call (everything else launches exactly one)
# kernels per entry point where that is not one (the launch count bench.py reports)
_LAUNCHES = {'dprnn_utt_stats': 2, 'dprnn_att_rowscale': 3, 'dprnn_gemm_atb_dual': 2, 'dprnn_gemm_atb_tc': 2,
             'dprnn_gemm_atb_tc_colsum': 2, 'dprnn_groupnorm_bwd': 3, 'dprnn_groupnorm_bwd_h16': 3, 'dprnn_col_sum': 2, 'dprnn_prelu_bwd': 2}


class _Lib:
    def __init__(self):
        self.launches = 0          # running count of kernel launches issued through call()
        self.timing = None         # optional {entry-point name: [(start_event, end_event), ...]}
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} is missing. Build it with `python -m tss_with_dprnn_b200.build` (needs nvcc). '
                'There is no CPU or PyTorch fallback for the separation path.')
        self.cdll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (ret, args) in self.protos.items():
            fn = getattr(self.cdll, name)          # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = ret, args

    def last_error(self) -> str:
        return self.cdll.dprnn_last_error().decode()

    def build_info(self) -> str:
        return self.cdll.dprnn_build_info().decode()

    def call(self, name: str, *args):
        """Invoke an int-returning entry point; tensors are passed as device pointers, None as NULL."""
        if name not in self.protos:      # an undeclared symbol would be called with ctypes' default (32-bit int) arguments
            raise AttributeError(f'{name} is not declared in include/dprnn_b200.h')
        conv = []
        for a in args:
            if isinstance(a, torch.Tensor):
                conv.append(a.data_ptr())
            else:
                conv.append(a)
        timed = self.timing is not None and name in self.timing
        if timed:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        rc = getattr(self.cdll, name)(*conv)
        if rc != 0:
            raise RuntimeError(f'{name} failed (rc={rc}): {self.last_error()}')
        if timed:
            ev[1].record()
            self.timing[name].append(ev)
        self.launches += _LAUNCHES.get(name, 1)
        if name == 'dprnn_batchnorm_affine' and conv[7]:
            self.launches += 1     # training mode adds the batch-statistics kernel

    def query(self, name: str, *args):
        if name not in self.protos:
            raise AttributeError(f'{name} is not declared in include/dprnn_b200.h')
        return getattr(self.cdll, name)(*args)


_lib = None


def lib() -> _Lib:
    global _lib
    if _lib is None:
        _lib = _Lib()
    return _lib
