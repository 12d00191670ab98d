"""ctypes binding of libaig.so (include/aig.h) and the in-tree nvcc build.

The shared object lives next to its sources (``csrc/libaig.so``) so that it travels with
the repository snapshot to the GPU box.  There is no CPU implementation behind this
module: if the library cannot be loaded, or no B200 is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(CSRC, 'libaig.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_FLAGS = ['-std=c++17', '-O3', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-pthread', '-shared', '-ldl', '-lz', '-lpthread']
SOURCES = ['aig_api.cu', 'record_reader.cpp']
DEPENDS = ['aig_api.cu', 'record_reader.cpp', 'aig_common.cuh', 'mfcc_kernel.cuh', 'energy_kernel.cuh', 'score_kernel.cuh', 'fused_kernel.cuh', 'frontend_kernel.cuh', 'host_staging.h',
           'mel_program_ref.inc', 'mel_tables_ref.inc', os.path.join(INCLUDE, 'aig.h')]

AIG_OK = 0
ERROR_NAMES = {-1: 'AIG_ERR_ARGUMENT', -2: 'AIG_ERR_NO_DEVICE', -3: 'AIG_ERR_TABLES', -4: 'AIG_ERR_ALLOC'}


class AigError(RuntimeError):
    """A libaig call returned a negative status."""

    def __init__(self, code, message):
        self.code = code
        name = ERROR_NAMES.get(code, 'CUDA error %d' % (-(code + 1000)) if code <= -1000 else 'error')
        super().__init__('libaig: %s (%d): %s' % (name, code, message))


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: cannot build libaig.so')


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    for dep in DEPENDS:
        path = dep if os.path.isabs(dep) else os.path.join(CSRC, dep)
        if os.path.getmtime(path) > built:
            return True
    return False


def build(force=False, verbose=False):
    """Compile csrc/*.cu into csrc/libaig.so for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ['-I', INCLUDE, '-o', LIB_PATH + '.tmp'] + SOURCES
    if verbose:
        cmd += ['-Xptxas', '-v']
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (' '.join(cmd), proc.stderr[-4000:]))
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    if verbose:
        print(proc.stderr)
    return LIB_PATH


_p = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double

# name -> (restype, argtypes); must list every symbol include/aig.h declares
SIGNATURES = {
    'aig_abi_version': (_int, []),
    'aig_create': (_int, [_int, ctypes.c_uint64, ctypes.POINTER(_p)]),
    'aig_destroy': (_int, [_p]),
    'aig_last_error': (ctypes.c_char_p, [_p]),
    'aig_synchronize': (_int, [_p]),
    'aig_set_tables': (_int, [_p, _p, _int, _int, _p, _int, _p, _dbl]),
    'aig_tables_are_reference': (_int, [_p]),
    'aig_mfcc': (_int, [_p, _p, _i64, _p, _int, _int]),
    'aig_normalize_images': (_int, [_p, _p, _i64, _p]),
    'aig_energy': (_int, [_p, _p, _i64, _int, _p, _p, _p, _p]),
    'aig_heatmap': (_int, [_p, _p, _i64, _int, _int, _p]),
    'aig_energy_heatmap': (_int, [_p, _p, _i64, _int, _p, _p, _p, _int, _int]),
    'aig_resize_mask': (_int, [_p, _p, _i64, _int, _int, _p]),
    'aig_mfcc_energy': (_int, [_p, _p, _i64, _int, _int, _p, _p, _p, _p]),
    'aig_iou_sweep': (_int, [_p, _p, _p, _i64, _p, _int, _p, _p, _p, _p]),
    'aig_iou_sweep_clips': (_int, [_p, _p, _p, _i64, _i64, _p, _int, _p, _p, _p]),
    'aig_ciou_sweep': (_int, [_p, _p, _p, _p, _p, _p, _i64, _int, _int, _p, _int, _p, _p, _p, _p]),
    'aig_power_spectrum': (_int, [_p, _p, _int, _i64, _p, _p]),
    'aig_audio_mfcc': (_int, [_p, _p, _int, _i64, _p, _p]),
    'aig_filtfilt': (_int, [_p, _p, _int, _i64, _int, _p, _p, _p, _int, _p]),
    'aig_normalize_mfcc': (_int, [_p, _p, _i64, _p]),
    'aig_tile_mfcc': (_int, [_p, _p, _i64, _int, _p]),
    'aig_split_triplets': (_int, [_p, _p, _i64, _p]),
    'aig_triplet_mse': (_int, [_p, _p, _p, _i64, _p]),
    'aig_overlay': (_int, [_p, _p, _p, _i64, _int, _int, ctypes.c_float, _p, _p]),
    'aig_crc32c': (ctypes.c_uint32, [_p, ctypes.c_size_t, _int]),
    'aig_records_open': (_int, [ctypes.c_char_p, ctypes.POINTER(_p)]),
    'aig_records_close': (_int, [_p]),
    'aig_records_count': (_i64, [_p]),
    'aig_record_context_int64': (_int, [_p, _int, ctypes.c_char_p, _p, _int, ctypes.POINTER(_int)]),
    'aig_record_sequence_size': (_int, [_p, _int, ctypes.c_char_p, ctypes.POINTER(_i64), ctypes.POINTER(_i64)]),
    'aig_record_sequence_read': (_int, [_p, _int, ctypes.c_char_p, _p, _i64]),
    'aig_records_last_error': (ctypes.c_char_p, []),
    'aig_comm_unique_id': (_int, [_p]),
    'aig_comm_init': (_int, [_p, _p, _int, _int]),
    'aig_allreduce_counts': (_int, [_p, _p, _int]),
    'aig_comm_destroy': (_int, [_p]),
    'aig_auc': (_int, [_p, _p, _int, _p]),
    'aig_launch_count': (_i64, [_p]),
    'aig_set_option': (_int, [_p, ctypes.c_char_p, _i64]),
    'aig_profile_read': (_int, [_p, _p, _p]),
    'aig_selftest': (_int, [_p, _int, _p]),
}

_lib = None


def load(build_if_missing=True):
    """dlopen csrc/libaig.so and attach the prototypes.  Raises if it is missing and cannot be built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError('%s is missing; run __graft_entry__.build()' % LIB_PATH)
        build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export the symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.aig_abi_version() != 1:
        raise RuntimeError('libaig.so ABI version %d does not match this binding (1)' % lib.aig_abi_version())
    _lib = lib
    return lib
