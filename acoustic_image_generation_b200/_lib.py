"""ctypes binding of libaig.so (include/aig.h) and the in-tree nvcc build.

The shared object lives next to its sources (``csrc/libaig.so``) so that it travels with
the repository snapshot to the GPU box.  There is no CPU implementation behind this
module: if the library cannot be loaded, or no B200 is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import shutil
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(CSRC, 'libaig.so')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

NVCC_FLAGS = ['-std=c++17', '-O3', '-lineinfo', '-gencode', 'arch=compute_100a,code=sm_100a',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-pthread', '-shared', '-ldl', '-lz', '-lpthread']
SOURCES = ['aig_api.cu', 'record_reader.cpp', 'host_copy.cpp']
DEPENDS = ['aig_api.cu', 'record_reader.cpp', 'host_copy.cpp', 'aig_common.cuh', 'mfcc_kernel.cuh', 'energy_kernel.cuh', 'heatmap_kernel.cuh', 'score_kernel.cuh', 'mask_packed_kernel.cuh', 'fused_kernel.cuh', 'frontend_kernel.cuh',
           'host_staging.h',
           'mel_program_ref.inc', 'mel_tables_ref.inc', os.path.join(INCLUDE, 'aig.h')]

AIG_OK = 0
ABI_VERSION = 2
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


BUILD_ID_MARKER = b'AIG_BUILD_ID='
_lock = threading.RLock()


def source_build_id():
    """Hash of everything the binary is made from (every file of DEPENDS, by content, plus the compiler flags).  build()
    stamps it into the library (aig_build_id()); load() refuses a library whose stamp differs, so a shipped libaig.so
    is always the committed source - modification times play no part."""
    h = hashlib.sha256()
    h.update(' '.join(NVCC_FLAGS + SOURCES).encode())
    for dep in DEPENDS:
        path = dep if os.path.isabs(dep) else os.path.join(CSRC, dep)
        h.update(os.path.basename(path).encode() + b'\0')
        with open(path, 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:24]


def binary_build_id(path=None):
    """The stamp inside a built library, read from the file (no dlopen); None when absent."""
    path = path or LIB_PATH
    try:
        with open(path, 'rb') as fh:
            blob = fh.read()
    except OSError:
        return None
    at = blob.find(BUILD_ID_MARKER)
    if at < 0:
        return None
    end = blob.find(b'\0', at)
    return blob[at + len(BUILD_ID_MARKER):end].decode('ascii', 'replace')


def needs_build():
    return binary_build_id() != source_build_id()


def build(force=False, verbose=False):
    """Compile csrc/*.cu into csrc/libaig.so for sm_100a (cross-compiles without a GPU)."""
    with _lock:
        if not force and not needs_build():
            return LIB_PATH
        cmd = [_nvcc()] + NVCC_FLAGS + ['-DAIG_BUILD_ID_STRING="%s"' % source_build_id(), '-I', INCLUDE,
                                        '-o', LIB_PATH + '.tmp'] + SOURCES
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
    'aig_build_id': (ctypes.c_char_p, []),
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
    'aig_acivw_batch': (_int, [_p, _p, _p, _i64, _int, _p, _int, _p, _p, _p, _p, _p, _p, _p, _p]),
    'aig_resize_mask': (_int, [_p, _p, _i64, _int, _int, _p]),
    'aig_mfcc_energy': (_int, [_p, _p, _i64, _int, _int, _p, _p, _p, _p]),
    'aig_mfcc_energy_heatmap': (_int, [_p, _p, _i64, _int, _int, _p, _p, _p, _p, _p, _int, _int]),
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
    'aig_comm_init_all': (_int, [_p, _int]),
    'aig_group_allreduce_counts': (_int, [_p, _p, _int, _int]),
    'aig_auc': (_int, [_p, _p, _int, _p]),
    'aig_launch_count': (_i64, [_p]),
    'aig_set_option': (_int, [_p, ctypes.c_char_p, _i64]),
    'aig_profile_read': (_int, [_p, _p, _p]),
    'aig_selftest': (_int, [_p, _int, _p]),
}

_lib = None


def load(build_if_missing=True):
    """dlopen csrc/libaig.so and attach the prototypes.

    The library must carry the build id of the sources next to it (source_build_id): a missing or stale binary is
    rebuilt when ``build_if_missing`` (needs nvcc), otherwise - and if the stamp still differs afterwards - this raises.
    Thread-safe: the loaders' worker threads may race here."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        want = source_build_id()
        if binary_build_id() != want:
            if not build_if_missing:
                raise RuntimeError('%s is missing or was built from other sources (build id %s, sources %s); run '
                                   '__graft_entry__.build()' % (LIB_PATH, binary_build_id(), want))
            build(force=True)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the library does not export the symbol
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.aig_abi_version() != ABI_VERSION:
            raise RuntimeError('libaig.so ABI version %d does not match this binding (%d)' % (lib.aig_abi_version(), ABI_VERSION))
        have = (lib.aig_build_id() or b'').decode()
        if have != want:
            raise RuntimeError('libaig.so build id %s does not match the sources (%s)' % (have, want))
        _lib = lib
        return lib
