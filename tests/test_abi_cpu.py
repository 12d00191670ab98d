"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, exports every
symbol include/aig.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import acoustic_image_generation_b200 as aig
from acoustic_image_generation_b200 import _lib, metrics_io, tables
from oracle import acoustic_oracle as oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    _lib.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'aig.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(aig_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 17
    for name in names:
        assert hasattr(lib, name), 'libaig.so does not export %s' % name
    assert sorted(_lib.SIGNATURES) == names, 'ctypes binding and header disagree'


def test_abi_version(lib):
    assert lib.aig_abi_version() == _lib.ABI_VERSION == 2
    header = open(os.path.join(ROOT, 'include', 'aig.h')).read()
    assert '#define AIG_ABI_VERSION %d' % _lib.ABI_VERSION in header


def test_binary_carries_the_build_id_of_the_sources(lib, tmp_path, monkeypatch):
    """The shipped libaig.so is tied to the sources next to it by content hash, not by modification time: the stamp in
    the file, the stamp the loaded library reports and the hash of DEPENDS + compiler flags must all agree, and a
    library with another stamp is refused."""
    want = _lib.source_build_id()
    assert _lib.binary_build_id() == want and not _lib.needs_build()
    assert lib.aig_build_id().decode() == want
    # any change to a source file changes the id ...
    blob = open(_lib.LIB_PATH, 'rb').read()
    fake = tmp_path / 'libaig.so'
    fake.write_bytes(blob.replace(_lib.BUILD_ID_MARKER + want.encode(), _lib.BUILD_ID_MARKER + b'0' * len(want)))
    assert _lib.binary_build_id(str(fake)) == '0' * len(want)
    # ... and load() refuses a stale binary when it may not rebuild it
    monkeypatch.setattr(_lib, 'LIB_PATH', str(fake))
    monkeypatch.setattr(_lib, '_lib', None)
    with pytest.raises(RuntimeError, match='built from other sources'):
        _lib.load(build_if_missing=False)


def test_shared_object_is_in_tree_and_sm100a():
    assert os.path.dirname(_lib.LIB_PATH).endswith(os.path.join('acoustic_image_generation_b200', 'csrc'))
    assert os.path.exists(_lib.LIB_PATH)
    assert 'arch=compute_100a,code=sm_100a' in _lib.NVCC_FLAGS


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    h = ctypes.c_void_p()
    assert lib.aig_create(0, 0, ctypes.byref(h)) == -2
    assert b'no CUDA device' in lib.aig_last_error(None)
    with pytest.raises(aig.AigError):
        aig.AcousticPath(0)
    with pytest.raises(aig.AigError):
        aig.find_logen(np.zeros((36, 48, 12), np.float32))


def test_auc_host_entry_matches_oracle(lib, golden):
    g = golden('auc')
    a = golden('acivw_iou'); f = golden('flickr_ciou')
    thr = list(oracle.REFERENCE_THRESHOLDS)
    assert aig.auc(thr, aig.success_rates(a['pos11'], a['num'])) == float(g['acivw11'])
    assert aig.auc(thr, aig.success_rates(f['pos11'], f['num'])) == float(g['flickr11'])
    t101 = np.linspace(0, 1, 101)
    v101 = aig.success_rates(f['pos101'], f['num'])
    assert aig.auc(t101, v101) == oracle.auc(t101, v101)
    rng = np.random.default_rng(0)
    for k in (2, 3, 8, 9, 11, 101, 130, 300, 1000):
        t = np.linspace(0, 1, k)
        v = np.sort(rng.random(k))[::-1]
        assert aig.auc(t, v) == oracle.auc(t, v), k
    with pytest.raises(aig.AigError):
        aig.auc([0.0, 0.5, 0.2], [1, 1, 1])


def test_product_tables_equal_oracle_tables():
    bank, dct, lifter, mfnorm = tables.reference_tables()
    ob, od, ol, om = oracle.reference_tables()
    assert np.array_equal(bank, ob) and np.array_equal(dct, od) and np.array_equal(lifter, ol) and mfnorm == om
    assert np.array_equal(aig.createfilters(256, 20, 300, 4000, 8000), oracle.createfilters(256, 20, 300, 4000, 8000))


def test_generated_mel_program_matches_tables():
    """The straight-line program compiled into the kernel is the float32 image of the oracle's bank."""
    bank, dct, lifter, mfnorm = oracle.reference_tables()
    text = open(os.path.join(ROOT, 'acoustic_image_generation_b200', 'csrc', 'mel_program_ref.inc')).read()
    comp = {'x': 0, 'y': 1, 'z': 2, 'w': 3}
    got = np.zeros((512, 24), np.float32)
    folded = np.zeros((24, 12), np.float32)
    slab = None
    for line in text.splitlines():
        m = re.match(r'MEL_SLAB_BEGIN\((\d+)\)', line)
        if m:
            slab = int(m.group(1)); continue
        m = re.match(r'MEL_BIN1\((\d+), (\w), (\d+), \d, ([^)]+)\)', line)
        if m:
            k = 32 * slab + 4 * int(m.group(1)) + comp[m.group(2)]
            got[k, int(m.group(3))] = float.fromhex(m.group(4).rstrip('f')); continue
        m = re.match(r'MEL_BIN2\((\d+), (\w), (\d+), \d, ([^,]+), (\d+), \d, ([^)]+)\)', line)
        if m:
            k = 32 * slab + 4 * int(m.group(1)) + comp[m.group(2)]
            got[k, int(m.group(3))] = float.fromhex(m.group(4).rstrip('f'))
            got[k, int(m.group(5))] = float.fromhex(m.group(6).rstrip('f')); continue
        m = re.match(r'MEL_DONE\((\d+), (.+)\)', line)
        if m:
            folded[int(m.group(1))] = [float.fromhex(v.strip().rstrip('f')) if 'x' in v else float(v.strip().rstrip('f'))
                                       for v in m.group(2).split(',')]
    assert np.array_equal(got, bank.astype(np.float32))
    assert np.array_equal(folded, (dct * mfnorm * lifter[None, :]).astype(np.float32))


def test_metric_files_round_trip(tmp_path):
    for thr, pos in zip(oracle.REFERENCE_THRESHOLDS, range(11)):
        p = metrics_io.write_accuracy_file(str(tmp_path), thr, pos, 10)
        assert os.path.basename(p) == 'intersection_{}_accuracy.txt'.format(thr * 1.0)
        assert open(p).read() == 'iou {:6f}'.format(pos / 10)
        assert metrics_io.read_accuracy_file(str(tmp_path), thr) == round(pos / 10, 6)
    p = metrics_io.write_area_file(str(tmp_path), 0.4321987)
    assert open(p).read() == 'area 0.432199'


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is test infrastructure: nothing in the package, tools/ or examples/ may import it; bench.py and
    __graft_entry__.py may (CPU baseline legs and smoke())."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pattern = re.compile(r'^\s*(from\s+oracle\b|import\s+oracle\b)', re.M)
    offenders = []
    for sub in ('acoustic_image_generation_b200', 'tools', 'examples'):
        for base, _, files in os.walk(os.path.join(root, sub)):
            for name in files:
                if name.endswith('.py') and pattern.search(open(os.path.join(base, name)).read()):
                    offenders.append(os.path.relpath(os.path.join(base, name), root))
    assert not offenders, offenders


def test_streaming_slot_fill_copies_exactly(lib):
    """host_copy.cpp (the non-temporal fill of the pinned staging slots) is plain host code: every source / destination
    alignment, sizes from 0 to 2 MiB, nothing written outside the destination range."""
    raw = ctypes.CDLL(_lib.LIB_PATH)
    raw.aig_host_copy_streaming.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    raw.aig_host_copy_streaming.restype = None
    assert raw.aig_host_copy_streaming_supported() in (0, 1)
    rng = np.random.default_rng(7)
    src = rng.integers(0, 256, size=(2 << 20) + 256, dtype=np.uint8)
    sizes = [0, 1, 31, 32, 4095, 4096, 4097, 4096 + 127, 4096 + 128, 65536 + 33] + [int(v) for v in rng.integers(0, 2 << 20, 40)]
    for k, n in enumerate(sizes):
        so, do = (k * 7) % 64, (k * 13) % 64
        dst = np.full(n + do + 96, 0xA5, np.uint8)
        raw.aig_host_copy_streaming(dst.ctypes.data + do, src.ctypes.data + so, n)
        assert np.array_equal(dst[do:do + n], src[so:so + n]), (n, so, do)
        assert (dst[:do] == 0xA5).all() and (dst[do + n:] == 0xA5).all(), (n, so, do)
