"""The reference's on-disk format: GZIP TFRecord files of tf.train.SequenceExample.

``RecordFile`` reads them through libaig's dependency-free C++ reader (include/aig.h, csrc/record_reader.cpp) - no
TensorFlow - and ``parse_acoustic_example`` / ``parse_flickr_example`` return what the reference's ``_parse_sequence``
produces (dataloader/outdoor_data_mfcc.py:260-344, dataloader/frames.py:246-341).  ``write_sequence_examples`` writes the
same format (convert_data.py:247-279) so synthetic ACIVW-shaped data sets can be produced without TensorFlow.
"""
from __future__ import annotations

import ctypes
import gzip
import struct

import numpy as np

from . import _lib
from ._lib import AigError


class RecordFile:
    """A TFRecord file (GZIP or plain) opened with aig_records_open; every record's CRC-32C is verified on open."""

    def __init__(self, path):
        self._lib = _lib.load()
        self._r = ctypes.c_void_p()
        code = self._lib.aig_records_open(str(path).encode(), ctypes.byref(self._r))
        if code != 0:
            raise AigError(code, (self._lib.aig_records_last_error() or b'').decode())

    def close(self):
        if getattr(self, '_r', None) is not None and self._r.value:
            self._lib.aig_records_close(self._r)
            self._r = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return int(self._lib.aig_records_count(self._r))

    def _check(self, code):
        if code != 0:
            raise AigError(code, (self._lib.aig_records_last_error() or b'').decode())

    def context(self, record, key, capacity=16):
        """int64 context feature as a NumPy array (scalars come back with shape [1])."""
        buf = (ctypes.c_int64 * capacity)()
        count = ctypes.c_int()
        self._check(self._lib.aig_record_context_int64(self._r, record, key.encode(), buf, capacity, ctypes.byref(count)))
        if count.value > capacity:
            return self.context(record, key, count.value)
        return np.array(buf[:count.value], dtype=np.int64)

    def sequence(self, record, key, dtype):
        """The feature list's byte strings decoded like tf.decode_raw(..., dtype): array [steps, elements per step]."""
        steps, nbytes = ctypes.c_int64(), ctypes.c_int64()
        self._check(self._lib.aig_record_sequence_size(self._r, record, key.encode(), ctypes.byref(steps), ctypes.byref(nbytes)))
        out = np.empty(nbytes.value, dtype=np.uint8)
        self._check(self._lib.aig_record_sequence_read(self._r, record, key.encode(), out.ctypes.data, out.nbytes))
        arr = out.view(np.dtype(dtype))
        return arr.reshape(steps.value, -1) if steps.value else arr.reshape(0, 0)


def _scalar(records, index, key):
    values = records.context(index, key)
    if values.size == 0:
        raise AigError(-1, 'context feature holds no value: ' + key)
    return int(values[0])


def parse_acoustic_example(records, index, flip=True):
    """_parse_sequence of the ACIVW loader (outdoor_data_mfcc.py:260-344) for the modalities present in the record:
    {'classes', 'location', 'audio_images' [T,H,W,D] float32 (flipped left-right and up-down when ``flip``, :314-315),
    'audio_samples' [T*mics, samples] int32, 'video_images' [T,H,W,3] uint8}."""
    out = {'classes': _scalar(records, index, 'classes'), 'location': _scalar(records, index, 'location')}
    try:
        h, w, d = (_scalar(records, index, 'audio_image/' + k) for k in ('height', 'width', 'depth'))
        img = records.sequence(index, 'audio/image', np.float32).reshape(-1, h, w, d)
        out['audio_images'] = np.ascontiguousarray(img[:, ::-1, ::-1, :]) if flip else img
    except AigError:
        pass
    try:
        samples = _scalar(records, index, 'audio_data/samples')
        out['audio_samples'] = records.sequence(index, 'audio/data', np.int32).reshape(-1, samples)
    except AigError:
        pass
    try:
        h, w, d = (_scalar(records, index, 'video/' + k) for k in ('height', 'width', 'depth'))
        out['video_images'] = records.sequence(index, 'video/image', np.uint8).reshape(-1, h, w, d)
    except AigError:
        pass
    return out


def parse_flickr_example(records, index):
    """The FlickrSoundNet variant (dataloader/frames.py:246-341): no flips (:311-312 are commented out there) and the
    annotator boxes 'xmin', 'xmax', 'ymin', 'ymax' as int32 [3] (frames.py:290-299)."""
    out = parse_acoustic_example(records, index, flip=False)
    for key in ('xmin', 'xmax', 'ymin', 'ymax'):
        try:
            out[key] = records.context(index, key).astype(np.int32)
        except AigError:
            pass
    return out


def iterate_examples(paths, parse=parse_acoustic_example, workers=4, prefetch=8, **parse_kwargs):
    """Every example of every file in ``paths``, in order, read and parsed ``workers`` files at a time with up to
    ``prefetch`` files ahead of the consumer - the role of ``TFRecordDataset(...).map(_parse_sequence,
    num_parallel_calls=4).prefetch(...)`` in the reference's loaders (dataloader/outdoor_data_mfcc.py:62-82).  Inflating,
    CRC checking and the protobuf walk run inside libaig with the GIL released, so the threads overlap."""
    from concurrent.futures import ThreadPoolExecutor

    def load(path):
        with RecordFile(path) as rec:
            return [parse(rec, i, **parse_kwargs) for i in range(len(rec))]

    paths = list(paths)
    if workers <= 1:
        for path in paths:
            yield from load(path)
        return
    with ThreadPoolExecutor(max_workers=workers) as pool:
        pending = []
        upcoming = iter(paths)
        for path in upcoming:
            pending.append(pool.submit(load, path))
            if len(pending) >= max(prefetch, 1):
                break
        while pending:
            examples = pending.pop(0).result()
            nxt = next(upcoming, None)
            if nxt is not None:
                pending.append(pool.submit(load, nxt))
            yield from examples


# ---- writer (convert_data.py:247-279) -----------------------------------------------------------------
def _varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _field(number, payload):
    return _varint((number << 3) | 2) + _varint(len(payload)) + payload


def _int64_feature(values):
    packed = b''.join(_varint(int(v)) for v in np.atleast_1d(values))
    return _field(3, _field(1, packed))                 # Feature.int64_list { value: packed }


def _bytes_feature(blob):
    return _field(1, _field(1, bytes(blob)))            # Feature.bytes_list { value }


_CRC_TABLE = None


def _masked_crc32c(data):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        table = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            table.append(c)
        _CRC_TABLE = table
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    c ^= 0xFFFFFFFF
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def encode_sequence_example(context, feature_lists):
    """Serialise one SequenceExample: ``context`` maps names to int64 scalars/lists, ``feature_lists`` maps names to
    lists of byte strings (one per step)."""
    ctx = b''.join(_field(1, _field(1, k.encode()) + _field(2, _int64_feature(v))) for k, v in context.items())
    fl = b''.join(_field(1, _field(1, k.encode()) + _field(2, b''.join(_field(1, _bytes_feature(s)) for s in steps)))
                  for k, steps in feature_lists.items())
    return _field(1, ctx) + _field(2, fl)


def write_sequence_examples(path, examples, compress=True):
    """Write serialised examples (bytes) as a TFRecord file, GZIP-compressed like the reference's
    (TFRecordCompressionType.GZIP, convert_data.py:247-248)."""
    opener = gzip.open if compress else open
    with opener(path, 'wb') as fh:
        for blob in examples:
            header = struct.pack('<Q', len(blob))
            fh.write(header + struct.pack('<I', _masked_crc32c(header)) + blob + struct.pack('<I', _masked_crc32c(blob)))
    return path
