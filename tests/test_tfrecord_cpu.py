"""N3: the GZIP TFRecord / SequenceExample reader of libaig against records serialised by the official protobuf
runtime (message types built from descriptors that restate tensorflow/core/example/{example,feature}.proto), plus the
package's own writer.  CPU only - the reader is host code."""
import gzip
import struct

import numpy as np
import pytest

from acoustic_image_generation_b200 import AigError, synth, tfrecord


def _tf_example_messages():
    """tf.train.SequenceExample and friends, declared at run time (no TensorFlow needed)."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name='aig_test_example.proto', package='aigtest', syntax='proto3')
    L = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields, nested=()):
        m = fd.message_type.add(name=name)
        for fname, number, ftype, label, type_name in fields:
            f = m.field.add(name=fname, number=number, type=ftype, label=label)
            if type_name:
                f.type_name = type_name
        return m

    msg('BytesList', [('value', 1, L.TYPE_BYTES, L.LABEL_REPEATED, None)])
    msg('FloatList', [('value', 1, L.TYPE_FLOAT, L.LABEL_REPEATED, None)])
    msg('Int64List', [('value', 1, L.TYPE_INT64, L.LABEL_REPEATED, None)])
    msg('Feature', [('bytes_list', 1, L.TYPE_MESSAGE, L.LABEL_OPTIONAL, '.aigtest.BytesList'),
                    ('float_list', 2, L.TYPE_MESSAGE, L.LABEL_OPTIONAL, '.aigtest.FloatList'),
                    ('int64_list', 3, L.TYPE_MESSAGE, L.LABEL_OPTIONAL, '.aigtest.Int64List')])
    msg('FeatureList', [('feature', 1, L.TYPE_MESSAGE, L.LABEL_REPEATED, '.aigtest.Feature')])
    for holder, value in (('Features', 'Feature'), ('FeatureLists', 'FeatureList')):
        m = msg(holder, [('entry', 1, L.TYPE_MESSAGE, L.LABEL_REPEATED, '.aigtest.%s.Entry' % holder)])
        e = m.nested_type.add(name='Entry')              # map<string, V> is wire-identical to repeated {key=1, value=2}
        e.field.add(name='key', number=1, type=L.TYPE_STRING, label=L.LABEL_OPTIONAL)
        e.field.add(name='value', number=2, type=L.TYPE_MESSAGE, label=L.LABEL_OPTIONAL, type_name='.aigtest.' + value)
    msg('SequenceExample', [('context', 1, L.TYPE_MESSAGE, L.LABEL_OPTIONAL, '.aigtest.Features'),
                            ('feature_lists', 2, L.TYPE_MESSAGE, L.LABEL_OPTIONAL, '.aigtest.FeatureLists')])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, 'GetMessageClass', None)
    return {n: get(pool.FindMessageTypeByName('aigtest.' + n)) for n in ('SequenceExample', 'Feature')}


def _official_example(context, feature_lists):
    cls = _tf_example_messages()['SequenceExample']
    ex = cls()
    for k, v in context.items():
        e = ex.context.entry.add(key=k)
        e.value.int64_list.value.extend(int(x) for x in np.atleast_1d(v))
    for k, steps in feature_lists.items():
        e = ex.feature_lists.entry.add(key=k)
        for s in steps:
            e.value.feature.add().bytes_list.value.append(bytes(s))
    return ex.SerializeToString()


def _frame(blob):
    header = struct.pack('<Q', len(blob))
    return header + struct.pack('<I', tfrecord._masked_crc32c(header)) + blob + struct.pack('<I', tfrecord._masked_crc32c(blob))


def test_native_crc32c_both_paths():
    """libaig's CRC-32C: the SSE4.2 path (when the CPU has it) and the table loop against the known answer, the
    Python writer's implementation, and each other on unaligned ragged buffers."""
    from acoustic_image_generation_b200 import _lib
    lib = _lib.load()
    crc = lambda b, table: lib.aig_crc32c(b, len(b), table)
    assert crc(b'123456789', 0) == crc(b'123456789', 1) == 0xE3069283
    rng = np.random.default_rng(4)
    buf = rng.integers(0, 256, 100003, dtype=np.uint8).tobytes()
    for start, n in ((0, 100003), (1, 9), (3, 64), (5, 4097), (7, 1), (2, 65536)):
        piece = buf[start:start + n]
        mv = np.frombuffer(buf, np.uint8)[start:start + n]                 # genuinely unaligned start address
        fast = lib.aig_crc32c(mv.ctypes.data, n, 0)
        assert fast == lib.aig_crc32c(mv.ctypes.data, n, 1) == crc(piece, 1)
        c = fast
        assert tfrecord._masked_crc32c(piece) == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_crc32c_known_answer():
    # CRC-32C("123456789") = 0xE3069283; the TFRecord mask is rot-right 15 plus 0xA282EAD8
    c = 0xE3069283
    assert tfrecord._masked_crc32c(b'123456789') == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


@pytest.mark.parametrize('compress', [True, False])
def test_reader_on_officially_serialised_records(tmp_path, compress):
    images = synth.sigmoid_images(12, 5)                                  # one second of 12 fps acoustic images
    audio = synth.audio_rows(12, 6, np.int32)
    video = np.random.default_rng(0).integers(0, 255, (3, 8, 10, 3), dtype=np.uint8)
    ctx = {'classes': 7, 'location': 2, 'audio_image/height': 36, 'audio_image/width': 48, 'audio_image/depth': 12,
           'audio_data/mics': 1, 'audio_data/samples': 1024, 'video/height': 8, 'video/width': 10, 'video/depth': 3,
           'xmin': [10, 0, 30], 'xmax': [100, 0, 298], 'ymin': [5, 0, 0], 'ymax': [50, 0, 224], 'neg': -5}
    lists = {'audio/image': [f.tobytes() for f in images], 'audio/data': [a.tobytes() for a in audio],
             'video/image': [v.tobytes() for v in video]}
    blobs = [_official_example(ctx, lists), _official_example({'classes': 1, 'location': 0}, {})]
    path = str(tmp_path / ('data.tfrecord'))
    payload = b''.join(_frame(b) for b in blobs)
    with (gzip.open(path, 'wb') if compress else open(path, 'wb')) as fh:
        fh.write(payload)
    with tfrecord.RecordFile(path) as rec:
        assert len(rec) == 2
        ex = tfrecord.parse_flickr_example(rec, 0)
        assert ex['classes'] == 7 and ex['location'] == 2
        assert np.array_equal(ex['audio_images'], images)
        assert np.array_equal(ex['audio_samples'], audio)
        assert np.array_equal(ex['video_images'], video)
        assert ex['xmax'].tolist() == [100, 0, 298] and ex['ymin'].dtype == np.int32
        assert rec.context(0, 'neg').tolist() == [-5]
        flipped = tfrecord.parse_acoustic_example(rec, 0)['audio_images']      # outdoor_data_mfcc.py:314-315
        assert np.array_equal(flipped, images[:, ::-1, ::-1, :])
        assert tfrecord.parse_acoustic_example(rec, 1) == {'classes': 1, 'location': 0}
        with pytest.raises(AigError):
            rec.context(0, 'missing')
        with pytest.raises(AigError):
            rec.sequence(1, 'audio/image', np.float32)


def test_own_writer_is_byte_identical_to_protobuf_and_round_trips(tmp_path):
    images = synth.sigmoid_images(3, 9)
    ctx = {'classes': 3, 'location': 1, 'audio_image/height': 36, 'audio_image/width': 48, 'audio_image/depth': 12}
    lists = {'audio/image': [f.tobytes() for f in images]}
    mine = tfrecord.encode_sequence_example(ctx, lists)
    assert mine == _official_example(ctx, lists)
    path = tfrecord.write_sequence_examples(str(tmp_path / 'Data_001.tfrecord'), [mine] * 5)
    with tfrecord.RecordFile(path) as rec:
        assert len(rec) == 5
        assert np.array_equal(tfrecord.parse_acoustic_example(rec, 4, flip=False)['audio_images'], images)


def test_corruption_is_detected(tmp_path):
    blob = tfrecord.encode_sequence_example({'classes': 1, 'location': 1}, {})
    good = _frame(blob)
    bad = bytearray(good)
    bad[14] ^= 0xFF
    for name, payload in (('flip.tfrecord', bytes(bad)), ('short.tfrecord', good[:-3])):
        p = str(tmp_path / name)
        open(p, 'wb').write(payload)
        with pytest.raises(AigError):
            tfrecord.RecordFile(p)
    with pytest.raises(AigError):
        tfrecord.RecordFile(str(tmp_path / 'does_not_exist.tfrecord'))
    empty = str(tmp_path / 'empty.tfrecord')
    open(empty, 'wb').close()
    with tfrecord.RecordFile(empty) as rec:
        assert len(rec) == 0


def test_malformed_messages_never_crash_the_reader(tmp_path):
    """Valid framing (checksums recomputed) around damaged SequenceExample bytes: every accessor either answers or
    raises AigError - truncated varints, over-long length prefixes and wrong wire types must not read out of bounds."""
    images = synth.sigmoid_images(2, 4)
    good = tfrecord.encode_sequence_example(
        {'classes': 3, 'location': 1, 'audio_image/height': 36, 'audio_image/width': 48, 'audio_image/depth': 12,
         'xmin': [1, 2, 3]}, {'audio/image': [f.tobytes()[:64] for f in images], 'audio/data': [b'\x01\x02\x03\x04'] * 3})
    rng = np.random.default_rng(11)
    blobs = [good[:k] for k in range(0, len(good), 7)]                      # every kind of truncation
    for _ in range(300):                                                     # byte flips, biased towards the tags
        b = bytearray(good)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        blobs.append(bytes(b))
    blobs.append(b'\xff' * 40)                                               # endless varint
    blobs.append(b'\x0a\xff\xff\xff\xff\x0f')                                # length prefix far past the end
    path = str(tmp_path / 'fuzz.tfrecord')
    with open(path, 'wb') as fh:
        fh.write(b''.join(_frame(b) for b in blobs))
    answered = raised = 0
    with tfrecord.RecordFile(path) as rec:
        assert len(rec) == len(blobs)
        for i in range(len(rec)):
            for call in (lambda: rec.context(i, 'classes'), lambda: rec.context(i, 'xmin'),
                         lambda: rec.sequence(i, 'audio/image', np.uint8), lambda: rec.sequence(i, 'audio/data', np.uint8),
                         lambda: tfrecord.parse_acoustic_example(rec, i)):
                try:
                    call()
                    answered += 1
                except (AigError, ValueError, KeyError):
                    raised += 1
    assert answered > 0 and raised > 0


def test_parallel_iteration_keeps_file_and_record_order(tmp_path):
    paths = []
    for r in range(9):
        blobs = [tfrecord.encode_sequence_example(
            {'classes': 10 * r + k, 'location': r, 'audio_image/height': 36, 'audio_image/width': 48, 'audio_image/depth': 12},
            {'audio/image': [f.tobytes() for f in synth.sigmoid_images(2, 50 + 3 * r + k)]}) for k in range(1 + r % 3)]
        paths.append(tfrecord.write_sequence_examples(str(tmp_path / ('Data_%03d.tfrecord' % r)), blobs))
    serial = list(tfrecord.iterate_examples(paths, workers=1))
    for workers, prefetch in ((4, 8), (2, 1), (16, 3)):
        got = list(tfrecord.iterate_examples(paths, workers=workers, prefetch=prefetch))
        assert [e['classes'] for e in got] == [e['classes'] for e in serial] == [10 * r + k for r in range(9) for k in range(1 + r % 3)]
        assert all(np.array_equal(a['audio_images'], b['audio_images']) for a, b in zip(got, serial))
    unflipped = list(tfrecord.iterate_examples(paths[:1], flip=False))
    assert np.array_equal(unflipped[0]['audio_images'][:, ::-1, ::-1, :], serial[0]['audio_images'])
    assert list(tfrecord.iterate_examples([])) == []
    with pytest.raises(AigError):
        list(tfrecord.iterate_examples(paths[:2] + [str(tmp_path / 'missing.tfrecord')]))
