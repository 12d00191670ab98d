"""End-to-end example without TensorFlow: write a small ACIVW-shaped data set as GZIP TFRecords (the reference's on-disk
format), read it back with libaig's reader, and run the scoring path of iouenergythreshold.py on the GPU for all 11
thresholds at once, producing the same intersection_{tau}_accuracy.txt / area.txt files.

    python examples/evaluate_tfrecords.py [out_dir]

The "reconstructed" images stand in for the UNet's output (out of scope here): the real image blended with noise.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import acoustic_image_generation_b200 as aig  # noqa: E402
from acoustic_image_generation_b200 import evaluate, synth, tfrecord  # noqa: E402


def write_dataset(data_dir, n_records=6, frames_per_record=12):
    """One record per second of video (12 acoustic frames), like convert_data.py:231-279."""
    paths = []
    for r in range(n_records):
        images = synth.smooth_images(frames_per_record, 100 + r)
        audio = synth.audio_rows(frames_per_record, 200 + r, np.int32)
        blob = tfrecord.encode_sequence_example(
            {'classes': r % 3, 'location': 1, 'audio_image/height': 36, 'audio_image/width': 48, 'audio_image/depth': 12,
             'audio_data/mics': 1, 'audio_data/samples': 1024},
            {'audio/image': [f.tobytes() for f in images], 'audio/data': [a.tobytes() for a in audio]})
        paths.append(tfrecord.write_sequence_examples(os.path.join(data_dir, 'Data_%03d.tfrecord' % (r + 1)), [blob]))
    return paths


def main(out_dir=None):
    out_dir = out_dir or tempfile.mkdtemp(prefix='aig_example_')
    os.makedirs(out_dir, exist_ok=True)
    paths = write_dataset(out_dir)
    path = aig.AcousticPath(0)
    ev = evaluate.AcivwEvaluation(path)
    rng = np.random.default_rng(0)
    for ex in tfrecord.iterate_examples(paths, workers=4):                   # parallel read + parse, flips as in _parse_sequence
        data = path.normalize_images(ex['audio_images'])                     # _map_func_acoustic_images
        mfcc = path.normalize_mfcc(path.build_spectrograms(ex['audio_samples']))   # _build_spectrograms_function + _map_func_mfcc
        assert mfcc.shape == (12, 12)
        recon = np.clip(data * np.float32(0.8) + np.float32(0.2) * rng.random(data.shape, dtype=np.float32), 0, 1)
        ev.add_batch(data, recon)
    res = ev.finish(out_dir)
    print('frames scored: %d   success rates: %s   AUC: %.4f' % (res['num'], np.round(res['rates'], 3).tolist(), res['auc']))
    print('metric files in', out_dir, ':', sorted(f for f in os.listdir(out_dir) if f.endswith('.txt')))
    return res, out_dir


if __name__ == '__main__':
    main(*sys.argv[1:2])
