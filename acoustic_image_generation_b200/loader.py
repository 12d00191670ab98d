"""The hot-path side of the reference's input pipeline as device tensors (SURVEY.md section 8(f), rows N1-N3 composed).

``SoundDataLoader`` / ``ActionsDataLoader`` (dataloader/outdoor_data_mfcc.py:62-101) read GZIP TFRecords, parse them four at
a time, build the audio MFCC with ``_build_spectrograms_function`` (:78-82, :796-824), low-pass the waveform (:558-575),
min-max normalise every acoustic image (:657-679) and every MFCC vector (:681-703) and batch the result; the trainers then
tile the MFCC vector into the ``mfccmap`` conditioning image (trainer/mfcctrainer.py:38-40).  ``AcousticBatches`` yields
exactly those tensors, already on the GPU: records are inflated and parsed by libaig's reader on a few threads, everything
after the upload is one of libaig's kernels.  Video frames, labels and shuffling policies stay with the caller (the
records' 'classes' / 'location' come along as NumPy arrays).
"""
from __future__ import annotations

import numpy as np

from . import tfrecord
from .api import AcousticPath, FRAME_H, FRAME_W, MFCC_NUM, _torch


class AcousticBatches:
    """Iterate batches of ``batch_frames`` acoustic frames over TFRecord files.

    Each batch is a dict of CUDA tensors
      'acoustic'  float32 [B, 36, 48, 12]  acoustic images, flipped as in _parse_sequence (:314-315) and min-max normalised
                                           per frame (_map_func_acoustic_images)
      'mfcc'      float32 [B, 12]          audio MFCC per frame's 1024 samples, min-max normalised per vector (_map_func_mfcc)
      'mfccmap'   float32 [B, 36, 48, 12]  the MFCC vector tiled over the image (only with ``tile=True``)
      'filtered'  float32 [B, 1024]        low-passed waveform (only with ``low_pass=True``; butter_lowpass_filter, :558-575)
      'filtered_mfcc' float32 [B, 12]      MFCC of the low-passed waveform, normalised per vector (with ``low_pass=True``):
                                           _map_func_audio_samples_build_spectrogram returns both MFCCs (:788-794) and
                                           _map_func_mfcc normalises both (:692-693)
    plus NumPy int64 arrays 'classes' and 'location' [B].  The last batch may be short unless ``drop_last``.
    Records without audio samples yield no 'mfcc' / 'mfccmap' / 'filtered'.
    """

    def __init__(self, paths, batch_frames=16, path=None, device=0, workers=4, prefetch=8, flip=True, tile=True,
                 low_pass=False, drop_last=False):
        if batch_frames < 1:
            raise ValueError('batch_frames must be positive')
        self.paths = list(paths)
        self.batch_frames = int(batch_frames)
        self.path = path if path is not None else AcousticPath(device)
        self.workers, self.prefetch = workers, prefetch
        self.flip, self.tile, self.low_pass, self.drop_last = flip, tile, low_pass, drop_last

    def __iter__(self):
        images, audio, classes, location = [], [], [], []
        pending, has_audio = 0, None
        for ex in tfrecord.iterate_examples(self.paths, workers=self.workers, prefetch=self.prefetch, flip=self.flip):
            if 'audio_images' not in ex:
                continue
            t = len(ex['audio_images'])
            images.append(ex['audio_images'])
            if has_audio is None:
                has_audio = 'audio_samples' in ex
            elif has_audio != ('audio_samples' in ex):
                raise ValueError('records with and without audio samples cannot share a batch stream')
            if 'audio_samples' in ex:
                rows = ex['audio_samples']
                if len(rows) != t:            # several microphones per frame: the reference keeps mics = 1 (convert_data.py)
                    raise ValueError('record holds %d audio rows for %d acoustic frames' % (len(rows), t))
                audio.append(rows)
            classes.append(np.full(t, ex['classes'], np.int64))
            location.append(np.full(t, ex['location'], np.int64))
            pending += t
            while pending >= self.batch_frames:
                yield self._emit(images, audio, classes, location, self.batch_frames)
                pending -= self.batch_frames
        if pending and not self.drop_last:
            yield self._emit(images, audio, classes, location, pending)

    @staticmethod
    def _take(chunks, n):
        """Remove and return the first n rows of a list of arrays."""
        out, need = [], n
        while need:
            head = chunks[0]
            if len(head) <= need:
                out.append(chunks.pop(0))
                need -= len(head)
            else:
                out.append(head[:need])
                chunks[0] = head[need:]
                need = 0
        return np.concatenate(out, 0) if len(out) > 1 else np.ascontiguousarray(out[0])

    def _emit(self, images, audio, classes, location, n):
        """Host arrays go to the device through libaig itself (``device_out=True``): ordinary (pageable) NumPy memory is
        staged through the handle's pinned ring by its copy threads (host_staging.h, 4-5x the driver's pageable path) and
        the kernel's result stays on the device - no torch ``.to(device)`` copy in between."""
        path = self.path
        batch = {'classes': self._take(classes, n), 'location': self._take(location, n)}
        batch['acoustic'] = path.normalize_images(self._take(images, n), device_out=True).reshape(n, FRAME_H, FRAME_W, MFCC_NUM)
        if audio:
            wav = self._take(audio, n)
            if self.low_pass:
                filtered = path.butter_lowpass_filter(wav, device_out=True)           # float32 [n, 1024] on the device
                batch['filtered'] = filtered
                batch['filtered_mfcc'] = path.normalize_mfcc(path.build_spectrograms(filtered))
            mfcc = path.normalize_mfcc(path.build_spectrograms(wav, device_out=True))
            batch['mfcc'] = mfcc
            if self.tile:
                batch['mfccmap'] = path.tile_mfcc(mfcc)
        return batch
