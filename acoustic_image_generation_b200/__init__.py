"""B200-native acoustic-image front end and localisation scoring (one hot path of
IIT-PAVIS/Acoustic-Image-Generation): 512-bin pixel spectra -> 12-channel MFCC image ->
energy heat map -> IoU / consensus-IoU threshold sweep -> AUC, as hand-written sm_100a
kernels behind a C ABI (include/aig.h, csrc/libaig.so), with Python drop-ins that keep the
reference's function names and signatures (api.py).

Importing the package does not touch the GPU; the first compute call loads libaig.so and
fails loudly if the library or a B200 is missing (there is no CPU fallback).
"""
from . import synth, tables  # noqa: F401
from .api import (AcousticPath, AigError, REFERENCE_THRESHOLDS, _build_spectrograms_function,  # noqa: F401
                  _map_func_acoustic_images, _map_func_audio_samples_build_spectrogram, _map_func_mfcc, _normalize_acoustic_images_rescaled, _normalize_mfcc,
                  auc, butter_lowpass_filter, createfilters, default_path, find_logen,
                  get_feats, rates_as_written, success_rates)
from .metrics_io import read_accuracy_file, write_accuracy_file, write_area_file  # noqa: F401

__all__ = ['AcousticPath', 'AigError', 'REFERENCE_THRESHOLDS', 'auc', 'createfilters', 'default_path',
           'find_logen', 'get_feats', 'rates_as_written', 'success_rates', 'synth', 'tables',
           'read_accuracy_file', 'write_accuracy_file', 'write_area_file']
