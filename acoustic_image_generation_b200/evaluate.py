"""Batch drivers that play the role of the reference's evaluation scripts for the scoring path.

``AcivwEvaluation`` follows the post-``session.run`` loop of iouenergythreshold.py:213-236 (real vs
reconstructed acoustic image -> find_logen -> mean masks -> IoU -> success count) and
``FlickrEvaluation`` the loop of showimages_bb.py:287-328 (reconstructed image + annotator boxes ->
consensus IoU), but for all thresholds in one pass, batched on the GPU, sharded over ranks, and with
the counters kept on the device until ``finish()``.  ``finish`` all-reduces the count vector, and on
rank 0 writes the same ``intersection_{tau}_accuracy.txt`` / ``area.txt`` files the reference's
areaundercurve.py and meanstd.py consume.
"""
from __future__ import annotations

import numpy as np

from . import metrics_io
from .api import REFERENCE_THRESHOLDS, AcousticPath, auc, rates_as_written, success_rates


class _Evaluation:
    def __init__(self, path=None, thresholds=REFERENCE_THRESHOLDS, device=0):
        import torch
        self.path = path if path is not None else AcousticPath(device)
        self.thresholds = tuple(float(t) for t in thresholds)
        dev = torch.device('cuda', self.path.device)
        self._thr = torch.tensor(self.thresholds, dtype=torch.float64, device=dev)
        self.counts = torch.zeros(len(self.thresholds) + 1, dtype=torch.int64, device=dev)   # pos..., num

    def finish(self, data_dir=None, group=None):
        """Sum the counters over the ranks; returns {'pos', 'num', 'rates', 'auc', 'auc_exact'} and (rank 0) writes the
        metric files.

        The reduction runs on the handle's own NCCL communicator (aig_allreduce_counts, enqueued on the handle's stream
        behind the sweeps that filled the vector); a torch.distributed process whose handle has not joined one yet joins
        it here (the 128-byte id travels through the torch group, which is used for nothing else).
        'auc' follows the reference's pipeline to the letter: areaundercurve.py integrates the rates it parses back from
        the ``'iou {:6f}'`` files, i.e. rounded to six decimals; 'auc_exact' integrates the unrounded rates."""
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if distributed and not self.path.has_comm:
            self.path.init_comm(group=group)
        self.path.allreduce_counts(self.counts)
        host = self.counts.cpu().numpy()
        pos, num = host[:-1], int(host[-1])
        rates = success_rates(pos, num) if num else np.full(len(pos), np.nan)
        curve = num and len(pos) > 1
        area_exact = auc(self.thresholds, rates) if curve else float('nan')
        area = auc(self.thresholds, rates_as_written(rates)) if curve else float('nan')
        rank0 = not distributed or dist.get_rank(group) == 0
        if data_dir is not None and rank0 and num:
            for t, p in zip(self.thresholds, pos):
                metrics_io.write_accuracy_file(data_dir, t, int(p), num)
            metrics_io.write_area_file(data_dir, area)
        return {'pos': pos, 'num': num, 'rates': rates, 'auc': area, 'auc_exact': area_exact}


class AcivwEvaluation(_Evaluation):
    """ACIVW / AVIA: IoU between the energy masks of the real and the generated acoustic image."""

    def add_batch(self, data, reconstructed, normalize_first=False):
        """data, reconstructed: [B, 36, 48, 12] float32 (NumPy or CUDA tensors); returns per-frame (I, U).
        One kernel launch (aig_acivw_batch): energies and masks never leave the SM."""
        inter, union, _, _ = self.path.acivw_batch(data, reconstructed, self._thr, pos=self.counts[:-1],
                                                   num=self.counts[-1:], normalize_first=normalize_first)
        return inter, union


class FlickrEvaluation(_Evaluation):
    """FlickrSoundNet: consensus IoU between the up-sampled energy mask and the annotator boxes."""

    def __init__(self, path=None, thresholds=REFERENCE_THRESHOLDS, device=0, out_hw=(224, 298)):
        super().__init__(path, thresholds, device)
        self.out_hw = out_hw

    def add_batch(self, reconstructed, xmin, xmax, ymin, ymax, normalize_first=False):
        _, mask = self.path.energy(reconstructed, normalize_first=normalize_first)
        i2, u2, _, _ = self.path.ciou_sweep(mask, xmin, xmax, ymin, ymax, self._thr, out_hw=self.out_hw,
                                            pos=self.counts[:-1], num=self.counts[-1:])
        return i2, u2


def render_heatmaps(path, reconstructed, frames_bgr=None, out_hw=(224, 298), alpha=0.7, normalize_first=False):
    """The per-frame body of showvideo.py:217-233 / showimages.py:144-150 for a batch, on the GPU: find_logen ->
    bilinear up-sampling -> min/max normalisation -> jet colour map blended over the gray video frame.
    Returns RGB uint8 [N, H, W, 3] (PNG encoding and ffmpeg muxing stay with the caller)."""
    _, _, heat = path.energy_heatmap(reconstructed, normalize_first, *out_hw, want_energy=False, want_mask=False)
    return path.overlay(heat, frames_bgr, alpha)
