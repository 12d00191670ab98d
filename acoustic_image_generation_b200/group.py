"""One process, several GPUs: the shape of the reference's own callers.

The reference's evaluation scripts are single processes that feed batches to one GPU (iouenergythreshold.py:140-236,
showimages_bb.py:141-328; one whole run per threshold, scripts/iou.bash:47-53).  ``AcousticPathGroup`` keeps that
shape - one Python process, plain NumPy arrays in and out - and spreads every call over all visible B200s: one libaig
handle and one host thread per device (the C calls release the GIL, and a handle's pinned staging ring and copy streams
are its own), frames sharded contiguously (``sharding.shard_range``), results written straight into slices of the
caller's arrays, the per-threshold success counts kept on each device and summed once - through a communicator made
with ncclCommInitAll (``aig_comm_init_all`` / ``aig_group_allreduce_counts``, one grouped NCCL call from this thread),
or on the host when NCCL is absent.  Frames are independent, so every result equals the one-GPU result bit for bit.

``bench.py --single-process`` measures this path; the one-process-per-GPU form (torchrun) stays the headline.
"""
from __future__ import annotations

import ctypes
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib, metrics_io
from .api import (FRAME_H, FRAME_PIXELS, FRAME_W, FFT_LEN, HEAT_H, HEAT_W, MFCC_NUM, REFERENCE_THRESHOLDS, AcousticPath,
                  AigError, _torch, auc, rates_as_written, success_rates)
from .sharding import shard_range


class AcousticPathGroup:
    """Per-device handles behind one set of batch calls.  ``devices``: CUDA ordinals (default: all visible)."""

    def __init__(self, devices=None, thresholds=REFERENCE_THRESHOLDS, use_nccl=True):
        torch = _torch()
        if devices is None:
            devices = list(range(torch.cuda.device_count()))
        devices = [int(d) for d in devices]
        if not devices or len(set(devices)) != len(devices):
            raise ValueError('devices must be a non-empty list of distinct CUDA ordinals, got %r' % (devices,))
        self.devices = devices
        self.paths = [AcousticPath(d) for d in devices]
        self.thresholds = tuple(float(t) for t in thresholds)
        self._pool = ThreadPoolExecutor(max_workers=len(devices), thread_name_prefix='aig-dev')
        self._thr = [torch.tensor(self.thresholds, dtype=torch.float64, device=torch.device('cuda', d)) for d in devices]
        self.counts = [torch.zeros(len(self.thresholds) + 1, dtype=torch.int64, device=torch.device('cuda', d)) for d in devices]
        self.nccl = False
        if use_nccl and len(devices) > 1:
            lib = _lib.load()
            handles = (ctypes.c_void_p * len(devices))(*[p._h for p in self.paths])
            code = lib.aig_comm_init_all(handles, len(devices))
            if code == 0:
                self.nccl = True
                for p in self.paths:
                    p._has_comm = True
            elif code != -2:                      # AIG_ERR_NO_DEVICE = no NCCL library: fall back to the host sum
                raise AigError(code, (lib.aig_last_error(self.paths[0]._h) or b'').decode())

    # -- plumbing -----------------------------------------------------------------------------
    def close(self):
        self._pool.shutdown(wait=True)
        for p in self.paths:
            p.close()
        self.paths = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __len__(self):
        return len(self.paths)

    def shards(self, n):
        """[(device index, first frame, stop frame)] of the non-empty shards of n frames."""
        out = []
        for i in range(len(self.paths)):
            lo, hi = shard_range(n, i, len(self.paths))
            if hi > lo:
                out.append((i, lo, hi))
        return out

    def _run(self, n, fn, chunk=None):
        """fn(device index, path, lo, hi) over [0, n), one host thread per device; re-raises the first failure.

        ``chunk=None``: one contiguous shard per device (``shards``).  ``chunk=c``: the frames are cut into pieces of c
        and every device thread takes the next piece when it has finished its last one.  On boxes whose GPUs do not see
        the same host bandwidth (profiles/r02_h2d_scale.txt: 23 GB/s for four of the eight links, 36 GB/s for the other
        four when all run at once) equal shards finish with the slowest link while dynamic pieces keep every link busy
        to the end; results do not depend on which device computed which piece."""
        if chunk is None or len(self.paths) == 1 or n <= chunk:
            futures = [self._pool.submit(fn, i, self.paths[i], lo, hi) for i, lo, hi in self.shards(n)]
            return [f.result() for f in futures]
        pieces = [(lo, min(lo + chunk, n)) for lo in range(0, n, chunk)]
        lock, state = threading.Lock(), {'next': 0}

        def worker(i):
            out = []
            while True:
                with lock:
                    k = state['next']
                    state['next'] = k + 1
                if k >= len(pieces):
                    return out
                out.append(fn(i, self.paths[i], *pieces[k]))
        futures = [self._pool.submit(worker, i) for i in range(len(self.paths))]
        return [r for f in futures for r in f.result()]

    def synchronize(self):
        for p in self.paths:
            p.synchronize()

    @property
    def launch_count(self):
        return sum(p.launch_count for p in self.paths)

    @staticmethod
    def _host(x, dtype, shape_tail):
        arr = np.ascontiguousarray(x, dtype=dtype)
        per = int(np.prod(shape_tail))
        if arr.size % per:
            raise ValueError('array of %d values is not a whole number of %s frames' % (arr.size, shape_tail))
        return arr.reshape((arr.size // per,) + tuple(shape_tail))

    # -- stages 1 + 2 -------------------------------------------------------------------------
    def mfcc_energy(self, power, flip=False, normalize_first=True, out=None, chunk_frames=64):
        """[N, 36, 48, 512] float32 spectra (NumPy) -> (mfcc f32 [N,36,48,12], energy f64 [N,36,48], mask u8 [N,36,48]),
        frames spread over the devices in pieces of ``chunk_frames`` taken dynamically (None: one equal shard per
        device); ``out`` supplies the result arrays (e.g. pinned)."""
        power = self._host(power, np.float32, (FRAME_H, FRAME_W, FFT_LEN))
        n = len(power)
        if out is None:
            out = (np.empty((n, FRAME_H, FRAME_W, MFCC_NUM), np.float32), np.empty((n, FRAME_H, FRAME_W), np.float64),
                   np.empty((n, FRAME_H, FRAME_W), np.uint8))
        mfcc, energy, mask = out
        self._run(n, lambda i, p, lo, hi: p.mfcc_energy(power[lo:hi], flip=flip, normalize_first=normalize_first,
                                                       out=(mfcc[lo:hi], energy[lo:hi], mask[lo:hi])), chunk=chunk_frames)
        return mfcc, energy, mask

    def energy_heatmap(self, images, normalize_first=False, out_h=HEAT_H, out_w=HEAT_W):
        """find_logen -> up-sampling -> normalisation for a batch (showvideo.py:226-228): float32 [N, out_h, out_w]."""
        images = self._host(images, np.float32, (FRAME_H, FRAME_W, MFCC_NUM))
        heat = np.empty((len(images), out_h, out_w), np.float32)
        self._run(len(images), lambda i, p, lo, hi: p.energy_heatmap(images[lo:hi], normalize_first, out_h, out_w,
                                                                     want_energy=False, want_mask=False, out=heat[lo:hi]))
        return heat

    # -- stage 3: evaluations with device-resident counters -------------------------------------
    def reset_counts(self):
        for c in self.counts:
            c.zero_()

    def add_acivw_batch(self, real, reconstructed, normalize_first=False):
        """The ACIVW evaluation step (iouenergythreshold.py:213-229) on a batch sharded over the devices; the success
        counts accumulate on each device.  Returns per-frame (I, U) as NumPy int64."""
        real = self._host(real, np.float32, (FRAME_H, FRAME_W, MFCC_NUM))
        recon = self._host(reconstructed, np.float32, (FRAME_H, FRAME_W, MFCC_NUM))
        if len(real) != len(recon):
            raise ValueError('image batches differ: %d vs %d frames' % (len(real), len(recon)))
        inter, union = np.empty(len(real), np.int64), np.empty(len(real), np.int64)

        def work(i, p, lo, hi):
            a, b, _, _ = p.acivw_batch(real[lo:hi], recon[lo:hi], self._thr[i], pos=self.counts[i][:-1],
                                       num=self.counts[i][-1:], normalize_first=normalize_first)
            inter[lo:hi], union[lo:hi] = a, b
        self._run(len(real), work)
        return inter, union

    def add_flickr_batch(self, reconstructed, xmin, xmax, ymin, ymax, normalize_first=False, out_hw=(HEAT_H, HEAT_W)):
        """The FlickrSoundNet consensus-IoU step (showimages_bb.py:288-321) on a batch sharded over the devices."""
        recon = self._host(reconstructed, np.float32, (FRAME_H, FRAME_W, MFCC_NUM))
        boxes = [self._host(v, np.int32, (3,)) for v in (xmin, xmax, ymin, ymax)]
        i2, u2 = np.empty(len(recon), np.int64), np.empty(len(recon), np.int64)

        def work(i, p, lo, hi):
            _, mask = p.energy(recon[lo:hi], normalize_first=normalize_first)
            a, b, _, _ = p.ciou_sweep(mask, *[bx[lo:hi] for bx in boxes], self._thr[i], out_hw=out_hw,
                                      pos=self.counts[i][:-1], num=self.counts[i][-1:])
            i2[lo:hi], u2[lo:hi] = a, b
        self._run(len(recon), work)
        return i2, u2

    def reduce_counts(self):
        """Sum of the per-device count vectors as NumPy int64 [K + 1] (pos..., num).  With the NCCL communicator every
        device ends up holding the sum (one grouped all-reduce); otherwise the vectors are added on the host."""
        if self.nccl:
            lib = _lib.load()
            handles = (ctypes.c_void_p * len(self.paths))(*[p._h for p in self.paths])
            bufs = (ctypes.c_void_p * len(self.paths))(*[c.data_ptr() for c in self.counts])
            code = lib.aig_group_allreduce_counts(handles, bufs, len(self.paths), int(self.counts[0].numel()))
            if code != 0:
                raise AigError(code, (lib.aig_last_error(self.paths[0]._h) or b'').decode())
            self.paths[0].synchronize()
            total = self.counts[0].cpu().numpy().copy()
            # the devices now all hold the total: keep it on device 0 only, so that further batches accumulate once
            for c in self.counts[1:]:
                c.zero_()
            return total
        self.synchronize()
        return np.sum(np.stack([c.cpu().numpy() for c in self.counts], 0), axis=0)

    def finish(self, data_dir=None):
        """{'pos', 'num', 'rates', 'auc', 'auc_exact'} of everything added so far; writes the reference's metric files."""
        host = self.reduce_counts()
        pos, num = host[:-1], int(host[-1])
        rates = success_rates(pos, num) if num else np.full(len(pos), np.nan)
        curve = num and len(pos) > 1
        area_exact = auc(self.thresholds, rates) if curve else float('nan')
        area = auc(self.thresholds, rates_as_written(rates)) if curve else float('nan')
        if data_dir is not None and num:
            for t, p in zip(self.thresholds, pos):
                metrics_io.write_accuracy_file(data_dir, t, int(p), num)
            metrics_io.write_area_file(data_dir, area)
        return {'pos': pos, 'num': num, 'rates': rates, 'auc': area, 'auc_exact': area_exact}
