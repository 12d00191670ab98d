"""Host-side mirror of the reference's interface for the acoustic-image path.

The reference exposes this path as Python callables on NumPy arrays.  The functions
here keep those names, argument meanings and return conventions and run the
arithmetic on a B200 through libaig.so (include/aig.h):

    createfilters, get_feats                 dataloader/outdoor_data_mfcc.py:826-876
    _normalize_acoustic_images_rescaled,
    _map_func_acoustic_images                dataloader/outdoor_data_mfcc.py:657-679
    find_logen                               iouenergythreshold.py:294-323
    iou_sweep / ciou_sweep / auc             iouenergythreshold.py:213-236,
                                             showimages_bb.py:287-328, areaundercurve.py:26-40

Arrays may be NumPy arrays (host; results come back as NumPy) or CUDA tensors / anything
exposing ``__cuda_array_interface__`` or ``__dlpack__`` (device; results are torch CUDA tensors,
nothing is copied to the host).  There is no CPU fallback: without libaig.so and a B200 every compute
call raises.
"""
from __future__ import annotations

import ctypes
import threading

import numpy as np

from . import _lib, tables
from ._lib import AigError

FRAME_H, FRAME_W, FRAME_PIXELS, MFCC_NUM, FFT_LEN = 36, 48, 1728, 12, 512
HEAT_H, HEAT_W = 224, 298                      # cv2.resize(map, (298, 224)), showimages.py:147
REFERENCE_THRESHOLDS = (0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)   # areaundercurve.py:27

_TORCH_DTYPES = {}


def _torch():
    import torch
    if not _TORCH_DTYPES:
        _TORCH_DTYPES.update({np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
                              np.dtype(np.uint8): torch.uint8, np.dtype(np.int32): torch.int32,
                              np.dtype(np.int64): torch.int64})
    return torch


def _is_torch(x):
    return type(x).__module__.startswith('torch') and hasattr(x, 'data_ptr')


def _is_cuda_tensor(x):
    return _is_torch(x) and x.is_cuda


class _Arg:
    """A caller buffer resolved to (pointer, keep-alive object)."""

    __slots__ = ('ptr', 'keep', 'shape', 'on_device', 'torch_device')

    def __init__(self, x, dtype, writable=False):
        dtype = np.dtype(dtype)
        self.torch_device = None
        if not _is_torch(x) and not hasattr(x, '__cuda_array_interface__') and not isinstance(x, np.ndarray) \
                and hasattr(x, '__dlpack__'):
            x = _torch().from_dlpack(x)          # DLPack producers (CuPy, JAX, TF ...): zero-copy view as a torch tensor
        if _is_torch(x):
            torch = _torch()
            want = _TORCH_DTYPES[dtype]
            if x.dtype != want or not x.is_contiguous():
                if writable:
                    raise ValueError('output tensor must be contiguous %s' % want)
                x = x.to(want).contiguous()
            self.ptr, self.keep, self.shape = x.data_ptr(), x, tuple(x.shape)
            self.on_device = x.is_cuda
            if x.is_cuda:
                self.torch_device = x.device
        elif hasattr(x, '__cuda_array_interface__'):
            cai = x.__cuda_array_interface__
            if np.dtype(cai['typestr']) != dtype or cai.get('strides'):
                raise ValueError('device array must be contiguous %s' % dtype)
            self.ptr, self.keep, self.shape, self.on_device = cai['data'][0], x, tuple(cai['shape']), True
        else:
            arr = np.asarray(x)
            if writable:
                if arr.dtype != dtype or not arr.flags.c_contiguous or not arr.flags.writeable:
                    raise ValueError('output array must be a writable C-contiguous %s array' % dtype)
            else:
                arr = np.ascontiguousarray(arr, dtype=dtype)
            self.ptr, self.keep, self.shape, self.on_device = arr.ctypes.data, arr, arr.shape, False


class AcousticPath:
    """One libaig handle: a CUDA device plus a stream.

    ``stream=None`` uses the legacy default stream (ordered with torch's default stream);
    pass ``torch.cuda.current_stream().cuda_stream`` to enqueue on a torch stream.  A handle is
    not thread-safe; the module-level drop-ins below keep one per thread, like giving each of
    the reference's four tf.data workers (outdoor_data_mfcc.py:82) its own.
    """

    def __init__(self, device=0, stream=None, tables_=None):
        self._lib = _lib.load()
        self._h = ctypes.c_void_p()
        self.device = int(device)
        code = self._lib.aig_create(self.device, int(stream or 0), ctypes.byref(self._h))
        if code != 0:
            raise AigError(code, (self._lib.aig_last_error(None) or b'').decode())
        self._tables_key = None
        self.set_tables(*(tables_ if tables_ is not None else tables.reference_tables()))

    # -- plumbing -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None) is not None and self._h.value:
            self._lib.aig_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, code):
        if code != 0:
            raise AigError(code, (self._lib.aig_last_error(self._h) or b'').decode())

    def synchronize(self):
        self._check(self._lib.aig_synchronize(self._h))

    @property
    def launch_count(self):
        return int(self._lib.aig_launch_count(self._h))

    def set_option(self, name, value):
        """Tuning / instrumentation knobs of include/aig.h: mfcc_variant, chain_chunk_frames, chain_overlap, profile."""
        self._check(self._lib.aig_set_option(self._h, name.encode(), int(value)))

    def set_mfcc_variant(self, variant):
        self.set_option('mfcc_variant', variant)

    @property
    def has_comm(self):
        """True once init_comm / AcousticPathGroup has joined this handle to a native (NCCL) communicator."""
        return getattr(self, '_has_comm', False)

    def selftest(self, which):
        """Device self-tests of the energy stage's arithmetic shortcuts: 0 lifter division, 1 table exp, 2 hoisted-
        reciprocal min-max division (see aig_selftest)."""
        out = (ctypes.c_uint64 * 4)()
        self._check(self._lib.aig_selftest(self._h, int(which), out))
        return [int(v) for v in out]

    def profile_read(self):
        """{'mfcc': (ms, launches), 'energy': (...), 'other': (...)} gathered since the last read (needs profile=1)."""
        ms = (ctypes.c_double * 3)()
        cnt = (ctypes.c_int64 * 3)()
        self._check(self._lib.aig_profile_read(self._h, ms, cnt))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(('mfcc', 'energy', 'other'))}

    def _a(self, x, dtype, writable=False):
        """Resolve a caller buffer and refuse CUDA tensors that live on another device than this handle's."""
        arg = _Arg(x, dtype, writable)
        if arg.torch_device is not None and arg.torch_device.index not in (None, self.device):
            raise ValueError('tensor is on %s but this AcousticPath runs on cuda:%d' % (arg.torch_device, self.device))
        return arg

    def _empty(self, shape, dtype, like, device_out=False):
        """Output buffer on the side the input lives on; ``device_out`` forces a CUDA tensor for a host input (the library
        then uploads the input itself - through its pinned staging ring when the array is pageable - and the result
        never returns to the host)."""
        if device_out or like.torch_device is not None or like.on_device:
            torch = _torch()
            dev = like.torch_device if like.torch_device is not None else torch.device('cuda', self.device)
            return torch.empty(shape, dtype=_TORCH_DTYPES[np.dtype(dtype)], device=dev)
        return np.empty(shape, dtype=dtype)

    # -- multi-GPU ----------------------------------------------------------------------------
    def init_comm(self, rank=None, world=None, group=None):
        """Join the NCCL communicator of libaig (aig_comm_init).  The 128-byte unique id is created on rank 0
        and shipped through the already-initialised torch.distributed group, which is used only for that."""
        import torch.distributed as dist
        rank = dist.get_rank(group) if rank is None else rank
        world = dist.get_world_size(group) if world is None else world
        ident = (ctypes.c_uint8 * 128)()
        if rank == 0:
            code = self._lib.aig_comm_unique_id(ident)
            if code != 0:
                raise AigError(code, (self._lib.aig_last_error(None) or b'').decode())
        box = [bytes(ident)]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = (ctypes.c_uint8 * 128).from_buffer_copy(box[0])
        self._check(self._lib.aig_comm_init(self._h, ident, int(rank), int(world)))
        self._has_comm = True

    def init_comm_from_id(self, ident, rank, world):
        """Join a communicator whose 128-byte id (aig_comm_unique_id) was shipped by other means (a file, MPI ...)."""
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(bytes(ident))
        self._check(self._lib.aig_comm_init(self._h, buf, int(rank), int(world)))
        self._has_comm = True

    @staticmethod
    def comm_unique_id():
        ident = (ctypes.c_uint8 * 128)()
        lib = _lib.load()
        code = lib.aig_comm_unique_id(ident)
        if code != 0:
            raise AigError(code, (lib.aig_last_error(None) or b'').decode())
        return bytes(ident)

    def allreduce_counts(self, counts):
        """In-place sum of an int64 count vector over the ranks of init_comm (identity without a communicator);
        a CUDA tensor stays on the device and the reduction is enqueued on the handle's stream."""
        arg = _Arg(counts, np.int64, writable=True)
        self._check(self._lib.aig_allreduce_counts(self._h, arg.ptr, int(np.prod(arg.shape))))
        return counts

    # -- tables -------------------------------------------------------------------------------
    def set_tables(self, filter_mat, dct_base, lifter, mfnorm):
        bank = np.ascontiguousarray(filter_mat, dtype=np.float64)
        dct = np.ascontiguousarray(dct_base, dtype=np.float64)
        lift = np.ascontiguousarray(lifter, dtype=np.float64)
        if bank.ndim != 2 or dct.ndim != 2 or dct.shape[0] != bank.shape[1] or lift.shape != (dct.shape[1],):
            raise ValueError('inconsistent table shapes %s %s %s' % (bank.shape, dct.shape, lift.shape))
        cur = self._tables_key          # the tables the handle holds: (bank, dct, lifter, mfnorm) copies
        if (cur is not None and float(mfnorm) == cur[3] and bank.shape == cur[0].shape and dct.shape == cur[1].shape
                and np.array_equal(bank, cur[0]) and np.array_equal(dct, cur[1]) and np.array_equal(lift, cur[2])):
            return                      # a memcmp-speed check: the reference's callers pass the same tables on every call
        self._check(self._lib.aig_set_tables(self._h, bank.ctypes.data, bank.shape[0], bank.shape[1],
                                             dct.ctypes.data, dct.shape[1], lift.ctypes.data, float(mfnorm)))
        self._tables_key = (bank.copy(), dct.copy(), lift.copy(), float(mfnorm))
        self.fft_len, self.filter_num, self.mfcc_num = bank.shape[0], bank.shape[1], dct.shape[1]

    @property
    def tables_are_reference(self):
        return self._lib.aig_tables_are_reference(self._h) == 1

    # -- stage 1 ------------------------------------------------------------------------------
    def mfcc_rows(self, beam, flip180=False, frame_pixels=FRAME_PIXELS, out=None):
        """[n, fft_len] float32 power rows -> [n, mfcc_num] float32 (get_feats + np.float32)."""
        a = self._a(beam, np.float32)
        n = int(np.prod(a.shape)) // self.fft_len
        if n * self.fft_len != int(np.prod(a.shape)):
            raise ValueError('beam has %s elements, not a multiple of fft_len=%d' % (a.shape, self.fft_len))
        res = out if out is not None else self._empty((n, self.mfcc_num), np.float32, a)
        o = self._a(res, np.float32, writable=True)
        if int(np.prod(o.shape)) != n * self.mfcc_num:
            raise ValueError('out has shape %s, expected %d x %d values' % (o.shape, n, self.mfcc_num))
        self._check(self._lib.aig_mfcc(self._h, a.ptr, n, o.ptr, int(bool(flip180)), int(frame_pixels)))
        return res

    def mfcc_image(self, power, flip=False):
        """[N, 36, 48, 512] float32 -> [N, 36, 48, 12] float32 MFCC acoustic images; ``flip`` applies the
        flip_left_right + flip_up_down of _parse_sequence (outdoor_data_mfcc.py:314-315)."""
        a = self._a(power, np.float32)
        if len(a.shape) != 4 or a.shape[1:] != (FRAME_H, FRAME_W, FFT_LEN):
            raise ValueError('expected [N, 36, 48, 512], got %s' % (a.shape,))
        rows = self.mfcc_rows(a.keep, flip180=flip, frame_pixels=FRAME_PIXELS)
        return rows.reshape(a.shape[0], FRAME_H, FRAME_W, MFCC_NUM)

    # -- callers either side of the path ----------------------------------
    @staticmethod
    def _audio_arg(audio):
        """Audio rows as a 4-byte buffer: int32 (tfrecord samples) stays int32, everything else becomes float32."""
        if _is_torch(audio):
            is_int = audio.dtype == _torch().int32
        else:
            is_int = np.asarray(audio).dtype == np.int32
        return _Arg(audio, np.int32 if is_int else np.float32), int(is_int)

    def power_spectrum(self, audio, window='tukey'):
        """[n, 1024] audio -> float32 [n, 512] power: window, 1024-point rFFT, drop Nyquist, |.|^2
        (outdoor_data_mfcc.py:799-804).  ``window``: 'tukey' (reference), None (frames.py variant) or a float64[1024]."""
        a, is_int = self._audio_arg(audio)
        n = self._frames(a.shape, 1024)
        win = tables.tukey_window() if isinstance(window, str) else window
        win = None if win is None else np.ascontiguousarray(win, dtype=np.float64)
        if win is not None and win.shape != (1024,):
            raise ValueError('window must have 1024 entries')
        res = self._empty((n, 512), np.float32, a)
        self._check(self._lib.aig_power_spectrum(self._h, a.ptr, is_int, n, None if win is None else win.ctypes.data,
                                                 self._a(res, np.float32, True).ptr))
        return res

    def build_spectrograms(self, audio, device_out=False):
        """_build_spectrograms_function (outdoor_data_mfcc.py:796-824): [n, 1024] audio -> float32 [n, 12] MFCC,
        spectrum and MFCC kernels chained on the device."""
        a, is_int = self._audio_arg(audio)
        n = self._frames(a.shape, 1024)
        win = tables.tukey_window()
        res = self._empty((n, self.mfcc_num), np.float32, a, device_out)
        self._check(self._lib.aig_audio_mfcc(self._h, a.ptr, is_int, n, win.ctypes.data, self._a(res, np.float32, True).ptr))
        return res

    def butter_lowpass_filter(self, data, cutoff=125, order=10, sample_rate=12288, device_out=False):
        """butter_lowpass_filter (outdoor_data_mfcc.py:571-575): zero-phase order-10 Butterworth low-pass along rows,
        float32 out."""
        a, is_int = self._audio_arg(data)
        if len(a.shape) < 1:
            raise ValueError('data must have at least one axis')
        length = a.shape[-1]
        n = int(np.prod(a.shape)) // length
        b, acoef, zi = tables.butter_lowpass(sample_rate, cutoff, order)
        res = self._empty(a.shape, np.float32, a, device_out)
        self._check(self._lib.aig_filtfilt(self._h, a.ptr, is_int, n, int(length), b.ctypes.data, acoef.ctypes.data,
                                           zi.ctypes.data, len(b), self._a(res, np.float32, True).ptr))
        return res

    def normalize_mfcc(self, mfcc):
        """Per-vector float32 min-max of [n, 12] MFCCs (_normalize_mfcc, outdoor_data_mfcc.py:696-703)."""
        a = self._a(mfcc, np.float32)
        n = self._frames(a.shape, MFCC_NUM)
        res = self._empty(a.shape, np.float32, a)
        self._check(self._lib.aig_normalize_mfcc(self._h, a.ptr, n, self._a(res, np.float32, True).ptr))
        return res

    def tile_mfcc(self, mfcc, normalize=False):
        """mfccmap: [B, 12] -> [B, 36, 48, 12] (trainer/mfcctrainer.py:38-40), optionally normalising each vector first."""
        a = self._a(mfcc, np.float32)
        n = self._frames(a.shape, MFCC_NUM)
        res = self._empty((n, FRAME_H, FRAME_W, MFCC_NUM), np.float32, a)
        self._check(self._lib.aig_tile_mfcc(self._h, a.ptr, n, int(bool(normalize)), self._a(res, np.float32, True).ptr))
        return res

    def split_triplets(self, images):
        """The four channel triplets tf.slice(x, [0,0,0,3t], [-1,36,48,3]) of [N, 36, 48, 12] images as one contiguous
        [4, N, 36, 48, 3] array (trainer/mfcctrainer.py:105-112)."""
        a = self._a(images, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * MFCC_NUM)
        res = self._empty((4, n, FRAME_H, FRAME_W, 3), np.float32, a)
        self._check(self._lib.aig_split_triplets(self._h, a.ptr, n, self._a(res, np.float32, True).ptr))
        return res

    def triplet_mse(self, target, generated):
        """float64[5]: tf.losses.mean_squared_error of the whole [N, 36, 48, 12] images and of each of their four channel
        triplets (trainer/mfcctrainer.py:103, 114-117), both images read once."""
        a, b = self._a(target, np.float32), self._a(generated, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * MFCC_NUM)
        if self._frames(b.shape, FRAME_PIXELS * MFCC_NUM) != n or n == 0:
            raise ValueError('triplet_mse needs two non-empty image batches of the same size')
        out = np.empty(5, np.float64)
        self._check(self._lib.aig_triplet_mse(self._h, a.ptr, b.ptr, n, out.ctypes.data))
        return out

    # -- stage 2 ------------------------------------------------------------------------------
    def normalize_images(self, images, device_out=False):
        """Per-frame (x - min) / max(x - min) in float32 over [N, 36, 48, 12] (outdoor_data_mfcc.py:672-679)."""
        a = self._a(images, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * MFCC_NUM)
        res = self._empty(a.shape, np.float32, a, device_out)
        self._check(self._lib.aig_normalize_images(self._h, a.ptr, n, self._a(res, np.float32, True).ptr))
        return res

    @staticmethod
    def _frames(shape, per_frame):
        total = int(np.prod(shape))
        if total % per_frame:
            raise ValueError('shape %s is not a whole number of frames of %d values' % (shape, per_frame))
        return total // per_frame

    def energy(self, images, normalize_first=False, want_scaled=False, want_mean=False):
        """find_logen + mean mask over a batch.  Returns (energy f64 [N,36,48], mask u8 [N,36,48][, scaled][, mean])."""
        a = self._a(images, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * MFCC_NUM)
        energy = self._empty((n, FRAME_H, FRAME_W), np.float64, a)
        mask = self._empty((n, FRAME_H, FRAME_W), np.uint8, a)
        scaled = self._empty((n, FRAME_H, FRAME_W, MFCC_NUM), np.float32, a) if want_scaled else None
        mean = self._empty((n,), np.float64, a) if want_mean else None
        self._check(self._lib.aig_energy(
            self._h, a.ptr, n, int(bool(normalize_first)),
            self._a(scaled, np.float32, True).ptr if want_scaled else None,
            self._a(energy, np.float64, True).ptr, self._a(mask, np.uint8, True).ptr,
            self._a(mean, np.float64, True).ptr if want_mean else None))
        out = (energy, mask)
        if want_scaled:
            out += (scaled,)
        if want_mean:
            out += (mean,)
        return out

    def find_logen(self, mfcc, inplace=True):
        """Drop-in for find_logen(mfcc) -> float64 [36, 48] (iouenergythreshold.py:294-323).

        Like the reference, the caller's contiguous float32 array is left scaled by
        1/lifter * mfnorm (``mfcc /= lifter; mfcc *= mfnorm`` act on a reshape view);
        ``inplace=False`` suppresses that side effect."""
        if isinstance(mfcc, np.ndarray) and mfcc.dtype == np.float32 and mfcc.flags.c_contiguous and inplace:
            if mfcc.size != FRAME_PIXELS * MFCC_NUM:
                raise ValueError('find_logen expects 36*48*12 values, got %d' % mfcc.size)
            energy = np.empty((FRAME_H, FRAME_W), np.float64)
            self._check(self._lib.aig_energy(self._h, mfcc.ctypes.data, 1, 0, mfcc.ctypes.data,
                                             energy.ctypes.data, None, None))
            return energy
        if _is_cuda_tensor(mfcc) and inplace and mfcc.is_contiguous() and mfcc.dtype == _torch().float32:
            energy = _torch().empty((FRAME_H, FRAME_W), dtype=_torch().float64, device=mfcc.device)
            self._check(self._lib.aig_energy(self._h, mfcc.data_ptr(), 1, 0, mfcc.data_ptr(),
                                             energy.data_ptr(), None, None))
            return energy
        energy, _ = self.energy(mfcc, normalize_first=False)
        return energy.reshape(FRAME_H, FRAME_W)

    def heatmap(self, energy, out_h=HEAT_H, out_w=HEAT_W):
        """cv2.resize(map, (out_w, out_h)) + imshow's implicit min/max normalisation, float32 [N, out_h, out_w]."""
        a = self._a(energy, np.float64)
        n = self._frames(a.shape, FRAME_PIXELS)
        res = self._empty((n, out_h, out_w), np.float32, a)
        self._check(self._lib.aig_heatmap(self._h, a.ptr, n, int(out_h), int(out_w), self._a(res, np.float32, True).ptr))
        return res

    def energy_heatmap(self, images, normalize_first=False, out_h=HEAT_H, out_w=HEAT_W, want_energy=True, want_mask=True,
                       out=None):
        """find_logen -> cv2.resize -> imshow normalisation for a batch in ONE kernel launch (showvideo.py:226-228):
        returns (energy f64 [N,36,48], mask u8 [N,36,48], heat f32 [N,out_h,out_w]); energy / mask are None when not
        wanted (they then never leave the SM).  ``out`` supplies the heat buffer."""
        a = self._a(images, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * MFCC_NUM)
        energy = self._empty((n, FRAME_H, FRAME_W), np.float64, a) if want_energy else None
        mask = self._empty((n, FRAME_H, FRAME_W), np.uint8, a) if want_mask else None
        heat = out if out is not None else self._empty((n, out_h, out_w), np.float32, a)
        o_heat = self._a(heat, np.float32, True)
        if int(np.prod(o_heat.shape)) != n * out_h * out_w:
            raise ValueError('out has shape %s, expected %d x %d x %d' % (o_heat.shape, n, out_h, out_w))
        self._check(self._lib.aig_energy_heatmap(self._h, a.ptr, n, int(bool(normalize_first)),
                                                 self._a(energy, np.float64, True).ptr if want_energy else None,
                                                 self._a(mask, np.uint8, True).ptr if want_mask else None,
                                                 o_heat.ptr, int(out_h), int(out_w)))
        return energy, mask, heat

    def overlay(self, heat, frames_bgr=None, alpha=0.7):
        """Jet-coloured heat map blended over the gray video frame (showvideo.py:224-229), RGB uint8 [N, H, W, 3].
        heat: float32 [N, H, W] in [0, 1] (from ``heatmap``); frames_bgr: uint8 [N, H, W, 3] (OpenCV order) or None."""
        a = self._a(heat, np.float32)
        if len(a.shape) != 3:
            raise ValueError('heat must be [N, H, W], got %s' % (a.shape,))
        n, hh, ww = a.shape
        b = None
        if frames_bgr is not None:
            b = self._a(frames_bgr, np.uint8)
            if tuple(b.shape) != (n, hh, ww, 3):
                raise ValueError('frames must be [N, H, W, 3] uint8 matching the heat maps')
        lut = tables.jet_lut()
        res = self._empty((n, hh, ww, 3), np.uint8, a)
        self._check(self._lib.aig_overlay(self._h, a.ptr, b.ptr if b is not None else None, n, hh, ww, float(alpha),
                                          lut.ctypes.data, self._a(res, np.uint8, True).ptr))
        return res

    def resize_mask(self, mask, out_h=HEAT_H, out_w=HEAT_W):
        """1.0 * (cv2.resize(mask * 1.0, (out_w, out_h)) > 0.5) as uint8 [N, out_h, out_w] (showimages_bb.py:303-304)."""
        a = self._a(mask, np.uint8)
        n = self._frames(a.shape, FRAME_PIXELS)
        res = self._empty((n, out_h, out_w), np.uint8, a)
        self._check(self._lib.aig_resize_mask(self._h, a.ptr, n, int(out_h), int(out_w), self._a(res, np.uint8, True).ptr))
        return res

    def mfcc_energy(self, power, flip=False, normalize_first=True, want_mean=False, out=None):
        """Stages 1 + 2 chained on the device: [N,36,48,512] f32 -> (mfcc f32 [N,36,48,12], energy f64, mask u8).

        ``out=(mfcc, energy, mask)`` supplies the result buffers (e.g. pinned host arrays, or device tensors
        reused across calls) instead of allocating them."""
        a = self._a(power, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * FFT_LEN)
        if out is not None:
            mfcc, energy, mask = out
        else:
            mfcc = self._empty((n, FRAME_H, FRAME_W, MFCC_NUM), np.float32, a)
            energy = self._empty((n, FRAME_H, FRAME_W), np.float64, a)
            mask = self._empty((n, FRAME_H, FRAME_W), np.uint8, a)
        o_mfcc, o_energy, o_mask = self._a(mfcc, np.float32, True), self._a(energy, np.float64, True), self._a(mask, np.uint8, True)
        if (int(np.prod(o_mfcc.shape)), int(np.prod(o_energy.shape)), int(np.prod(o_mask.shape))) != (
                n * FRAME_PIXELS * MFCC_NUM, n * FRAME_PIXELS, n * FRAME_PIXELS):
            raise ValueError('out buffers do not match %d frames' % n)
        mean = self._empty((n,), np.float64, a) if want_mean else None
        self._check(self._lib.aig_mfcc_energy(
            self._h, a.ptr, n, int(bool(flip)), int(bool(normalize_first)), o_mfcc.ptr, o_energy.ptr, o_mask.ptr,
            self._a(mean, np.float64, True).ptr if want_mean else None))
        return (mfcc, energy, mask) + ((mean,) if want_mean else ())

    def mfcc_energy_heatmap(self, power, flip=False, normalize_first=True, out_h=HEAT_H, out_w=HEAT_W, out=None):
        """Stages 1 + 2 including the up-sampled, normalised heat map in the one persistent kernel (opt-in,
        aig_mfcc_energy_heatmap): [N,36,48,512] f32 -> (mfcc f32 [N,36,48,12], energy f64, mask u8, heat f32 [N,out_h,out_w]).
        ``out=(mfcc, energy, mask, heat)`` supplies the result buffers."""
        a = self._a(power, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * FFT_LEN)
        if out is not None:
            mfcc, energy, mask, heat = out
        else:
            mfcc = self._empty((n, FRAME_H, FRAME_W, MFCC_NUM), np.float32, a)
            energy = self._empty((n, FRAME_H, FRAME_W), np.float64, a)
            mask = self._empty((n, FRAME_H, FRAME_W), np.uint8, a)
            heat = self._empty((n, out_h, out_w), np.float32, a)
        args = [self._a(mfcc, np.float32, True), self._a(energy, np.float64, True), self._a(mask, np.uint8, True),
                self._a(heat, np.float32, True)]
        sizes = (n * FRAME_PIXELS * MFCC_NUM, n * FRAME_PIXELS, n * FRAME_PIXELS, n * out_h * out_w)
        if tuple(int(np.prod(x.shape)) for x in args) != sizes:
            raise ValueError('out buffers do not match %d frames of %d x %d heat maps' % (n, out_h, out_w))
        self._check(self._lib.aig_mfcc_energy_heatmap(self._h, a.ptr, n, int(bool(flip)), int(bool(normalize_first)),
                                                      args[0].ptr, args[1].ptr, args[2].ptr, None, args[3].ptr,
                                                      int(out_h), int(out_w)))
        return mfcc, energy, mask, heat

    # -- stage 3 ------------------------------------------------------------------------------
    @staticmethod
    def _thresholds(thresholds):
        """Threshold vector as a float64 buffer: NumPy (host) or a CUDA tensor kept on the device."""
        if _is_torch(thresholds):
            arg = _Arg(thresholds, np.float64)
        else:
            arg = _Arg(np.ascontiguousarray(thresholds, dtype=np.float64), np.float64)
        if len(arg.shape) != 1:
            raise ValueError('thresholds must be a vector')
        return arg

    @staticmethod
    def _counters(thr, pos, num):
        """(pos, num) accumulators: NumPy by default, or caller-supplied int64 arrays / CUDA tensors
        (device-resident counters let a whole evaluation run without touching the host)."""
        if pos is None:
            pos = np.zeros(thr.shape[0], np.int64)
        elif not _is_torch(pos):
            pos = np.ascontiguousarray(pos, dtype=np.int64)
        if num is None or isinstance(num, (int, np.integer)):
            num = np.array([0 if num is None else int(num)], np.int64)
        p, c = _Arg(pos, np.int64, True), _Arg(num, np.int64, True)
        if int(np.prod(p.shape)) != thr.shape[0] or int(np.prod(c.shape)) != 1:
            raise ValueError('pos must hold one int64 per threshold and num exactly one int64')
        return pos, num, p, c

    @staticmethod
    def _num_result(num):
        return int(num[0]) if isinstance(num, np.ndarray) else num

    def iou_sweep(self, mask_a, mask_b, thresholds=REFERENCE_THRESHOLDS, pos=None, num=None):
        """Mask IoU and success counts (iouenergythreshold.py:224-229) for all thresholds at once.

        Returns (inter int64 [n], union int64 [n], pos int64 [K], num int).  ``pos`` / ``num`` from a
        previous call can be passed back in to accumulate across batches."""
        a, b = self._a(mask_a, np.uint8), self._a(mask_b, np.uint8)
        n = self._frames(a.shape, FRAME_PIXELS)
        if self._frames(b.shape, FRAME_PIXELS) != n:
            raise ValueError('mask batches differ: %s vs %s' % (a.shape, b.shape))
        thr = self._thresholds(thresholds)
        pos, cnt, p_arg, c_arg = self._counters(thr, pos, num)
        inter = self._empty((n,), np.int64, a)
        union = self._empty((n,), np.int64, a)
        self._check(self._lib.aig_iou_sweep(self._h, a.ptr, b.ptr, n, thr.ptr, thr.shape[0],
                                            self._a(inter, np.int64, True).ptr, self._a(union, np.int64, True).ptr,
                                            p_arg.ptr, c_arg.ptr))
        return inter, union, pos, self._num_result(cnt)

    def iou_sweep_clips(self, mask_a, mask_b, frames_per_clip, thresholds=REFERENCE_THRESHOLDS, pos=None):
        """Per-clip success counts for a stream of fixed-length clips in one launch: returns (inter [n], union [n],
        pos int64 [n_clips, K]); pos accumulates into a caller-supplied array / CUDA tensor when given."""
        a, b = self._a(mask_a, np.uint8), self._a(mask_b, np.uint8)
        n = self._frames(a.shape, FRAME_PIXELS)
        if self._frames(b.shape, FRAME_PIXELS) != n:
            raise ValueError('mask batches differ: %s vs %s' % (a.shape, b.shape))
        thr = self._thresholds(thresholds)
        n_clips = -(-n // int(frames_per_clip))
        if pos is None:
            pos = np.zeros((n_clips, thr.shape[0]), np.int64)
        p_arg = _Arg(pos, np.int64, True)
        if int(np.prod(p_arg.shape)) != n_clips * thr.shape[0]:
            raise ValueError('pos must be int64 [%d, %d]' % (n_clips, thr.shape[0]))
        inter = self._empty((n,), np.int64, a)
        union = self._empty((n,), np.int64, a)
        self._check(self._lib.aig_iou_sweep_clips(self._h, a.ptr, b.ptr, n, int(frames_per_clip), thr.ptr, thr.shape[0],
                                                  self._a(inter, np.int64, True).ptr, self._a(union, np.int64, True).ptr,
                                                  p_arg.ptr))
        return inter, union, pos

    def acivw_batch(self, real, reconstructed, thresholds=REFERENCE_THRESHOLDS, pos=None, num=None, normalize_first=False,
                    want_energy=False, want_masks=False):
        """The reference's ACIVW evaluation step for a batch in one kernel launch (iouenergythreshold.py:213-229): the real
        and the reconstructed [B, 36, 48, 12] image -> find_logen of each -> mean masks -> I, U -> success counts.

        Returns (inter int64 [B], union int64 [B], pos int64 [K], num) and, when asked for, (energy_real, energy_recon)
        float64 [B, 36, 48] and (mask_real, mask_recon) uint8 [B, 36, 48] appended.  Like the reference (which works on
        an np.stack copy) the inputs are not scaled in place."""
        a, b = self._a(real, np.float32), self._a(reconstructed, np.float32)
        n = self._frames(a.shape, FRAME_PIXELS * MFCC_NUM)
        if self._frames(b.shape, FRAME_PIXELS * MFCC_NUM) != n:
            raise ValueError('image batches differ: %s vs %s' % (a.shape, b.shape))
        thr = self._thresholds(thresholds)
        pos, cnt, p_arg, c_arg = self._counters(thr, pos, num)
        inter = self._empty((n,), np.int64, a)
        union = self._empty((n,), np.int64, a)
        energies = [self._empty((n, FRAME_H, FRAME_W), np.float64, a) for _ in range(2)] if want_energy else [None, None]
        masks = [self._empty((n, FRAME_H, FRAME_W), np.uint8, a) for _ in range(2)] if want_masks else [None, None]
        ptr = lambda x, dt: self._a(x, dt, True).ptr if x is not None else None
        self._check(self._lib.aig_acivw_batch(self._h, a.ptr, b.ptr, n, int(bool(normalize_first)), thr.ptr, thr.shape[0],
                                              ptr(inter, np.int64), ptr(union, np.int64), p_arg.ptr, c_arg.ptr,
                                              ptr(energies[0], np.float64), ptr(energies[1], np.float64),
                                              ptr(masks[0], np.uint8), ptr(masks[1], np.uint8)))
        out = (inter, union, pos, self._num_result(cnt))
        if want_energy:
            out += (tuple(energies),)
        if want_masks:
            out += (tuple(masks),)
        return out

    def ciou_sweep(self, mask, xmin, xmax, ymin, ymax, thresholds=REFERENCE_THRESHOLDS,
                   out_hw=(HEAT_H, HEAT_W), pos=None, num=None):
        """FlickrSoundNet consensus IoU and success counts (showimages_bb.py:288-321).

        mask u8 [n, 36, 48]; boxes int32 [n, 3] each.  Returns (2*I int64 [n], 2*U int64 [n], pos, num)."""
        a = self._a(mask, np.uint8)
        n = self._frames(a.shape, FRAME_PIXELS)
        boxes = [self._a(v, np.int32) for v in (xmin, xmax, ymin, ymax)]
        for bx in boxes:
            if int(np.prod(bx.shape)) != 3 * n:
                raise ValueError('boxes must be [n, 3] int32, got %s for n=%d' % (bx.shape, n))
        thr = self._thresholds(thresholds)
        pos, cnt, p_arg, c_arg = self._counters(thr, pos, num)
        inter2 = self._empty((n,), np.int64, a)
        union2 = self._empty((n,), np.int64, a)
        self._check(self._lib.aig_ciou_sweep(self._h, a.ptr, boxes[0].ptr, boxes[1].ptr, boxes[2].ptr, boxes[3].ptr,
                                             n, int(out_hw[0]), int(out_hw[1]), thr.ptr, thr.shape[0],
                                             self._a(inter2, np.int64, True).ptr, self._a(union2, np.int64, True).ptr,
                                             p_arg.ptr, c_arg.ptr))
        return inter2, union2, pos, self._num_result(cnt)


def auc(thresholds, values):
    """areaundercurve.py:32-37: sklearn.metrics.auc(threshold[::-1], value[::-1]) (trapezoid), float64."""
    thr = np.ascontiguousarray(thresholds, dtype=np.float64)
    val = np.ascontiguousarray(values, dtype=np.float64)
    if thr.shape != val.shape or thr.ndim != 1:
        raise ValueError('thresholds and values must be equal-length vectors')
    out = ctypes.c_double()
    code = _lib.load().aig_auc(thr.ctypes.data, val.ctypes.data, len(thr), ctypes.byref(out))
    if code != 0:
        raise AigError(code, 'aig_auc: thresholds must be monotonic with at least two entries')
    return out.value


def success_rates(pos, num):
    """``1.0 * pos / num`` (iouenergythreshold.py:236)."""
    return np.asarray(pos, dtype=np.float64) / np.float64(num)


def rates_as_written(rates):
    """The success rates as areaundercurve.py sees them: the evaluation scripts write each one as ``'iou {:6f}'``
    (iouenergythreshold.py:235-236) and areaundercurve.py:28-31 parses that text back, so its AUC is computed from values
    rounded to six decimals."""
    return np.array([float('{:6f}'.format(float(r))) for r in np.asarray(rates, dtype=np.float64)], dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# module-level drop-ins with the reference's names and signatures
# ---------------------------------------------------------------------------------------------
_local = threading.local()


def default_path(device=None):
    """The calling thread's handle (created on first use on the current CUDA device)."""
    path = getattr(_local, 'path', None)
    if device is None:
        if path is not None:
            return path
        try:
            device = _torch().cuda.current_device() if _torch().cuda.is_available() else 0
        except Exception:
            device = 0
    if path is None or path.device != device:
        path = AcousticPath(device)
        _local.path = path
    return path


createfilters = tables.createfilters


def get_feats(fft_len, beam, mfcc_num, dct_base, mfnorm, lifter, filter_mat):
    """Drop-in for get_feats (dataloader/outdoor_data_mfcc.py:851-876).

    NumPy in -> float64 [n, mfcc_num] out, like the reference (the values are the GPU's float32
    results; the reference's callers cast to float32 at :823 anyway).  CUDA tensor in -> float32
    CUDA tensor out."""
    path = default_path()
    bank = np.asarray(filter_mat)
    if bank.shape[0] != fft_len or np.shape(dct_base) != (bank.shape[1], mfcc_num):
        raise ValueError('table shapes do not match fft_len=%d mfcc_num=%d' % (fft_len, mfcc_num))
    path.set_tables(bank, dct_base, lifter, mfnorm)
    try:
        rows = path.mfcc_rows(beam)
    finally:
        path.set_tables(*tables.reference_tables())
    return rows.astype(np.float64) if isinstance(rows, np.ndarray) else rows


def _build_spectrograms_function(audio_data):
    """Drop-in for _build_spectrograms_function (outdoor_data_mfcc.py:796-824, iouenergythreshold.py:238-266)."""
    return default_path().build_spectrograms(audio_data)


def butter_lowpass_filter(data, cutoff=125, order=10, sample_rate=12288):
    """Drop-in for ActionsDataLoader.butter_lowpass_filter (outdoor_data_mfcc.py:571-575)."""
    return default_path().butter_lowpass_filter(data, cutoff, order, sample_rate)


def _map_func_audio_samples_build_spectrogram(audio_images, audio_wav, processed_images, action, location, filtered_audio_wav):
    """Drop-in for ActionsDataLoader._map_func_audio_samples_build_spectrogram (outdoor_data_mfcc.py:783-794): elements
    1 and 5 (the raw and the low-passed waveform, [T, 1024]) become their [T, 12] MFCCs."""
    path = default_path()
    return (audio_images, path.build_spectrograms(audio_wav), processed_images, action, location,
            path.build_spectrograms(filtered_audio_wav))


def _normalize_mfcc(mfcc):
    """Drop-in for ActionsDataLoader._normalize_mfcc (outdoor_data_mfcc.py:696-703) on one 12-vector."""
    return default_path().normalize_mfcc(mfcc).reshape(MFCC_NUM)


def _map_func_mfcc(audio_images, audio_samples, video_images, action, location, filtered_audio_samples):
    """Drop-in for ActionsDataLoader._map_func_mfcc (outdoor_data_mfcc.py:681-694): elements 1 and 5 ([T, 12] MFCCs of
    the raw and the low-passed audio) are normalised vector by vector."""
    path = default_path()
    return (audio_images, path.normalize_mfcc(audio_samples), video_images, action, location,
            path.normalize_mfcc(filtered_audio_samples))


def find_logen(mfcc):
    """Drop-in for find_logen (iouenergythreshold.py:294-323), including its in-place scaling of ``mfcc``."""
    return default_path().find_logen(mfcc, inplace=True)


def _normalize_acoustic_images_rescaled(image):
    """Drop-in for ActionsDataLoader._normalize_acoustic_images_rescaled (outdoor_data_mfcc.py:672-679): one
    [36, 48, 12] frame -> float32 (x - min) / max(x - min)."""
    out = default_path().normalize_images(image)
    return out.reshape(FRAME_H, FRAME_W, MFCC_NUM)


def _map_func_acoustic_images(audio_images, audio_samples, video_images, action, location, filtered_audio_samples):
    """Drop-in for ActionsDataLoader._map_func_acoustic_images (outdoor_data_mfcc.py:657-670): element 0 of the
    6-tuple ([T, 36, 48, 12] acoustic images) is normalised frame by frame, the rest pass through."""
    processed = default_path().normalize_images(audio_images)
    return processed, audio_samples, video_images, action, location, filtered_audio_samples
