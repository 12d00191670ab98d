"""CPU oracle for the acoustic-image front end and localisation scoring path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The product path (``acoustic_image_generation_b200``) never
does: it calls the CUDA kernels through ``libaig.so`` and fails loudly when the
library or a GPU is missing.

It is a NumPy restatement (written from scratch, float-for-float in the same
dtype and operation order) of the NumPy / OpenCV / sklearn / TF1 arithmetic the
reference IIT-PAVIS/Acoustic-Image-Generation runs on the host for this path.
Every function cites the reference file:line it follows (paths relative to the
reference checkout).

Parity pinning: the reference ships no tests and no golden vectors ("parity
unpinned" by the reference's own tests).  The oracle is instead pinned against
the reference's *own source* executed in the build container: ``oracle/
make_golden.py`` AST-extracts ``createfilters`` / ``get_feats`` / ``find_logen``
from ``/root/reference/iouenergythreshold.py`` and runs them (plus ``cv2.resize``,
``cv2.rectangle`` and ``sklearn.metrics.auc``, the third-party calls the
reference makes) on seeded inputs; the outputs are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` holds this restatement to
them (bit-exact wherever the arithmetic is deterministic NumPy; 1e-12 where the
reference goes through OpenCV's IPP resize whose rounding is not reproducible).

Third-party arithmetic on the path (not under /root/reference; the reference
pins no versions, README.md:7 only says "TensorFlow 1.14.0 >="): NumPy
``dot/exp/log/mean/sum`` (here 2.3.5), OpenCV ``resize(INTER_LINEAR)`` /
``rectangle`` (here 4.13.0), ``sklearn.metrics.auc`` (here 1.9.0).
"""
from __future__ import annotations

import numpy as np

# Reference geometry / constants -------------------------------------------------
FRAME_H = 36          # dataloader/outdoor_data_mfcc.py:445  reshape [-1, 36, 48, 12]
FRAME_W = 48
FRAME_PIXELS = FRAME_H * FRAME_W
FFT_LEN = 512         # dataloader/outdoor_data_mfcc.py:811
FILTER_NUM = 24       # :809
MFCC_NUM = 12         # :810
LIFTER_NUM = 22       # :806
LO_FREQ = 0           # :807
HI_FREQ = 6400        # :808
MEL_FLOOR = 0.001     # :858
HEAT_H = 224          # showimages.py:147  cv2.resize(map, (298, 224))
HEAT_W = 298
REFERENCE_THRESHOLDS = (0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0)  # areaundercurve.py:27


# ---------------------------------------------------------------------------
# F1  mel filter bank                     dataloader/outdoor_data_mfcc.py:826-849
# ---------------------------------------------------------------------------
def mel_band_edges(fft_len, filter_num, lo_freq, hi_freq, samp_freq):
    """FFT-bin index of the filter_num+2 triangle corners (outdoor_data_mfcc.py:830-840)."""
    to_mel = lambda hz: 1127 * (np.log(1 + (hz / 700.0)))
    to_hz = lambda mel: 700.0 * (np.exp(mel / 1127.0) - 1)
    centres_mel = np.linspace(to_mel(lo_freq), to_mel(hi_freq), filter_num + 2)
    centres_hz = to_hz(centres_mel)
    bins = centres_hz / float(samp_freq) * (fft_len - 1) * 2
    return np.floor(bins).astype('int')


def createfilters(fft_len, filter_num, lo_freq, hi_freq, samp_freq):
    """HTK-style triangular mel filter bank, float64 [fft_len, filter_num].

    Follows outdoor_data_mfcc.py:826-849 (identical copies: iouenergythreshold.py:269-292,
    showimages.py:191-214, ...): column f rises 0->1 over bins [p_f, p_{f+1}] and
    falls 1->0 over [p_{f+1}, p_{f+2}], both ramps written with np.linspace so the
    weights are bit-identical to the reference's.
    """
    edges = mel_band_edges(fft_len, filter_num, lo_freq, hi_freq, samp_freq)
    bank = np.zeros((fft_len, filter_num))
    for f in range(filter_num):
        left, peak, right = int(edges[f]), int(edges[f + 1]), int(edges[f + 2])
        bank[left:peak + 1, f] = np.linspace(0, 1, peak - left + 1)
        bank[peak:right + 1, f] = np.linspace(1, 0, right - peak + 1)
    return bank


def mfcc_constants(filter_num=FILTER_NUM, mfcc_num=MFCC_NUM, lifter_num=LIFTER_NUM):
    """(dct_base [filter_num, mfcc_num], lifter [mfcc_num], mfnorm) as float64.

    outdoor_data_mfcc.py:813-818 (and find_logen, iouenergythreshold.py:304-308):
    dct_base[j, m] = cos((m+1)*pi/filter_num*(j+0.5)); lifter[m] = 1+(L/2)*sin(pi*(m+1)/L);
    mfnorm = sqrt(2/filter_num).
    """
    j = np.arange(filter_num) + 0.5
    dct_base = np.zeros((filter_num, mfcc_num))
    for m in range(mfcc_num):
        dct_base[:, m] = np.cos((m + 1) * np.pi / filter_num * j)
    lifter = 1 + (lifter_num / 2) * np.sin(np.pi * (1 + np.arange(mfcc_num)) / lifter_num)
    mfnorm = np.sqrt(2.0 / filter_num)
    return dct_base, lifter, mfnorm


def reference_tables():
    """The fixed tables every reference caller builds (outdoor_data_mfcc.py:806-820)."""
    bank = createfilters(FFT_LEN, FILTER_NUM, LO_FREQ, HI_FREQ, 2 * HI_FREQ)
    dct_base, lifter, mfnorm = mfcc_constants()
    return bank, dct_base, lifter, mfnorm


# ---------------------------------------------------------------------------
# F2  spectrum rows -> MFCC rows          dataloader/outdoor_data_mfcc.py:851-876
# ---------------------------------------------------------------------------
def get_feats(fft_len, beam, mfcc_num, dct_base, mfnorm, lifter, filter_mat):
    """[n, fft_len] power rows -> [n, mfcc_num] float64 cepstra.

    Same operation order as outdoor_data_mfcc.py:851-876: filter-bank product
    (float32 rows x float64 bank => float64), floor at 0.001, natural log, DCT
    product, scale by mfnorm, scale by lifter, NaN and Inf replaced by 0.
    """
    rows = beam.shape[0]
    spectrum = np.reshape(beam, [rows, fft_len])
    mel = np.dot(spectrum, filter_mat)                 # :855
    np.copyto(mel, MEL_FLOOR, where=mel < MEL_FLOOR)   # :858
    mel = np.log(mel)                                  # :861
    cep = np.dot(mel, dct_base)                        # :864
    cep *= mfnorm                                      # :865
    cep *= lifter                                      # :868
    cep[np.isnan(cep)] = 0                             # :871
    cep[np.isinf(cep)] = 0                             # :872
    return np.reshape(cep, [rows, mfcc_num])


# ---------------------------------------------------------------------------
# F3  layout ops                          dataloader/outdoor_data_mfcc.py:314-315
# ---------------------------------------------------------------------------
def flip180(images):
    """tf.image.flip_left_right then flip_up_down on [T, H, W, C] (:314-315)."""
    return np.ascontiguousarray(images[:, ::-1, ::-1, :])


def mfcc_image(power, flip=False):
    """North-star stage 1: [N, 36, 48, 512] float32 power -> [N, 36, 48, 12] float32.

    The per-row operator is get_feats with the reference tables
    (outdoor_data_mfcc.py:820-823, result cast with np.float32 at :823) applied
    to the N*1728 pixel spectra; ``flip`` adds the 180-degree rotation of
    _parse_sequence (:314-315).
    """
    power = np.asarray(power, dtype=np.float32)
    n = power.shape[0]
    bank, dct_base, lifter, mfnorm = reference_tables()
    cep = get_feats(FFT_LEN, power.reshape(-1, FFT_LEN), MFCC_NUM, dct_base, mfnorm, lifter, bank)
    img = np.float32(cep).reshape(n, FRAME_H, FRAME_W, MFCC_NUM)
    return flip180(img) if flip else img


# ---------------------------------------------------------------------------
# F4  per-frame min-max                   dataloader/outdoor_data_mfcc.py:672-679
# ---------------------------------------------------------------------------
def normalize_acoustic_image(image):
    """float32 (x - min) / max(x - min) over all 36*48*12 values of one frame.

    Order as in _normalize_acoustic_images_rescaled (:674-678): subtract the
    minimum, then divide by the maximum of the *shifted* image.  A constant
    frame gives 0/0 = NaN, as in the reference.
    """
    x = np.asarray(image, dtype=np.float32)
    shifted = x - x.min()
    with np.errstate(invalid='ignore', divide='ignore'):
        return shifted / shifted.max()


def normalize_acoustic_images(images):
    """tf.map_fn of the above over the leading axis (:668)."""
    return np.stack([normalize_acoustic_image(f) for f in images], 0)


# ---------------------------------------------------------------------------
# F5  energy map                          iouenergythreshold.py:294-323
# ---------------------------------------------------------------------------
def find_logen(mfcc, inplace=True):
    """12-channel MFCC frame -> float64 [36, 48] energy map.

    iouenergythreshold.py:294-323.  Faithful details: the cepstra are divided by
    the lifter and *multiplied* by mfnorm in place (:310-311; each op computed in
    float64 and rounded back to the array's float32 on store - NumPy 2
    semantics), projected with dct_base transposed (:312-313, not a
    pseudo-inverse), exponentiated, summed over the 24 bands (:320) and
    inverted (:321).  With ``inplace`` (the reference behaviour) the caller's
    float32 buffer keeps the scaled values when it is a contiguous view.
    """
    dct_base, lifter, mfnorm = mfcc_constants()
    cep = np.reshape(mfcc, (-1, MFCC_NUM))
    if not inplace:
        cep = cep.copy()
    cep /= lifter[None, :]
    cep *= mfnorm
    bands = np.exp(np.dot(cep, dct_base.T))
    return np.reshape(1 / np.sum(bands, -1), (FRAME_H, FRAME_W))


# ---------------------------------------------------------------------------
# F7  mask                                iouenergythreshold.py:217-219
# ---------------------------------------------------------------------------
def mean_mask(energy):
    """1 * (map > np.mean(map)) as uint8 (iouenergythreshold.py:217-219)."""
    return (energy > np.mean(energy)).astype(np.uint8)


# ---------------------------------------------------------------------------
# F6  bilinear up-sampling + implicit normalise   showimages.py:146-148
# ---------------------------------------------------------------------------
def _linear_taps(n_src, n_dst):
    """Half-pixel-centre taps of cv2.resize(INTER_LINEAR): src = (dst+0.5)*n_src/n_dst-0.5,
    clamped at both borders.  Returns (i0, i1, w1) with float64 weights."""
    pos = (np.arange(n_dst) + 0.5) * (n_src / n_dst) - 0.5
    i0 = np.floor(pos).astype(np.int64)
    w1 = pos - i0
    low, high = i0 < 0, i0 >= n_src - 1
    i0 = np.where(low, 0, np.where(high, n_src - 1, i0))
    w1 = np.where(low | high, 0.0, w1)
    i1 = np.minimum(i0 + 1, n_src - 1)
    return i0, i1, w1


def resize_bilinear(image, out_h, out_w):
    """cv2.resize(image, (out_w, out_h)) for a float64 2-D array (showimages.py:147,
    showvideo.py:227).  Horizontal pass then vertical pass in float64.  OpenCV
    routes this call through IPP whose last-bit rounding is not reproducible, so
    this is pinned to cv2 at 1e-12, not bit-exactly."""
    src = np.asarray(image, dtype=np.float64)
    x0, x1, wx = _linear_taps(src.shape[1], out_w)
    y0, y1, wy = _linear_taps(src.shape[0], out_h)
    rows = src[:, x0] * (1 - wx)[None, :] + src[:, x1] * wx[None, :]
    return rows[y0, :] * (1 - wy)[:, None] + rows[y1, :] * wy[:, None]


def normalize_heatmap(up):
    """matplotlib Normalize(vmin=min, vmax=max) that imshow applies implicitly
    (showimages.py:148, showvideo.py:228): (x - min) / (max - min) over the
    up-sampled image."""
    lo, hi = up.min(), up.max()
    with np.errstate(invalid='ignore', divide='ignore'):
        return (up - lo) / (hi - lo)


def heatmap(energy, out_h=HEAT_H, out_w=HEAT_W):
    """energy [36,48] f64 -> normalised up-sampled heat map, float32 [out_h, out_w]."""
    return np.float32(normalize_heatmap(resize_bilinear(energy, out_h, out_w)))


def _linear_taps_exact(n_src, n_dst):
    """Integer form of _linear_taps: src = ((2d+1)*n_src - n_dst) / (2*n_dst).
    Returns (i0, i1, numerator of w1, denominator)."""
    den = 2 * n_dst
    t = (2 * np.arange(n_dst, dtype=np.int64) + 1) * n_src - n_dst
    i0 = np.floor_divide(t, den)
    num = t - i0 * den
    low, high = i0 < 0, i0 >= n_src - 1
    i0 = np.where(low, 0, np.where(high, n_src - 1, i0))
    num = np.where(low | high, 0, num)
    i1 = np.minimum(i0 + 1, n_src - 1)
    return i0, i1, num, den


def resize_mask(mask, out_h=HEAT_H, out_w=HEAT_W):
    """1.0 * (cv2.resize(m2 * 1.0, (out_w, out_h)) > 0.5) as uint8 (showimages_bb.py:303-304).

    Done in exact integer arithmetic: the bilinear value of a {0,1} image is a
    rational with denominator 4*out_h*out_w, so the strict ``> 0.5`` (including
    the exact-0.5 ties at e.g. output columns 74 and 223 for 48->298) is decided
    without rounding.  Checked against cv2 on 3.5e8 pixels in this container."""
    m = np.asarray(mask).astype(np.int64)
    x0, x1, xn, xd = _linear_taps_exact(m.shape[1], out_w)
    y0, y1, yn, yd = _linear_taps_exact(m.shape[0], out_h)
    rows = m[:, x0] * (xd - xn)[None, :] + m[:, x1] * xn[None, :]
    val = rows[y0, :] * (yd - yn)[:, None] + rows[y1, :] * yn[:, None]
    return (2 * val > xd * yd).astype(np.uint8)


# ---------------------------------------------------------------------------
# F8  ACIVW / AVIA IoU                    iouenergythreshold.py:216-229
# ---------------------------------------------------------------------------
def iou_pair(mask_a, mask_b):
    """(I, U, iou) of two {0,1} masks: I = sum(a and b), U = sum(a or b) as integers,
    iou = I / U in float64 (U == 0 gives NaN, which never counts as a success)."""
    a = np.asarray(mask_a) != 0
    b = np.asarray(mask_b) != 0
    inter = int(np.sum(np.logical_and(a, b)))
    union = int(np.sum(np.logical_or(a, b)))
    with np.errstate(invalid='ignore', divide='ignore'):
        score = np.float64(inter) / np.float64(union)
    return inter, union, score


# ---------------------------------------------------------------------------
# F9  FlickrSoundNet consensus IoU        showimages_bb.py:288-321
# ---------------------------------------------------------------------------
def boxes_to_consensus(xmin, xmax, ymin, ymax, out_h=HEAT_H, out_w=HEAT_W):
    """Weighted ground-truth map in {0, 0.5, 1} (float32 [out_h, out_w]).

    showimages_bb.py:288-296: up to three annotator boxes, a box is present iff
    xmax != 0, drawn filled with both corners inclusive and clipped to the image
    (cv2.rectangle, thickness -1), weight 0.5 each, sum capped at 1."""
    total = np.zeros((out_h, out_w), dtype=np.float32)
    for c in range(len(xmax)):
        if xmax[c] != 0:
            xa, xb = sorted((int(xmin[c]), int(xmax[c])))
            ya, yb = sorted((int(ymin[c]), int(ymax[c])))
            xa, ya = max(xa, 0), max(ya, 0)
            xb, yb = min(xb, out_w - 1), min(yb, out_h - 1)
            if xa <= xb and ya <= yb:
                total[ya:yb + 1, xa:xb + 1] += np.float32(0.5)
    return np.minimum(total, np.float32(1.0))


def consensus_iou(gt, pred_mask):
    """(2*I, 2*U, iou) for the consensus IoU of showimages_bb.py:306-318.

    I = sum((gt>0 and pred) * gt); U = sum((gt>0 or pred) + (gt - [gt>0])).
    Both are multiples of 0.5, returned doubled as exact integers; iou = I / U
    in float64."""
    g2 = np.rint(np.asarray(gt, dtype=np.float64) * 2).astype(np.int64)   # 0, 1, 2 half-units
    p = np.asarray(pred_mask) != 0
    box = g2 > 0
    inter2 = int(np.sum(np.where(box & p, g2, 0)))
    union2 = int(np.sum(2 * (box | p).astype(np.int64) + (g2 - 2 * box.astype(np.int64))))
    with np.errstate(invalid='ignore', divide='ignore'):
        score = np.float64(inter2) / np.float64(union2)
    return inter2, union2, score


# ---------------------------------------------------------------------------
# success-rate sweep and AUC              iouenergythreshold.py:227-236, areaundercurve.py:26-40
# ---------------------------------------------------------------------------
def success_counts(scores, thresholds):
    """pos[k] = #frames with iou > thresholds[k] (strict, NaN never counts); num = #frames.

    The reference re-runs the whole evaluation once per threshold
    (scripts/iou.bash:47-53, ``if iou_score > threshold: pos += 1; num += 1``);
    this computes all K counts in one pass."""
    s = np.asarray(scores, dtype=np.float64)
    t = np.asarray(thresholds, dtype=np.float64)
    with np.errstate(invalid='ignore'):
        pos = (s[None, :] > t[:, None]).sum(1).astype(np.int64)
    return pos, int(s.size)


def success_rates(pos, num):
    """1.0 * pos / num (iouenergythreshold.py:236)."""
    return np.asarray(pos, dtype=np.float64) / np.float64(num)


def rates_as_written(rates):
    """The success rates as areaundercurve.py:28-31 reads them: each was written as 'iou {:6f}'
    (iouenergythreshold.py:235-236) and is parsed back from that text, i.e. rounded to six decimals."""
    return np.array([float('iou {:6f}'.format(float(r)).split(' ')[1]) for r in np.asarray(rates, dtype=np.float64)])


def auc(thresholds, values):
    """sklearn.metrics.auc on the *reversed* arrays, as areaundercurve.py:32-37 does:
    x decreasing => direction -1 times the trapezoid sum of diff(x)*(y[1:]+y[:-1])/2."""
    x = np.asarray(thresholds, dtype=np.float64)[::-1]
    y = np.asarray(values, dtype=np.float64)[::-1]
    dx = np.diff(x)
    direction = 1.0
    if np.any(dx < 0):
        if np.all(dx <= 0):
            direction = -1.0
        else:
            raise ValueError('thresholds must be monotonic')
    return float(direction * (dx * (y[1:] + y[:-1]) / 2.0).sum())


# ---------------------------------------------------------------------------
# N1  audio front half                    dataloader/outdoor_data_mfcc.py:565-575, 796-805
# ---------------------------------------------------------------------------
AUDIO_SAMPLES = 1024      # _NUMBER_OF_SAMPLES, outdoor_data_mfcc.py:9
SAMPLE_RATE = 12288       # ActionsDataLoader.__init__ default, outdoor_data_mfcc.py:17


def tukey_window(m=AUDIO_SAMPLES, alpha=0.75):
    """scipy.signal.tukey(1024, alpha=0.75) (outdoor_data_mfcc.py:799; moved to scipy.signal.windows.tukey):
    cosine tapers over the first and last alpha/2 of the window, ones in between (symmetric form)."""
    n = np.arange(0, m)
    width = int(np.floor(alpha * (m - 1) / 2.0))
    n1, n3 = n[0:width + 1], n[m - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (m - 1))))
    w2 = np.ones(m - 2 * (width + 1))
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (m - 1))))
    return np.concatenate((w1, w2, w3))


def power_spectrum(audio, window=None):
    """[n, 1024] audio rows -> float64 [n, 512] power (outdoor_data_mfcc.py:800-804): window, rfft(1024), drop the
    Nyquist bin, |.|**2.  ``window=None`` is the variant of dataloader/frames.py:659-667 (no window)."""
    x = np.asarray(audio)
    if window is not None:
        x = x * np.reshape(np.tile(window, (x.shape[0], 1)), (x.shape[0], AUDIO_SAMPLES))
    else:
        # without the float64 window the dtype of the input decides the FFT precision: NumPy < 2 (the reference's
        # era) always transformed in float64, NumPy >= 2 keeps float32.  The oracle pins the float64 behaviour.
        x = x.astype(np.float64)
    spec = np.abs(np.fft.rfft(x, AUDIO_SAMPLES, axis=1))[:, :-1]
    return spec ** 2


def build_spectrograms(audio):
    """_build_spectrograms_function (outdoor_data_mfcc.py:796-824): [n, 1024] audio -> float32 [n, 12] MFCC."""
    bank, dct_base, lifter, mfnorm = reference_tables()
    power = power_spectrum(audio, tukey_window())
    return np.float32(get_feats(FFT_LEN, power, MFCC_NUM, dct_base, mfnorm, lifter, bank))


def butter_lowpass_filter(data, sample_rate=SAMPLE_RATE, cutoff=125, order=10):
    """butter_lowpass + butter_lowpass_filter (outdoor_data_mfcc.py:565-575): order-10 Butterworth low-pass at
    cutoff / (sample_rate / 2), zero-phase with scipy.signal.filtfilt, cast to float32.  scipy is the third-party
    arithmetic the reference itself calls here (unpinned there; 1.18.1 in this image)."""
    from scipy import signal
    b, a = signal.butter(order, cutoff / (0.5 * sample_rate), btype='low', analog=False)
    return np.float32(signal.filtfilt(b, a, data))


# ---------------------------------------------------------------------------
# N2  batch assembly for the models       outdoor_data_mfcc.py:681-703, trainer/mfcctrainer.py:38-40
# ---------------------------------------------------------------------------
def normalize_mfcc(mfcc):
    """_normalize_mfcc over rows: float32 (x - min) / max(x - min) per 12-vector (outdoor_data_mfcc.py:696-703)."""
    x = np.asarray(mfcc, dtype=np.float32).reshape(-1, MFCC_NUM)
    shifted = x - x.min(axis=1, keepdims=True)
    with np.errstate(invalid='ignore', divide='ignore'):
        return shifted / shifted.max(axis=1, keepdims=True)


def tile_mfcc(mfcc):
    """mfccmap (trainer/mfcctrainer.py:38-40): [B, 12] -> [B, 36, 48, 12], every pixel holds the clip's MFCC vector."""
    x = np.asarray(mfcc, dtype=np.float32).reshape(-1, 1, MFCC_NUM)
    return np.reshape(np.tile(x, (1, FRAME_PIXELS, 1)), (-1, FRAME_H, FRAME_W, MFCC_NUM))


def split_triplets(images):
    """tf.slice(x, [0,0,0,3t], [-1,36,48,3]) for t = 0..3 (trainer/mfcctrainer.py:105-112): [4, N, 36, 48, 3]."""
    x = np.asarray(images, dtype=np.float32).reshape(-1, FRAME_H, FRAME_W, MFCC_NUM)
    return np.stack([np.ascontiguousarray(x[..., 3 * t:3 * t + 3]) for t in range(4)])


def triplet_mse(target, generated):
    """tf.losses.mean_squared_error of the whole image and of each channel-triplet pair (trainer/mfcctrainer.py:103,
    114-117): float32 squared differences, averaged (here in float64; TF's float32 mean agrees to ~1e-7 relative)."""
    a, b = split_triplets(target), split_triplets(generated)
    sq = np.square(a - b)                                   # float32, like tf.squared_difference
    per = sq.reshape(4, -1).sum(axis=1, dtype=np.float64)
    return np.concatenate([[per.sum() / sq.size], per / (sq.size // 4)])


# ---------------------------------------------------------------------------
# N4  heat-map overlay                    showvideo.py:217-233, showimages.py:144-150
# ---------------------------------------------------------------------------
def overlay(heat, frame_bgr, lut, alpha=0.7):
    """One frame: gray = cv2.cvtColor(frame, COLOR_BGR2GRAY) (restated: OpenCV 4.x's 15-bit fixed point, checked against cv2
    in the tests); imshow(gray, cmap=gray) = min/max normalise + 256 levels; imshow(map, cmap=jet, alpha) = jet table
    lookup of the normalised heat map, blended alpha * jet + (1 - alpha) * gray in float32, rounded to uint8.
    PARITY UNPINNED against matplotlib (absent here; its figure resampling and Agg compositing are not reproduced)."""
    heat = np.asarray(heat, dtype=np.float32)
    ji = np.clip((heat * np.float32(256.0)).astype(np.int64), 0, 255)
    rgb = lut[ji].astype(np.float32)
    if frame_bgr is not None:
        f = np.asarray(frame_bgr).astype(np.int64)
        gray = (f[..., 0] * 3735 + f[..., 1] * 19235 + f[..., 2] * 9798 + 16384) >> 15
        lo, hi = gray.min(), gray.max()
        if hi > lo:
            gi = np.minimum(((gray - lo).astype(np.float32) / np.float32(hi - lo) * np.float32(256.0)).astype(np.int64), 255)
        else:
            gi = np.zeros_like(gray)
        a = np.float32(alpha)
        rgb = rgb * a + gi.astype(np.float32)[..., None] * (np.float32(1.0) - a)
    return np.rint(rgb).astype(np.uint8)


# ---------------------------------------------------------------------------
# whole-path composites used by the tests and the CPU baseline
# ---------------------------------------------------------------------------
def energy_stage(mfcc_images, normalize_first=True):
    """F4 -> F5 -> F7 for a batch: returns (energy f64 [N,36,48], mask u8 [N,36,48])."""
    energies, masks = [], []
    for frame in np.asarray(mfcc_images, dtype=np.float32):
        img = normalize_acoustic_image(frame) if normalize_first else frame.copy()
        e = find_logen(img)
        energies.append(e)
        masks.append(mean_mask(e))
    return np.stack(energies, 0), np.stack(masks, 0)


def acivw_sweep(energy_a, energy_b, thresholds):
    """iouenergythreshold.py:213-229 for a batch of energy-map pairs."""
    inter, union, score = [], [], []
    for ea, eb in zip(energy_a, energy_b):
        i, u, s = iou_pair(mean_mask(ea), mean_mask(eb))
        inter.append(i); union.append(u); score.append(s)
    pos, num = success_counts(score, thresholds)
    return np.array(inter, np.int64), np.array(union, np.int64), pos, num


def flickr_sweep(masks, xmin, xmax, ymin, ymax, thresholds, out_h=HEAT_H, out_w=HEAT_W):
    """showimages_bb.py:287-321 for a batch: masks u8 [N,36,48], boxes int32 [N,3] each."""
    inter2, union2, score = [], [], []
    for h in range(len(masks)):
        gt = boxes_to_consensus(xmin[h], xmax[h], ymin[h], ymax[h], out_h, out_w)
        i2, u2, s = consensus_iou(gt, resize_mask(masks[h], out_h, out_w))
        inter2.append(i2); union2.append(u2); score.append(s)
    pos, num = success_counts(score, thresholds)
    return np.array(inter2, np.int64), np.array(union2, np.int64), pos, num
