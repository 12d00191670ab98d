"""Generate tests/golden/*.npz by running the REFERENCE'S OWN SOURCE in this container.

Test tooling (see oracle/acoustic_oracle.py header).  Run once from the repo root:

    python oracle/make_golden.py [/root/reference]

The reference cannot be imported (every module imports TensorFlow 1.x, absent
here), but the hot path's arithmetic is plain NumPy: ``createfilters``,
``get_feats`` and ``find_logen`` are AST-extracted from
``iouenergythreshold.py`` (:269-292, :325-350, :294-323) and executed unmodified.
The inline scoring blocks (``iouenergythreshold.py:216-229``,
``showimages_bb.py:288-321``) are not functions in the reference; they are
replayed below call-for-call with the same NumPy / cv2 / sklearn calls and the
same dtypes.  TF-only ops (the 180-degree flip and the per-frame min-max,
``outdoor_data_mfcc.py:314-315,672-679``) have no runnable reference here; their
goldens come from float32 NumPy written to the TF op semantics.

Inputs are regenerated from seeds by ``acoustic_image_generation_b200.synth``;
each golden file stores a digest of its inputs so drift in the generator (e.g.
a different NumPy) is detected rather than mis-reported as a parity failure.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from acoustic_image_generation_b200 import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
REFERENCE_THRESHOLDS = [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]


def load_reference_functions(reference_root, rel='iouenergythreshold.py',
                             names=('createfilters', 'get_feats', 'find_logen')):
    """Compile the named top-level FunctionDefs of a reference file, unmodified."""
    with open(os.path.join(reference_root, rel)) as fh:
        tree = ast.parse(fh.read())
    scope = {'np': np}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), rel, 'exec'), scope)
    missing = [n for n in names if n not in scope]
    if missing:
        raise RuntimeError('reference functions not found: %s' % missing)
    return {n: scope[n] for n in names}


def reference_constants():
    """dct_base / lifter / mfnorm exactly as _build_spectrograms_function spells them
    (iouenergythreshold.py:248-263 == outdoor_data_mfcc.py:806-818)."""
    lifter_num, filter_num, mfcc_num = 22, 24, 12
    dct_base = np.zeros((filter_num, mfcc_num))
    for m in range(mfcc_num):
        dct_base[:, m] = np.cos((m + 1) * np.pi / filter_num * (np.arange(filter_num) + 0.5))
    lifter = 1 + (lifter_num / 2) * np.sin(np.pi * (1 + np.arange(mfcc_num)) / lifter_num)
    mfnorm = np.sqrt(2.0 / filter_num)
    return dct_base, lifter, mfnorm


def tf_minmax(frame):
    """float32 semantics of tf.reduce_min / subtract / tf.reduce_max / divide
    (outdoor_data_mfcc.py:672-679)."""
    x = np.asarray(frame, np.float32)
    x = x - np.min(x)
    with np.errstate(invalid='ignore', divide='ignore'):
        return x / np.max(x)


def replay_acivw_frame(ref, data_frame, recon_frame):
    """iouenergythreshold.py:216-226 for one frame (find_logen scales its argument in place,
    so it is handed copies, like the np.stack at :215 does)."""
    map1 = ref['find_logen'](np.array(data_frame, np.float32))
    m = 1 * (map1 > np.mean(map1))
    map2 = ref['find_logen'](np.array(recon_frame, np.float32))
    m2 = 1 * (map2 > np.mean(map2))
    intersection = np.logical_and(m, m2)
    union = np.logical_or(m, m2)
    with np.errstate(invalid='ignore', divide='ignore'):
        iou_score = np.sum(intersection) / np.sum(union)
    return int(np.sum(intersection)), int(np.sum(union)), float(iou_score)


def replay_flickr_frame(ref, cv2, recon_frame, xmin, xmax, ymin, ymax, out_w=298, out_h=224):
    """showimages_bb.py:288-318 for one frame."""
    m = np.zeros((3, out_h, out_w), dtype=np.float32)
    for contour in range(3):
        if xmax[contour] != 0:
            cv2.rectangle(m[contour], (int(xmin[contour]), int(ymin[contour])),
                          (int(xmax[contour]), int(ymax[contour])), (255, 255, 255), -1)
            m[contour] = m[contour] / 255.
            m[contour] = m[contour] / 2.
    mtot = np.sum(m, axis=0)
    mtot[mtot > 1.0] = 1.0
    map2 = ref['find_logen'](np.array(recon_frame, np.float32))
    m2 = 1 * (map2 > np.mean(map2))
    m2 = cv2.resize(m2 * 1.0, (out_w, out_h))
    m2 = 1.0 * (m2 > 0.5)
    intersection = np.logical_and(mtot, m2) * mtot
    union = np.logical_or(mtot, m2)
    box = 1 * (mtot > 0)
    unionbig = union + (mtot - box)
    with np.errstate(invalid='ignore', divide='ignore'):
        iou_score = np.sum(intersection) / np.sum(unionbig)
    return float(np.sum(intersection)), float(np.sum(unionbig)), float(iou_score), mtot, m2


def replay_flickr_mask(cv2, mask_small, xmin, xmax, ymin, ymax, out_w=298, out_h=224):
    """showimages_bb.py:288-296 and :303-318 for one frame whose mean-threshold mask `1 * (map2 > mean2)` is given
    (int64 [36, 48]) - the same statements and dtypes as replay_flickr_frame from `m2 = cv2.resize(...)` on."""
    m = np.zeros((3, out_h, out_w), dtype=np.float32)
    for contour in range(3):
        if xmax[contour] != 0:
            cv2.rectangle(m[contour], (int(xmin[contour]), int(ymin[contour])),
                          (int(xmax[contour]), int(ymax[contour])), (255, 255, 255), -1)
            m[contour] = m[contour] / 255.
            m[contour] = m[contour] / 2.
    mtot = np.sum(m, axis=0)
    mtot[mtot > 1.0] = 1.0
    m2 = np.asarray(mask_small, np.int64)
    m2 = cv2.resize(m2 * 1.0, (out_w, out_h))
    m2 = 1.0 * (m2 > 0.5)
    intersection = np.logical_and(mtot, m2) * mtot
    union = np.logical_or(mtot, m2)
    box = 1 * (mtot > 0)
    unionbig = union + (mtot - box)
    with np.errstate(invalid='ignore', divide='ignore'):
        iou_score = np.sum(intersection) / np.sum(unionbig)
    return np.sum(intersection), np.sum(unionbig), iou_score


def round2_goldens(cv2, metrics, acivw_pos11, n_pairs, flickr_pos11, n_fl):
    """Round-2 additions: (a) the AUC the way the reference's PIPELINE computes it - the evaluation scripts
    write each success rate as 'iou {:6f}' (iouenergythreshold.py:235-236), areaundercurve.py:28-37 parses the text back
    and integrates those six-decimal values; (b) consensus-IoU frames whose I / U is exactly 1/10 and 3/10: the dtype of
    the reference's ratio (float32 sum / float64 sum -> float64) decides whether `iou_score > 0.1` counts them."""
    import tempfile
    out = {}
    cases = (('acivw11', acivw_pos11, n_pairs), ('flickr11', flickr_pos11, n_fl),
             ('sevenths', np.array([7, 6, 5, 4, 3, 3, 2, 1, 1, 0, 0], np.int64), 7),
             ('thirds', np.array([3, 3, 2, 2, 2, 1, 1, 1, 0, 0, 0], np.int64), 3))
    for name, pos, num in cases:
        with tempfile.TemporaryDirectory() as data_dir:
            for threshold, p in zip(REFERENCE_THRESHOLDS, pos):                       # the writers, :235-236
                with open('{}'.format(data_dir) + "/intersection_{}_accuracy.txt".format(threshold * 1.0), "w") as outfile:
                    outfile.write('iou {:6f}'.format(1.0 * int(p) / num))
            value = np.zeros(11)                                                        # areaundercurve.py:26-37
            threshold = list(REFERENCE_THRESHOLDS)
            for i in range(len(threshold)):
                with open('{}'.format(data_dir) + "/intersection_{}_accuracy.txt".format(threshold[i]), "r") as outfile:
                    t = outfile.read()
                    value[i] = t.split(' ')[1]
            value = value[::-1]
            threshold = threshold[::-1]
            area = metrics.auc(threshold, value)
        out[name + '_pos'] = np.asarray(pos, np.int64)
        out[name + '_num'] = np.int64(num)
        out[name + '_auc_from_files'] = np.float64(area)
        out[name + '_area_text'] = np.array('area {:6f}'.format(area))
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'auc_files.npz'), **out)

    # (b) identity-size "up-sampling" (out 48 x 36), one annotator (weight 0.5) inside the predicted region:
    # I = 0.5 B, U = P - 0.5 B for a box of B pixels inside P predicted pixels
    masks = np.zeros((3, 36, 48), np.uint8)
    boxes = np.zeros((4, 3, 3), np.int32)                  # xmin, xmax, ymin, ymax  x  frame  x  annotator
    masks[0, 5, 10:21] = 1                                 # P = 11, box B = 2  ->  I / U = 1 / 10
    boxes[:, 0, 0] = (12, 13, 5, 5)
    masks[1, 8, 3:16] = 1                                  # P = 13, box B = 6  ->  3 / 10
    boxes[:, 1, 0] = (4, 9, 8, 8)
    masks[2, 20:24, 20:30] = 1                             # P = 40, two overlapping annotators -> generic control case
    boxes[:, 2, 0] = (18, 25, 19, 22)
    boxes[:, 2, 1] = (22, 33, 21, 26)
    rows = [replay_flickr_mask(cv2, masks[f], boxes[0, f], boxes[1, f], boxes[2, f], boxes[3, f], 48, 36) for f in range(3)]
    score = np.array([np.float64(r[2]) for r in rows])
    ratio = dict(masks=masks, xmin=boxes[0], xmax=boxes[1], ymin=boxes[2], ymax=boxes[3],
                 inter=np.array([np.float64(r[0]) for r in rows]), union=np.array([np.float64(r[1]) for r in rows]),
                 iou=score, ratio_dtype=np.array(str(np.asarray(rows[0][2]).dtype)),
                 counted=np.array([[bool(r[2] > t) for t in REFERENCE_THRESHOLDS] for r in rows]),
                 pos11=np.array([int(sum(bool(r[2] > t) for r in rows)) for t in REFERENCE_THRESHOLDS], np.int64))
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'ciou_ratio.npz'), **ratio)


def main(reference_root='/root/reference'):
    import cv2
    from sklearn import metrics

    ref = load_reference_functions(reference_root)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    dct_base, lifter, mfnorm = reference_constants()

    # ---- G1: filter bank and constants ------------------------------------------------
    bank = ref['createfilters'](512, 24, 0, 6400, 12800)
    bank_small = ref['createfilters'](256, 20, 300, 4000, 8000)   # a non-default geometry
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'filterbank.npz'),
                        filter_mat=bank, filter_mat_256_20=bank_small,
                        dct_base=dct_base, lifter=lifter, mfnorm=np.float64(mfnorm))

    # ---- G2: get_feats on pixel spectra ------------------------------------------------
    out = {}
    for kind, n, seed in (('chi2', 2, 0), ('lognormal', 1, 1), ('floor', 1, 2)):
        power = synth.power_frames(n, seed, kind)
        cep = ref['get_feats'](512, power.reshape(-1, 512), 12, dct_base, mfnorm, lifter, bank)
        out['mfcc_' + kind] = np.float32(cep).reshape(n, 36, 48, 12)     # outdoor_data_mfcc.py:823
        out['seed_' + kind] = np.int64(seed)
        out['digest_' + kind] = np.array(synth.digest(power))
    # a few rows kept in the reference's float64 for a tight oracle check
    out['mfcc_chi2_f64_rows'] = ref['get_feats'](
        512, synth.power_frames(2, 0, 'chi2').reshape(-1, 512)[:64], 12, dct_base, mfnorm, lifter, bank)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'mfcc.npz'), **out)

    # ---- G3: min-max, energy, mask -------------------------------------------------------
    mfcc_chi2 = out['mfcc_chi2']
    normed = np.stack([tf_minmax(f) for f in mfcc_chi2], 0)
    sig = synth.sigmoid_images(2, 3)
    smooth = synth.smooth_images(2, 4)
    energy = {}
    for name, imgs in (('normed', normed), ('sigmoid', sig), ('smooth', smooth)):
        maps = np.stack([ref['find_logen'](np.array(f, np.float32)) for f in imgs], 0)
        energy['energy_' + name] = maps
        energy['mean_' + name] = np.array([np.mean(mp) for mp in maps])
        energy['mask_' + name] = np.stack([(1 * (mp > np.mean(mp))).astype(np.uint8) for mp in maps], 0)
    # in-place side effect of find_logen: the argument after the call
    scaled = np.array(sig[0], np.float32)
    ref['find_logen'](scaled)
    energy['sigmoid0_after_call'] = scaled
    energy['normed_input'] = normed
    energy['digest_sigmoid'] = np.array(synth.digest(sig))
    energy['digest_smooth'] = np.array(synth.digest(smooth))
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'energy.npz'), **energy)

    # ---- G4: cv2.resize heat maps and mask resize ---------------------------------------
    heat = {}
    e0 = energy['energy_smooth'][0]
    heat['up_224_298'] = cv2.resize(e0, (298, 224))
    heat['up_224_224'] = cv2.resize(e0, (224, 224))
    mk = energy['mask_smooth']
    heat['mask_up_224_298'] = np.packbits(
        np.stack([(cv2.resize(m * 1.0, (298, 224)) > 0.5) for m in mk], 0).astype(np.uint8), axis=-1)
    heat['mask_up_224_224'] = np.packbits(
        np.stack([(cv2.resize(m * 1.0, (224, 224)) > 0.5) for m in mk], 0).astype(np.uint8), axis=-1)
    mk2 = energy['mask_sigmoid']
    heat['mask_sigmoid_up_224_298'] = np.packbits(
        np.stack([(cv2.resize(m * 1.0, (298, 224)) > 0.5) for m in mk2], 0).astype(np.uint8), axis=-1)
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'heatmap.npz'), **heat)

    # ---- G5: ACIVW IoU sweep -------------------------------------------------------------
    n_pairs = 48
    a = synth.smooth_images(n_pairs, 10)
    b = synth.smooth_images(n_pairs, 11)
    b[: n_pairs // 2] = a[: n_pairs // 2] * np.float32(0.9) + b[: n_pairs // 2] * np.float32(0.1)
    rows = [replay_acivw_frame(ref, a[h], b[h]) for h in range(n_pairs)]
    inter = np.array([r[0] for r in rows], np.int64)
    union = np.array([r[1] for r in rows], np.int64)
    score = np.array([r[2] for r in rows], np.float64)
    thr101 = np.linspace(0, 1, 101)
    acivw = dict(inter=inter, union=union, iou=score, digest_a=np.array(synth.digest(a)),
                 digest_b=np.array(synth.digest(b)),
                 pos11=np.array([int(np.sum(score > t)) for t in REFERENCE_THRESHOLDS], np.int64),
                 pos101=np.array([int(np.sum(score > t)) for t in thr101], np.int64),
                 num=np.int64(n_pairs))
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'acivw_iou.npz'), **acivw)

    # ---- G6: Flickr consensus IoU --------------------------------------------------------
    n_fl = 48
    pred = synth.smooth_images(n_fl, 20)
    xmin, xmax, ymin, ymax = synth.flickr_boxes(n_fl, 21)
    rows = [replay_flickr_frame(ref, cv2, pred[h], xmin[h], xmax[h], ymin[h], ymax[h]) for h in range(n_fl)]
    score = np.array([r[2] for r in rows], np.float64)
    flickr = dict(inter=np.array([r[0] for r in rows]), union=np.array([r[1] for r in rows]), iou=score,
                  gt0=rows[0][3], pred0=rows[0][4].astype(np.uint8),
                  digest_pred=np.array(synth.digest(pred)),
                  digest_boxes=np.array(synth.digest(np.stack([xmin, xmax, ymin, ymax]))),
                  pos11=np.array([int(np.sum(score > t)) for t in REFERENCE_THRESHOLDS], np.int64),
                  pos101=np.array([int(np.sum(score > t)) for t in thr101], np.int64),
                  num=np.int64(n_fl))
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'flickr_ciou.npz'), **flickr)

    # ---- G7: AUC (areaundercurve.py:32-37) -----------------------------------------------
    aucs = {}
    for name, pos, num, thr in (('acivw11', acivw['pos11'], n_pairs, REFERENCE_THRESHOLDS),
                                ('flickr11', flickr['pos11'], n_fl, REFERENCE_THRESHOLDS),
                                ('flickr101', flickr['pos101'], n_fl, list(thr101))):
        value = np.array([1.0 * p / num for p in pos])
        aucs[name] = np.float64(metrics.auc(list(thr)[::-1], value[::-1]))
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'auc.npz'), **aucs)
    round2_goldens(cv2, metrics, acivw['pos11'], n_pairs, flickr['pos11'], n_fl)

    # ---- G8: audio front half (N1): the reference's module-level _build_spectrograms_function and the loader's
    # butter_lowpass_filter, unmodified; `signal.tukey` moved to scipy.signal.windows in current scipy, so the
    # namespace they run in maps the old name to the new function -------------------------------------------------------
    import types
    from scipy import signal as sp_signal
    shim = types.SimpleNamespace(tukey=sp_signal.windows.tukey, butter=sp_signal.butter, filtfilt=sp_signal.filtfilt)
    with open(os.path.join(reference_root, 'iouenergythreshold.py')) as fh:
        tree = ast.parse(fh.read())
    scope = {'np': np, 'signal': shim, 'createfilters': ref['createfilters'], 'get_feats': ref['get_feats']}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == '_build_spectrograms_function':
            exec(compile(ast.Module([node], []), 'iouenergythreshold.py', 'exec'), scope)
    with open(os.path.join(reference_root, 'dataloader', 'outdoor_data_mfcc.py')) as fh:
        tree = ast.parse(fh.read())
    methods = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == 'ActionsDataLoader':
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name in ('butter_lowpass', 'butter_lowpass_filter'):
                    exec(compile(ast.Module([item], []), 'outdoor_data_mfcc.py', 'exec'), {'np': np, 'signal': shim}, methods)
    loader = types.SimpleNamespace(sample_rate=12288)
    loader.butter_lowpass = types.MethodType(methods['butter_lowpass'], loader)
    audio_i = synth.audio_rows(24, 90, np.int32)
    audio_f = synth.audio_rows(8, 91, np.float32, amplitude=1.0)
    front = dict(
        digest_int=np.array(synth.digest(audio_i)), digest_float=np.array(synth.digest(audio_f)),
        tukey=sp_signal.windows.tukey(1024, alpha=0.75),
        mfcc_int=scope['_build_spectrograms_function'](audio_i),
        mfcc_float=scope['_build_spectrograms_function'](audio_f),
        lowpass_int=methods['butter_lowpass_filter'](loader, audio_i),
        lowpass_float=methods['butter_lowpass_filter'](loader, audio_f),
        power_int_f64=(np.abs(np.fft.rfft(audio_i * sp_signal.windows.tukey(1024, alpha=0.75), 1024, axis=1))[:, :-1] ** 2)[:4])
    front['mfcc_lowpassed_int'] = scope['_build_spectrograms_function'](front['lowpass_int'])
    np.savez_compressed(os.path.join(GOLDEN_DIR, 'audio_front.npz'), **front)

    total = sum(os.path.getsize(os.path.join(GOLDEN_DIR, f)) for f in os.listdir(GOLDEN_DIR))
    print('golden vectors written to %s (%.1f KiB)' % (GOLDEN_DIR, total / 1024.0))


if __name__ == '__main__':
    main(*sys.argv[1:2])
