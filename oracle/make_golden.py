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
