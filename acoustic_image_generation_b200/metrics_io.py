"""The two plain-text metric files the evaluation scripts exchange.

``intersection_{tau}_accuracy.txt`` holds ``'iou {:6f}'.format(pos / num)`` for one IoU threshold
(iouenergythreshold.py:235-236, showimages_bb.py:327-328); ``area.txt`` holds
``'area {:6f}'.format(auc)`` (areaundercurve.py:39-40).  Writing them in the same format keeps the
reference's areaundercurve.py and meanstd.py usable on results produced here.
"""
from __future__ import annotations

import os


def accuracy_file_name(threshold):
    """File name the reference uses: the threshold is formatted as ``threshold * 1.0``."""
    return 'intersection_{}_accuracy.txt'.format(float(threshold) * 1.0)


def write_accuracy_file(data_dir, threshold, pos, num):
    path = os.path.join(data_dir, accuracy_file_name(threshold))
    with open(path, 'w') as outfile:
        outfile.write('iou {:6f}'.format(1.0 * int(pos) / int(num)))
    return path


def read_accuracy_file(data_dir, threshold):
    """The value areaundercurve.py:28-31 parses back: second space-separated token."""
    with open(os.path.join(data_dir, accuracy_file_name(threshold))) as infile:
        return float(infile.read().split(' ')[1])


def write_area_file(data_dir, auc_value):
    path = os.path.join(data_dir, 'area.txt')
    with open(path, 'w') as outfile:
        outfile.write('area {:6f}'.format(float(auc_value)))
    return path
