"""ORACLE — test infrastructure, NOT product code.

numpy restatement of the reference's confusion-matrix metrics (misc/metric_tool.py) used to check
dahitra_b200.metrics (device-side confusion matrix).  Integer work: parity is bit-exact for the
matrix; the derived scores are float64 formulas restated term by term.
"""
import numpy as np

EPS = np.finfo(np.float32).eps


def confuse_matrix(num_classes, label_gts, label_preds):
    """get_confuse_matrix / __fast_hist — misc/metric_tool.py:141-158"""
    cm = np.zeros((num_classes, num_classes))
    for lt, lp in zip(label_gts, label_preds):
        lt, lp = lt.flatten(), lp.flatten()
        mask = (lt >= 0) & (lt < num_classes)
        cm += np.bincount(num_classes * lt[mask].astype(int) + lp[mask], minlength=num_classes ** 2).reshape(num_classes, num_classes)
    return cm


def cm2score(hist):
    """misc/metric_tool.py:99-138"""
    n_class = hist.shape[0]
    tp = np.diag(hist)
    sum_a1 = hist.sum(axis=1)
    sum_a0 = hist.sum(axis=0)
    acc = tp.sum() / (hist.sum() + EPS)
    recall = tp / (sum_a1 + EPS)
    precision = tp / (sum_a0 + EPS)
    F1 = 2 * recall * precision / (recall + precision + EPS)
    iu = tp / (sum_a1 + hist.sum(axis=0) - tp + EPS)
    d = {'acc': acc, 'miou': np.nanmean(iu), 'mf1': np.nanmean(F1)}
    d.update(dict(zip(['iou_' + str(i) for i in range(n_class)], iu)))
    d.update(dict(zip(['F1_' + str(i) for i in range(n_class)], F1)))
    d.update(dict(zip(['precision_' + str(i) for i in range(n_class)], precision)))
    d.update(dict(zip(['recall_' + str(i) for i in range(n_class)], recall)))
    return d
