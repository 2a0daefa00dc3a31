"""Validation-path ops — the device-side counterpart of evaluation/metrics.py:runningScore.

`RunningScore` has the reference's interface (`update(label_trues, label_preds)`, `get_scores()`, `reset()`,
`confusion_matrix`) but keeps the confusion matrix on the GPU: `update_from_logits(labels, logits)` fuses the arg-max of
`semantics.data.max(1)[1]` (train.py:846) with the histogram, so neither the logits' arg-max nor the labels travel to
the host per batch (the reference does `.cpu().numpy()` on both, train.py:846-848).  The scores are computed from one
19 x 19 int64 read-back at the end.
"""
import ctypes as C

import numpy as np
import torch

from . import _cabi as A


class RunningScore(object):
    def __init__(self, n_classes, device=None):
        if n_classes > 32:
            raise NotImplementedError("RunningScore: up to 32 classes (shared-memory histogram)")
        self.n_classes = n_classes
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._hist = torch.zeros(n_classes, n_classes, device=self.device, dtype=torch.int64)

    # reference attribute (evaluation/metrics.py:10): a float numpy matrix
    @property
    def confusion_matrix(self):
        return self._hist.cpu().numpy().astype(np.float64)

    def _labels(self, t):
        t = torch.as_tensor(t)
        if not t.is_cuda:
            t = t.to(self.device, non_blocking=True)
        A.require_cuda(t)
        return t.long().contiguous()

    def update_from_logits(self, label_trues, logits):
        """hist[gt, argmax_c logits] += 1 for every pixel with 0 <= gt < n_classes; logits B x C x H x W (planar or
        channels-last), labels B x H x W."""
        gt = self._labels(label_trues)
        A.require_cuda(logits)
        x = logits.detach().float()
        if not (x.is_contiguous() or x.is_contiguous(memory_format=torch.channels_last)):
            x = x.contiguous()
        b, c, h, w = x.shape
        if gt.numel() != b * h * w:
            raise ValueError("labels %s do not match logits %s" % (tuple(gt.shape), tuple(x.shape)))
        A.call("segsde_confusion_update", A.ptr(x), None, A.ptr(gt), C.c_int(b), C.c_int(c), C.c_int64(h * w),
               C.c_int64(x.stride(0)), C.c_int64(x.stride(1)), C.c_int64(x.stride(3)), C.c_int(self.n_classes),
               A.ptr(self._hist), A.stream_ptr())

    def update(self, label_trues, label_preds):
        """Reference signature (metrics.py:19-25): integer label / prediction maps (host or device)."""
        gt, pred = self._labels(label_trues), self._labels(label_preds)
        if gt.numel() != pred.numel():
            raise ValueError("labels %s and predictions %s differ in size" % (tuple(gt.shape), tuple(pred.shape)))
        A.call("segsde_confusion_update", None, A.ptr(pred), A.ptr(gt), C.c_int(1), C.c_int(0), C.c_int64(gt.numel()),
               C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int(self.n_classes), A.ptr(self._hist), A.stream_ptr())

    def get_scores(self):
        """Same dictionary as the reference (metrics.py:27-57), from one read-back of the matrix."""
        hist = self.confusion_matrix
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = np.diag(hist).sum() / hist.sum()
            acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
            iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
            mean_iu = np.nanmean(iu)
            freq = hist.sum(axis=1) / hist.sum()
            fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
        return ({"Overall Acc: \t": acc, "Mean Acc : \t": acc_cls, "FreqW Acc : \t": fwavacc, "Mean IoU : \t": mean_iu},
                dict(zip(range(self.n_classes), iu)))

    def reset(self):
        self._hist.zero_()


runningScore = RunningScore      # the reference's class name
