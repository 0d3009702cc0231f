"""RunningScore / cal_text_score -- drop-in for the pixel metric of the reference's src/text_metrics.py:9-82.

The reference copies the full probability map to the host every training step and runs numpy.bincount
(src/train.py:176-181).  Here the 2x2 confusion matrix is accumulated on the device by one 12-byte-per-pixel kernel
(csrc/db_loss.cu: text_score_hist_kernel); only four integers cross PCIe.  (QuadMetric / the polygon evaluators of
that file need shapely and are out of scope)"""
import numpy as np
import torch

from . import _lib


class RunningScore:
    def __init__(self, n_classes):
        if n_classes != 2:
            raise ValueError("the DB pixel metric is binary (cfg.hps.no_classes = 2)")
        self.n_classes = n_classes
        self.confusion_matrix = np.zeros((n_classes, n_classes))

    def update_from_device(self, hist4):
        self.confusion_matrix += hist4.cpu().numpy().astype(np.float64).reshape(2, 2)

    def update(self, label_trues, label_preds):
        """Host arrays, as in the reference (kept for callers that already hold numpy labels)."""
        for lt, lp in zip(label_trues, label_preds):
            lt, lp = np.asarray(lt).flatten(), np.asarray(lp).flatten()
            m = (lt >= 0) & (lt < self.n_classes)
            self.confusion_matrix += np.bincount(self.n_classes * lt[m].astype(int) + lp[m],
                                                 minlength=self.n_classes ** 2).reshape(self.n_classes, self.n_classes)

    def get_scores(self):
        """src/text_metrics.py:35-58."""
        hist = self.confusion_matrix
        acc = np.diag(hist).sum() / (hist.sum() + 0.0001)
        acc_cls = np.nanmean(np.diag(hist) / (hist.sum(axis=1) + 0.0001))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist) + 0.0001)
        mean_iu = np.nanmean(iu)
        freq = hist.sum(axis=1) / (hist.sum() + 0.0001)
        fwavacc = (freq[freq > 0] * iu[freq > 0]).sum()
        return {'Overall Acc': acc, 'Mean Acc': acc_cls, 'FreqW Acc': fwavacc, 'Mean IoU': mean_iu}, dict(zip(range(self.n_classes), iu))

    def reset(self):
        self.confusion_matrix = np.zeros((self.n_classes, self.n_classes))


def text_score_hist(texts, gt_texts, training_masks, thresh=0.5, out=None):
    """Device 2x2 confusion matrix (uint64[4], [gt*2+pred]); accumulates into ``out`` when given."""
    _lib.require_cuda(texts, gt_texts, training_masks)
    n, h, w = texts.shape
    if texts.stride(2) != 1 or texts.stride(1) != w:
        texts = texts.contiguous()
    gt = gt_texts.float().contiguous()
    mk = training_masks.float().contiguous()
    if out is None:
        out = torch.zeros(4, dtype=torch.int64, device=texts.device)
    with torch.cuda.device(texts.device):
        _lib.check(_lib.lib().dbb_text_score_hist(texts.data_ptr(), texts.stride(0), gt.data_ptr(), mk.data_ptr(), n, h, w,
                                                  float(thresh), out.data_ptr(), _lib.stream_ptr()), "dbb_text_score_hist")
    return out


def cal_text_score(texts, gt_texts, training_masks, running_metric_text, thresh=0.5):
    """src/text_metrics.py:63-82: texts = preds[:, 0] (a channel view is fine, no copy is made)."""
    hist = text_score_hist(texts.detach().float(), gt_texts, training_masks, thresh)
    running_metric_text.update_from_device(hist)
    score_text, _ = running_metric_text.get_scores()
    return score_text
