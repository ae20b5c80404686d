"""Streaming statistical metrics (mirrors reference multinn/metrics/statistical.py:6-47). The reference builds
`tf.metrics.mean / accuracy / precision / recall` accumulators that `collect_metrics` updates batch by batch
(utils/training.py:180-213); here the accumulators are six float64 counters that live on the tensors' device.
Reporting code: torch is used as the array library so that nothing leaves the device until `result()`.

Quirk Q5 kept by default: the reference's f1 score uses the precision in place of the recall (statistical.py:37-38), so
`f1_score` equals the precision whenever it is positive; `true_f1` is the textbook value."""
import torch


class BaseMetrics:
    """loss / log_likelihood / perplexity are means over rows (log_probs = the positive per-row NLL the generators return,
    rnn_nade.py:279-302); accuracy / precision / recall count elements of the binary predictions against the targets."""

    def __init__(self, device='cpu'):
        self.acc = torch.zeros(7, dtype=torch.float64, device=device)   # sum nll, sum exp nll, rows, TP, FP, FN, correct
        self.elements = 0

    def update(self, log_probs, targets=None, predictions=None):
        lp = log_probs.double().reshape(-1)
        self.acc[0] += lp.sum()
        self.acc[1] += lp.exp().sum()
        self.acc[2] += lp.numel()
        if targets is not None and predictions is not None:
            t, p = targets.reshape(-1) > 0.5, predictions.reshape(-1) > 0.5
            if t.numel() != p.numel():
                raise ValueError(f'targets hold {t.numel()} elements, predictions {p.numel()}')
            self.acc[3] += (t & p).sum()
            self.acc[4] += (~t & p).sum()
            self.acc[5] += (t & ~p).sum()
            self.acc[6] += (t == p).sum()
            self.elements += t.numel()

    def allreduce(self):
        """Data parallel: sum the accumulators over the ranks (one collective)."""
        import torch.distributed as dist
        n = torch.tensor([float(self.elements)], dtype=torch.float64, device=self.acc.device)
        buf = torch.cat([self.acc, n])
        dist.all_reduce(buf)
        self.acc, self.elements = buf[:7], int(buf[7])

    def result(self):
        s_nll, s_ppl, rows, tp, fp, fn, correct = (float(a) for a in self.acc)
        rows = max(rows, 1.0)
        out = {'loss': s_nll / rows, 'log_likelihood': s_nll / rows, 'perplexity': s_ppl / rows, 'rows': int(self.acc[2])}
        if self.elements:
            precision = tp / (tp + fp) if tp + fp > 0 else 0.0           # tf.metrics.precision: 0 when nothing predicted
            recall = tp / (tp + fn) if tp + fn > 0 else 0.0
            out.update(accuracy=correct / self.elements, precision=precision, recall=recall,
                       f1_score=precision if precision > 0 else 0.0,                          # quirk Q5
                       true_f1=2 * precision * recall / (precision + recall) if precision + recall > 0 else 0.0)
        return out


def global_reconstruction_metrics(targets, predictions, eps=1e-7):
    """The reference's `metrics['global']` (core/multi_encoder_nn.py:117-152 -> encoders/pass_encoder.py:77-92 ->
    statistical.py:6-47): per track, tf.losses.log_loss(targets, decoded predictions) summed over the dimensions is the
    per-row 'log prob'; loss / log_likelihood are its row means, accuracy / precision / recall count elements; the
    global value of each metric is the mean over the tracks (multinn_core.py:402-405). With Pass encoders the decoded
    predictions are the generator's THRESHOLDED outputs (rnn_multinade.py:124-128), so every mismatch costs
    -log(1e-7) = 16.1 and the number is a scaled Hamming distance, not a likelihood. targets, predictions: [M,N,D]
    (track-major, any array type). Reporting code, float64 on the host."""
    import numpy as np
    to_np = lambda a: a.detach().cpu().numpy() if hasattr(a, 'detach') else np.asarray(a)
    t, p = to_np(targets).astype(np.float64), to_np(predictions).astype(np.float64)
    if t.shape != p.shape or t.ndim != 3:
        raise ValueError(f'targets {t.shape} and predictions {p.shape} must both be [tracks, rows, dims]')
    row = (-t * np.log(p + eps) - (1.0 - t) * np.log(1.0 - p + eps)).sum(2)            # losses_impl.log_loss
    tb, pb = t > 0.5, p > 0.5
    tp, fp, fn = (tb & pb).sum((1, 2)), (~tb & pb).sum((1, 2)), (tb & ~pb).sum((1, 2))
    safe = lambda a, b: np.where(b > 0, a / np.maximum(b, 1), 0.0)
    precision, recall = safe(tp, tp + fp), safe(tp, tp + fn)
    loss = row.mean(1)
    with np.errstate(over='ignore'):
        perplexity = np.exp(row).mean(1)
    per_track = dict(loss=loss, log_likelihood=loss, perplexity=perplexity, accuracy=(tb == pb).mean((1, 2)),
                     precision=precision, recall=recall, f1_score=np.where(precision > 0, precision, 0.0))   # quirk Q5
    out = {k: float(v.mean()) for k, v in per_track.items()}
    out['tracks'] = {k: v.tolist() for k, v in per_track.items()}
    return out
