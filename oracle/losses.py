"""Oracle (test infrastructure only): CPU restatement of the reference's loss value —
`MMCTransformer.losses` (models/MMCTransformer.py:159-179) with `sigmoid_focal_loss`
(models/losses.py:5-53, alpha = 0.7, gamma = 2.0, reduction 'none'), masked and summed.
Pinned against the reference by oracle/make_golden.py (tests/golden/losses_cases.npz)."""
import torch
import torch.nn.functional as F


def sigmoid_focal_loss(inputs, targets, alpha: float = 0.7, gamma: float = 2.0):
    """models/losses.py:36-48"""
    inputs, targets = inputs.float(), targets.float()
    p = torch.sigmoid(inputs)
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss


def losses(masks, out_cls_logits, gt_cls_labels):
    """models/MMCTransformer.py:170-179: logits [B,T,1], labels [B,T], masks [B,1,T] -> scalar"""
    cls = sigmoid_focal_loss(out_cls_logits, gt_cls_labels.unsqueeze(-1))
    return (cls * masks.transpose(1, 2).contiguous()).sum()


def losses_grad(masks, out_cls_logits, gt_cls_labels, batch_size: int = 1):
    """d(cls_loss / batch_size) / d(logits) (main.py:326-333 `final_loss.backward()`), closed form in float64:
    dL/dx = alpha_t [ (p - t)(1 - p_t)^g - g ce (1 - p_t)^(g - 1) p (1 - p)(2 t - 1) ], masked."""
    alpha, gamma = 0.7, 2.0
    x = out_cls_logits.double()
    t = gt_cls_labels.unsqueeze(-1).double()
    p = torch.sigmoid(x)
    ce = F.binary_cross_entropy_with_logits(x, t, reduction="none")
    om = 1 - (p * t + (1 - p) * (1 - t))
    d = (p - t) * om ** gamma - gamma * ce * om ** (gamma - 1) * p * (1 - p) * (2 * t - 1)
    d = (alpha * t + (1 - alpha) * (1 - t)) * d
    return (d * masks.transpose(1, 2) / batch_size).float()
