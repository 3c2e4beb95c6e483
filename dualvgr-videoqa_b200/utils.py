"""Mirror of the reference's top-level utils.py: todevice (:3-8), common_loss (:10-18), loss_dependence (:20-31).
train.py does ``from utils import *`` and calls these on the per-layer GAT outputs (train.py:146-154)."""
import torch

from dualvgr_videoqa_b200 import autograd as ag

__all__ = ["todevice", "common_loss", "loss_dependence"]


def todevice(tensor, device):
    if isinstance(tensor, (list, tuple)):
        assert isinstance(tensor[0], torch.Tensor)
        return [todevice(t, device) for t in tensor]
    elif isinstance(tensor, torch.Tensor):
        return tensor.to(device)


def common_loss(emb1, emb2):
    """mean over [B,N,N] of (E1^ E1^T - E2^ E2^T)^2 with E^ = row-normalised, node-centred embeddings."""
    B, N, _ = emb1.shape
    return ag.PairLossFn.apply(emb1, emb2, 0, 1.0 / (B * N * N))


def loss_dependence(emb1, emb2, dim):
    """HSIC summed over the batch: sum_b tr(R K1 R K2), R = I - 11^T/dim (dim = number of nodes)."""
    if dim != emb1.shape[1]:
        raise ValueError("loss_dependence: dim must equal the number of nodes")
    return ag.PairLossFn.apply(emb1, emb2, 1, 1.0)
