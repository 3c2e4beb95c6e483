"""Mirror of reference model/init_weight.py: N(0, 0.01) weights, zero bias for Linear; N(0, 0.01) for Conv1d."""
import torch.nn as nn


def init_weight(m):
    if isinstance(m, nn.Linear):
        nn.init.normal_(m.weight, mean=0.0, std=0.01)
        nn.init.constant_(m.bias, 0)
    if isinstance(m, nn.Conv1d):
        nn.init.normal_(m.weight, mean=0.0, std=0.01)
