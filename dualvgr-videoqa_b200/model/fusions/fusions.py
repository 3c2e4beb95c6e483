"""MFB fusion block — mirror of the one class of reference model/fusions/fusions.py that the hot path constructs
(MFB, :382-453, built at model/models.py:109 with mm_dim=256, factor=2, ELU in/out, no dropout, no normalisation)."""
import torch.nn as nn

from dualvgr_videoqa_b200 import autograd as ag   # made importable by model/__init__.py


class MFB(nn.Module):
    def __init__(self, input_dims, output_dim, mm_dim=256, factor=2, activ_input='elu', activ_output='elu',
                 normalize=False, dropout_input=0., dropout_pre_norm=0., dropout_output=0.):
        super().__init__()
        if (activ_input, activ_output, normalize) != ('elu', 'elu', False) or factor != 2 or \
                dropout_input or dropout_pre_norm or dropout_output:
            raise NotImplementedError("the sm_100a MFB path implements the configuration DualVGR uses "
                                      "(ELU in/out, factor 2, no dropout, no normalisation)")
        self.input_dims, self.output_dim, self.mm_dim, self.factor = input_dims, output_dim, mm_dim, factor
        self.linear0 = nn.Linear(input_dims[0], mm_dim * factor)
        self.linear1 = nn.Linear(input_dims[1], mm_dim * factor)
        self.linear_out = nn.Linear(mm_dim, output_dim)
        self.n_params = sum(p.numel() for p in self.parameters() if p.requires_grad)
        for lin in (self.linear0, self.linear1, self.linear_out):      # reference model/init_weight.py
            nn.init.normal_(lin.weight, mean=0.0, std=0.01)
            nn.init.constant_(lin.bias, 0)

    def forward(self, x):
        slot = ag.PairSlot()        # the two input gradients land in the halves of ONE buffer (the unit stack's stacked layout)
        x0 = ag.linear(x[0], self.linear0.weight, self.linear0.bias, act="elu", act_grad_folded=True, dx_slot=(slot, 0))
        x1 = ag.linear(x[1], self.linear1.weight, self.linear1.bias, act="elu", act_grad_folded=True, dx_slot=(slot, 1))
        z = ag.MfbPairFn.apply(x0, x1)
        return ag.linear(z, self.linear_out.weight, self.linear_out.bias, act="elu")
