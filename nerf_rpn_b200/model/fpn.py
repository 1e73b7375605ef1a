"""FPN neck. Mirrors nerf_rpn/model/fpn.py:59-132 for the configuration the reference uses
(FPN([128,256,512,512], 256, 4), feature_extractor.py:304): per level a 1^3 lateral conv and a 3^3 output conv, created in
the reference's order (lateral_i, fpn_i alternating) so seeds and state_dict keys (lateral_convs.{i}.*, fpn_convs.{i}.*)
match.  forward (fpn.py:134-161: top-down nearest-upsample accumulation, then the 3^3 convs) runs inside the backbone's captured
engine; called on its own it launches the same kernels eagerly (model/_eager.py)."""
from torch import nn


class FPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 extra_convs_on_inputs=True, relu_before_extra_convs=False, upsample_cfg=dict(mode='nearest')):
        super().__init__()
        assert isinstance(in_channels, list)
        if start_level != 0 or end_level != -1 or add_extra_convs or num_outs != len(in_channels) or \
                upsample_cfg.get("mode", "nearest") != "nearest" or "scale_factor" in upsample_cfg:
            raise NotImplementedError("nerf_rpn_b200.FPN implements the reference's usage: all input levels, no extra convs, "
                                      "size-based nearest up-sampling")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_ins, self.num_outs = len(in_channels), num_outs
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.num_ins):
            self.lateral_convs.append(nn.Conv3d(in_channels[i], out_channels, 1))
            self.fpn_convs.append(nn.Conv3d(out_channels, out_channels, 3, padding=1))

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, inputs):
        """fpn.py:134-161 stand-alone: tuple of (N, C_i, w, l, h) fp32 CUDA -> tuple of (N, out_channels, w, l, h)."""
        from ._eager import fpn_forward
        return fpn_forward(self, inputs, getattr(self, "precision", None))
