"""Small end-to-end run for compute-sanitizer: voxelizer + RVT-T-like backbone (narrow fused kernels) + a
wide-dim model (ln_rows / TMA GEMMs / attention_core / cast_xh paths), eager launches, 2 steps each."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rvt_b200  # noqa: E402
from oracle import backbone_oracle as bo  # noqa: E402
from oracle import voxel_oracle as vo  # noqa: E402
from tests.test_host_cpu import make_cfg  # noqa: E402

dev = torch.device('cuda:0')
x, y, p, t = vo.synth_events(3, 20001, 24, 32, hot_fraction=0.2, hot_pixels=2)
sh = rvt_b200.StackedHistogram(10, 24, 32, 10)
sh.construct(*(torch.from_numpy(a).to(dev) for a in (x, y, p, t)))
for embed, dh, hw, part in ((32, 32, (64, 96), (2, 3)), (64, 32, (64, 96), (2, 3))):   # stage dims up to 256 / 512
    spec = bo.BackboneSpec(embed_dim=embed, dim_head=dh, partition_size=part)
    m = rvt_b200.RNNDetector(make_cfg(spec))
    m.load_state_dict(bo.synth_params(spec, 1), strict=True)
    m = m.to(dev).eval()
    st = None
    with torch.no_grad():
        for step in range(2):
            xin = bo.synth_events_tensor(step, 1, 20, *hw).to(dev)
            out, st = m(xin, st)
torch.cuda.synchronize()
print('sanitize run ok')
