"""Pins what `gap_resize_u8_to_nhwc_bf16` / `gap_resize_nearest_i64` must reproduce: the reference's JointResize
(dataset.py:136-153) calls torchvision.transforms.functional.resize on TENSORS, which for BILINEAR is
F.interpolate(mode="bilinear", align_corners=False, antialias=True) and for NEAREST F.interpolate(mode="nearest").
The GPU test (tests/test_gpu_thin_layers.py) compares the kernels with these F.interpolate forms."""
import pytest
import torch
import torch.nn.functional as F

TF = pytest.importorskip("torchvision.transforms.functional")


@pytest.mark.parametrize("ih,iw,oh,ow", [(50, 70, 32, 32), (20, 20, 64, 64), (256, 256, 256, 256)])
def test_torchvision_tensor_resize_is_antialiased_interpolate(ih, iw, oh, ow):
    g = torch.Generator().manual_seed(ih + ow)
    x = torch.rand(3, ih, iw, generator=g)
    a = TF.resize(x, (oh, ow), interpolation=TF.InterpolationMode.BILINEAR)
    b = F.interpolate(x[None], size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)[0]
    assert torch.equal(a, b)
    lab = (torch.rand(ih, iw, generator=g) < 0.3).long()
    c = TF.resize(lab.unsqueeze(0), (oh, ow), interpolation=TF.InterpolationMode.NEAREST).squeeze(0)
    d = F.interpolate(lab[None, None].float(), size=(oh, ow), mode="nearest")[0, 0].long()
    assert torch.equal(c, d)
