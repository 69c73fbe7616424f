"""RoI max pooling on the NHWC bf16 feature map (csrc/roi_pool_nhwc.cu) through the C ABI
against the oracle's restatement of ROIPoolForward / ROIPoolBackward
(oracle/scda_oracle.c, following extensions/_roi_pooling/src/roi_pooling_kernel.cu:39-91,
137-201) on the same bf16-representable values.  Forward: bit-exact values and the same
argmax element (max of bf16 numbers is exact).  Backward: fp32 sums in a different order."""
import numpy as np
import pytest

import _inputs

pytestmark = pytest.mark.gpu


def _bf16_round(a):
    import torch
    return torch.from_numpy(a).bfloat16().float().numpy()


@pytest.mark.parametrize("shape,R,pool,iw,ih,wh", [
    ((1, 512, 32, 64), 512, (7, 7), 1024, 512, (16, 512)),      # the model's operating point
    ((2, 64, 20, 24), 40, (7, 7), 384, 320, (4, 300)),
    ((1, 8, 16, 16), 9, (3, 5), 256, 256, (8, 200)),
    ((1, 256, 38, 50), 64, (6, 6), 800, 600, (16, 400)),
])
def test_roi_pool_nhwc_matches_oracle(cuda_lib, oracle_mod, shape, R, pool, iw, ih, wh):
    import torch
    from scda_b200 import tc
    B, C, H, W = shape
    feat = _bf16_round(np.maximum(_inputs.features(shape, 0), 0))       # post-ReLU map: exact ties at 0
    rois = _inputs.rois_uniform(R, 1, img_w=iw, img_h=ih, wh=wh, batch=B)
    rois[0, 1:] = [-40, -40, 30, 30]
    rois[1, 1:] = [100, 100, 90, 90]                                      # inverted -> 1 x 1 RoI
    rois[2, 1:] = [iw + 200, ih + 200, iw + 300, ih + 300]                # outside: every bin empty
    scale = 1 / 16.
    o_or, a_or = oracle_mod.roi_pool_forward(feat, rois, pool[0], pool[1], scale)
    f_dev = torch.from_numpy(feat).cuda().permute(0, 2, 3, 1).contiguous().bfloat16()
    r_dev = torch.from_numpy(rois).cuda()
    out, arg = tc.roi_pool_nhwc(f_dev, r_dev, pool[0], pool[1], scale)
    out = out.float().cpu().numpy().reshape(R, C, pool[0], pool[1])
    arg = arg.cpu().numpy().view(np.uint16).reshape(R, C, pool[0], pool[1]).astype(np.int64)
    assert np.array_equal(out, o_or)
    # oracle argmax = flat NCHW offset (batch, c, h, w); ours = h * W + w
    empty = a_or < 0
    assert np.array_equal(arg == 0xFFFF, empty)
    assert np.array_equal(arg[~empty], (a_or[~empty] % (H * W)))

    g = _bf16_round(_inputs.features(o_or.shape, 2))
    gi_or = oracle_mod.roi_pool_backward(g, rois, a_or, feat.shape, scale)
    arg_dev = torch.from_numpy(arg.astype(np.uint16).view(np.int16).reshape(R, -1)).cuda()
    gi = tc.roi_pool_nhwc_bwd(torch.from_numpy(g).cuda().bfloat16().reshape(R, -1).contiguous(), arg_dev, r_dev,
                              (B, H, W, C), pool[0], pool[1])
    gi = gi.permute(0, 3, 1, 2).cpu().numpy()
    np.testing.assert_allclose(gi, gi_or, rtol=1e-4, atol=1e-4)
    assert abs(float(gi.sum()) - float(g[~empty].sum())) <= 1e-3 * max(1.0, abs(float(g[~empty].sum())))


def test_roi_pool_nhwc_rejects_bad_arguments(cuda_lib):
    import torch
    feat = torch.zeros(1, 8, 8, 12, device="cuda", dtype=torch.bfloat16)      # C = 12: not a multiple of 8
    rois = torch.zeros(1, 5, device="cuda")
    out = torch.zeros(1, 12 * 49, device="cuda", dtype=torch.bfloat16)
    arg = torch.zeros(1, 12 * 49, device="cuda", dtype=torch.int16)
    s = torch.cuda.current_stream().cuda_stream
    assert cuda_lib.scda_roi_pool_nhwc_bf16_fwd(feat.data_ptr(), 1 / 16., 1, 1, 8, 8, 12, 7, 7, rois.data_ptr(),
                                                out.data_ptr(), arg.data_ptr(), s) == 0
