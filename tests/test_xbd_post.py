"""xBD post-processing (dahitra_b200.xbd_post) against a numpy restatement of xBD_code/predict_test_cls.py:69-91 and
xBD_code/train.py:266-273."""
import numpy as np
import pytest
import torch


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


def _np_tta(fn, img_hwc):
    """predict_test_cls.py:69-91 for one image: 4 flips in, sigmoid, flips undone, mean"""
    inp = np.asarray([img_hwc, img_hwc[::-1, ...], img_hwc[:, ::-1, ...], img_hwc[::-1, ::-1, ...]], dtype="float")
    msk = _sigmoid(fn(inp.transpose((0, 3, 1, 2))))
    pred = [msk[0, ...], msk[1, :, ::-1, :], msk[2, :, :, ::-1], msk[3, :, ::-1, ::-1]]
    return np.asarray(pred).mean(axis=0)


@pytest.mark.gpu
def test_flip4_tta_and_damage_map_match_the_reference_recipe():
    from dahitra_b200.xbd_post import flip4_tta, damage_map
    rng = np.random.RandomState(0)
    W = torch.from_numpy(rng.randn(5, 6, 3, 3).astype(np.float32)).cuda()

    def net(x):                                   # a position-dependent stand-in network (so flips matter)
        y = torch.nn.functional.conv2d(x, W, padding=1)
        ramp = torch.linspace(-1, 1, x.shape[-1], device=x.device)[None, None, None, :]
        return y + ramp * torch.linspace(0, 1, x.shape[-2], device=x.device)[None, None, :, None]

    imgs = rng.rand(2, 32, 48, 6).astype(np.float32) * 2 - 1
    x = torch.from_numpy(imgs.transpose(0, 3, 1, 2)).cuda()
    got = flip4_tta(net, x).cpu().numpy()
    for b in range(2):
        ref = _np_tta(lambda a: net(torch.from_numpy(a.astype(np.float32)).cuda()).cpu().numpy().astype(np.float64), imgs[b])
        assert np.abs(got[b] - ref).max() < 1e-5
    out = torch.from_numpy(rng.randn(2, 5, 16, 16).astype(np.float32) * 3).cuda()
    msk = _sigmoid(out.cpu().numpy().astype(np.float64))
    ref = msk[:, 1:].argmax(axis=1) * (msk[:, 0] > 0.3)                      # train.py:266-273
    assert np.array_equal(damage_map(out).cpu().numpy(), ref)
