"""Device-side input path against the numpy restatement of the reference loaders; the restatement is
checked against torchvision's own to_tensor / normalize when torchvision is importable."""
import numpy as np
import pytest
import torch

from oracle import inputs_oracle as IO


def test_oracle_matches_torchvision_ops():
    TF = pytest.importorskip("torchvision.transforms.functional")
    from PIL import Image
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(64, 48, 3)).astype(np.uint8)
    ref = TF.normalize(TF.to_tensor(Image.fromarray(img)), mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5]).numpy()
    assert np.array_equal(IO.normalize_levir(img), ref)
    assert len(IO.tiles(np.zeros((1024, 1024, 3), np.uint8))) == 16


@pytest.mark.gpu
@pytest.mark.parametrize("kind,shape,tile", [("levir", (3, 256, 256), 0), ("xbd", (2, 128, 64), 0), ("levir", (2, 1024, 1024), 256),
                                             ("levir", (1, 512, 256), 128)])
def test_normalize_u8_bit_exact(kind, shape, tile):
    from dahitra_b200.inputs import normalize_u8
    N, H, W = shape
    rng = np.random.RandomState(1)
    img = rng.randint(0, 256, size=(N, H, W, 3)).astype(np.uint8)
    y = normalize_u8(torch.from_numpy(img).cuda(), kind=kind, tile=tile).cpu().numpy()
    norm = IO.normalize_levir if kind == "levir" else IO.normalize_xbd
    ref = np.stack([norm(t) for n in range(N) for t in (IO.tiles(img[n], tile) if tile else [img[n]])])
    assert y.shape == ref.shape and y.dtype == np.float32
    assert np.array_equal(y, ref)                         # same fp32 operations in the same order: bit-identical


@pytest.mark.gpu
def test_pipeline_with_u8_inputs(levir_template):
    """PairPipeline(inputs="u8_hwc") == normalising on the host like the reference loader and calling the module."""
    from dahitra_b200.networks import BASE_Transformer_UNet
    from dahitra_b200.pipeline import PairPipeline
    from oracle import synth
    net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
    net.load_state_dict(synth.synth_state_dict(levir_template, seed=3, style="default"))
    net = net.cuda().eval()
    rng = np.random.RandomState(3)
    batches = [(torch.from_numpy(rng.randint(0, 256, size=(2, 256, 256, 3)).astype(np.uint8)).pin_memory(),
                torch.from_numpy(rng.randint(0, 256, size=(2, 256, 256, 3)).astype(np.uint8)).pin_memory()) for _ in range(3)]
    got = [p.clone() for p in PairPipeline(net, out="argmax_u8", inputs="u8_hwc", kind="levir").run(batches)]
    assert len(got) == 3
    with torch.no_grad():
        for (a, b), g in zip(batches, got):
            xa = torch.from_numpy(np.stack([IO.normalize_levir(i) for i in a.numpy()])).cuda()
            xb = torch.from_numpy(np.stack([IO.normalize_levir(i) for i in b.numpy()])).cuda()
            ref = net(xa, xb).argmax(1).to(torch.uint8).cpu()
            assert torch.equal(g, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("out", ["argmax_u8", "argmax", "logits"])
def test_pipeline_results_equal_module_calls(levir_template, out):
    """The three-stream pipeline (upload | forward | download) returns, batch for batch, what calling the module on each
    batch returns — seven distinct batches through two slots, so every buffer is reused at least three times."""
    from dahitra_b200.networks import BASE_Transformer_UNet
    from dahitra_b200.pipeline import PairPipeline
    from oracle import synth
    net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
    net.load_state_dict(synth.synth_state_dict(levir_template, seed=3, style="default"))
    net = net.cuda().eval()
    g = torch.Generator().manual_seed(5)
    batches = [(torch.randn(2, 3, 256, 256, generator=g).pin_memory(), torch.randn(2, 3, 256, 256, generator=g).pin_memory())
               for _ in range(7)]
    got = [p.clone() for p in PairPipeline(net, out=out).run(batches)]
    assert len(got) == 7
    with torch.no_grad():
        for (a, b), r in zip(batches, got):
            y = net(a.cuda(), b.cuda())
            ref = y if out == "logits" else y.argmax(1)
            assert torch.equal(r.to(ref.dtype), ref.cpu())
