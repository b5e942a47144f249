"""The plain-C host of the C ABI (examples/dahitra_infer.c): a consumer with no Python and no PyTorch in its process.

CPU: it compiles as C99 against include/dahitra_b200.h, links against the in-tree library, reads the checkpoint file format and
runs dahitra_prepare_weights to the same bytes as the ctypes route of the engine.  GPU: the whole call sequence (prepare ->
upload -> dahitra_workspace_bytes -> dahitra_forward -> copy back) against the fp64 oracle and against the module's own path
(reference models/evaluator.py:156-180 around models/networks.py:1321-1357)."""
import os
import re
import subprocess

import pytest
import torch

from dahitra_b200 import _lib, checkpoints as CK, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FNV0, FNVP, M64 = 14695981039346656037, 1099511628211, (1 << 64) - 1


def fnv1a(raw: bytes) -> int:
    h = FNV0
    for b in raw:
        h = ((h ^ b) * FNVP) & M64
    return h


@pytest.fixture(scope="module")
def exe():
    if _lib.needs_build():
        _lib.build()
    return _lib.build_example()


def run(exe, *args):
    r = subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=600)
    return r.returncode, r.stdout, r.stderr


def test_state_dict_file_round_trip(tmp_path, levir_template):
    """DHSD0001 writer: every floating-point tensor, its name, shape and bytes; integer buffers (num_batches_tracked) skipped"""
    import struct
    sd = synth.synth_state_dict(levir_template, seed=4, style="default")
    sd["as_fp64"] = torch.arange(6, dtype=torch.float64).view(2, 3)
    path = str(tmp_path / "w.bin")
    n = CK.export_state_dict_bin({"model_G_state_dict": {"module." + k: v for k, v in sd.items()}}, path)
    want = {k: v for k, v in sd.items() if v.dtype.is_floating_point}
    assert n == len(want)
    raw = open(path, "rb").read()
    assert raw[:8] == b"DHSD0001" and struct.unpack("<i", raw[8:12])[0] == n
    pos, seen = 12, 0
    while pos < len(raw):
        ln, = struct.unpack("<i", raw[pos:pos + 4])
        name = raw[pos + 4:pos + 4 + ln].decode()
        pos += 4 + ln
        dtype, ndim, s0, s1, s2, s3, nbytes = struct.unpack("<ii4qq", raw[pos:pos + 48])
        pos += 48
        t = want[name]
        assert ndim == t.dim() and [s0, s1, s2, s3][:ndim] == list(t.shape) and dtype == int(t.dtype == torch.float64)
        assert raw[pos:pos + nbytes] == t.contiguous().numpy().tobytes(), name
        pos += nbytes
        seen += 1
    assert seen == n and pos == len(raw)


@pytest.mark.parametrize("variant", ["levir", "xbd"])
def test_c_host_prepares_the_same_bytes_as_the_engine(exe, tmp_path, variant):
    """`dahitra_infer --prepare-only` (C: file -> dh_tensor list -> dahitra_prepare_weights) against engine.prepare_weights_c
    (ctypes over live tensors): same number of floats, same bytes, same slot offsets.  No CUDA call is made."""
    import numpy as np
    from dahitra_b200.engine import prepare_weights_c
    if variant == "levir":
        from dahitra_b200.networks import BASE_Transformer_UNet
        net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
        vid, nc = 0, 2
    else:
        from dahitra_b200.xbd import BASE_Transformer_UNet as X
        net = X(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned", with_decoder_pos="learned", enc_depth=1, dec_depth=8)
        vid, nc = 1, 5
    sd = synth.synth_state_dict(net.state_dict(), seed=22, style="default")
    path = str(tmp_path / "w.bin")
    CK.export_state_dict_bin(sd, path)
    rc, out, err = run(exe, "--weights", path, "--prepare-only", "--variant", vid, "--nc", nc)
    assert rc == 0, err
    got = dict(kv.split("=") for kv in out.split())
    flat, offs = prepare_weights_c(sd, vid, nc)
    assert int(got["floats"]) == flat.numel() and int(got["slots"]) == len(offs)
    assert int(got["present"]) == sum(o >= 0 for o in offs)
    assert int(got["data_fnv1a"], 16) == fnv1a(flat.numpy().tobytes())
    assert int(got["offsets_fnv1a"], 16) == fnv1a(np.asarray(list(offs), dtype="<i8").tobytes())


def test_c_host_reports_errors(exe, tmp_path, levir_template):
    """a missing checkpoint key, a wrong file and a missing argument end with a message and a non-zero exit code, never a crash"""
    sd = synth.synth_state_dict(levir_template, seed=4, style="default")
    sd.pop("resnet.layer1.0.conv1.weight")
    path = str(tmp_path / "w.bin")
    CK.export_state_dict_bin(sd, path)
    rc, out, err = run(exe, "--weights", path, "--prepare-only")
    assert rc == 1 and "dahitra_prepare_weights" in err and "weight" in err
    bad = str(tmp_path / "bad.bin")
    open(bad, "wb").write(b"not a checkpoint")
    rc, out, err = run(exe, "--weights", bad, "--prepare-only")
    assert rc == 1 and "DHSD0001" in err
    rc, out, err = run(exe, "--prepare-only")
    assert rc == 1 and "--weights" in err


def test_example_is_plain_c_and_uses_only_declared_entry_points():
    src = open(os.path.join(ROOT, "examples", "dahitra_infer.c")).read()
    hdr = open(os.path.join(ROOT, "include", "dahitra_b200.h")).read()
    used = set(re.findall(r"\b(dahitra_[a-z0-9_]+)\s*\(", src)) - {"dahitra_infer"}
    declared = set(re.findall(r"\b(dahitra_[a-z0-9_]+)\s*\(", hdr))
    assert used and used <= declared, used - declared
    assert "torch" not in src.replace("PyTorch", "").replace("no Python", "") and "Python.h" not in src


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["levir", "xbd"])
def test_c_host_forward_vs_oracle(exe, tmp_path, variant, levir_template):
    """the C host end to end on the GPU.  LEVIR (2 pairs of 256x256): logits against the fp64 oracle with the tolerance of
    tests/test_gpu_forward.py (1e-4 + 1e-3 |ref|, absolute term widened to 2e-4 max|ref| on these default-scale weights);
    both variants (xBD: one 1024x1024 pair, 5 classes, x1 / x2 = the halves of one (B,6,H,W) tensor): the uint8 class map is the
    argmax of the logits, and logits are bit-identical to what the nn.Module returns for the same weights and images (whose
    parity with the oracle tests/test_gpu_forward.py holds)."""
    from oracle import dahitra_oracle as O                     # checker only
    if variant == "levir":
        from dahitra_b200.networks import BASE_Transformer_UNet
        sd = synth.synth_state_dict(levir_template, seed=3, style="default")
        net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
        x1, x2 = synth.synth_pair(2, 256, 256, seed=5, kind="uniform")
        args, vid, nc = (x1, x2), 0, 2
    else:
        from dahitra_b200.xbd import BASE_Transformer_UNet as X
        net = X(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned", with_decoder_pos="learned", enc_depth=1, dec_depth=8)
        sd = synth.synth_state_dict(net.state_dict(), seed=6, style="default")
        gen = torch.Generator().manual_seed(8)
        args, vid, nc = (torch.randint(0, 256, (1, 6, 1024, 1024), generator=gen).float() / 127 - 1,), 1, 5
    wpath, ipath, opath = (str(tmp_path / n) for n in ("w.bin", "x.bin", "y.bin"))
    CK.export_state_dict_bin(sd, wpath)
    CK.write_pairs_bin(ipath, *args)
    rc, out, err = run(exe, "--weights", wpath, "--input", ipath, "--output", opath, "--variant", vid, "--nc", nc, "--repeat", 5)
    print(out.strip())
    assert rc == 0, err
    assert "forward ok" in out and "pairs_per_s=" in out
    logits, cmap = CK.read_result_bin(opath)
    assert logits.shape == (args[0].shape[0], nc) + tuple(args[0].shape[2:]) and torch.isfinite(logits).all()
    assert torch.equal(cmap.long(), logits.argmax(1))
    if variant == "levir":
        ref = O.forward_levir(sd, x1, x2, dtype=torch.float64)
        d = (logits.double() - ref).abs()
        print(f"[C host] max|d| vs fp64 oracle {float(d.max()):.3e} (max|ref| {float(ref.abs().max()):.2f})")
        assert bool((d <= max(1e-4, 2e-4 * float(ref.abs().max())) + 1e-3 * ref.abs()).all())
        assert float((logits.argmax(1) == ref.argmax(1)).float().mean()) >= 0.999
    net.load_state_dict(sd)
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(*[a.cuda() for a in args]).cpu()
    assert torch.equal(y, logits)
