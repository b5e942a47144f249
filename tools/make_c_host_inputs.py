"""Write the files the plain-C host runs of tools/gpu_runs/gpu_r02a{n,o,p,q}.sh read (git-ignored, under examples/bin/):
san_w.bin (LEVIR variant, synthetic default-scale weights, seed 3), san_x.bin (one 256x256 pair), san_w_xbd.bin (xBD variant,
5 classes, seed 6).  Run on the CPU side before the gpurun call; remove the files afterwards (110 MB travel with every snapshot)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dahitra_b200 import checkpoints as CK, synth  # noqa: E402
from dahitra_b200.networks import BASE_Transformer_UNet  # noqa: E402
from dahitra_b200.xbd import BASE_Transformer_UNet as X  # noqa: E402

out = os.path.join(ROOT, "examples", "bin")
os.makedirs(out, exist_ok=True)
torch.manual_seed(123)
net = BASE_Transformer_UNet(3, 2, 'learned', resnet_stages_num=4, with_decoder_pos='learned', enc_depth=1, dec_depth=8)
CK.export_state_dict_bin(synth.synth_state_dict(net.state_dict(), seed=3, style="default"), os.path.join(out, "san_w.bin"))
CK.write_pairs_bin(os.path.join(out, "san_x.bin"), *synth.synth_pair(1, 256, 256, seed=5, kind="uniform"))
net = X(input_nc=3, output_nc=5, token_len=4, resnet_stages_num=4, with_pos="learned", with_decoder_pos="learned", enc_depth=1, dec_depth=8)
CK.export_state_dict_bin(synth.synth_state_dict(net.state_dict(), seed=6, style="default"), os.path.join(out, "san_w_xbd.bin"))
print("wrote", sorted(f for f in os.listdir(out) if f.startswith("san_")))
