"""Where the training step (native pixel decoders unless --stock) (config 4, batch 8, one GPU) spends its device time: torch.profiler, top kernels.
   python tools/profile_train_step.py > profiles/r02_train_step_profile.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from dahitra_b200.networks import define_G


class Args:
    net_G = "newUNetTrans"


torch.manual_seed(0)
net = define_G(Args(), gpu_ids=[0]).train()
net.native_training = "--stock" not in sys.argv
opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.01)
g = torch.Generator(device="cuda").manual_seed(100)
x1 = torch.rand(8, 3, 256, 256, device="cuda", generator=g) * 2 - 1
x2 = torch.rand(8, 3, 256, 256, device="cuda", generator=g) * 2 - 1
y = (torch.rand(8, 256, 256, device="cuda", generator=g) < 0.1).long()


net.channels_last_training = "--nchw" not in sys.argv


def step():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(net(x1, x2), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=150, max_name_column_width=90))
