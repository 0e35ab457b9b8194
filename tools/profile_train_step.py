"""Kernel-time breakdown of one training step (torch.profiler / CUPTI, no ncu replay): python tools/profile_train_step.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dagnn_b200 import _lib, data as D, sharding

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
_lib.build_library()
dev = torch.device("cuda:0")
B = bench.build_workload(wl, 1)
m = bench.build_module(wl).to(dev).train()
G = B.to(dev)
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
flat = sharding.FlatGradients(m.parameters())
y = torch.randint(0, D.CODE2_NUM_VOCAB, (D.CODE2_MAX_SEQ_LEN, int(B.num_graphs))).to(dev)


def step():
    flat.zero()
    pred = m(G)
    loss = sum(torch.nn.functional.cross_entropy(pred[k], y[k]) for k in range(len(pred))) / len(pred)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
