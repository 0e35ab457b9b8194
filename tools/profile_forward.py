"""Kernel-time breakdown of one forward of the hot path (torch.profiler / CUPTI): python tools/profile_forward.py [workload]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from dagnn_b200 import _lib

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
_lib.build_library()
dev = torch.device("cuda:0")
B = bench.build_workload(wl, 1)
m = bench.build_module(wl).to(dev)
G = B.to(dev)
with torch.no_grad():
    for _ in range(5):
        bench.hot_path(m, G, wl)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            bench.hot_path(m, G, wl)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
