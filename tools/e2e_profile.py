"""Where does the e2e step of bench.py go?  torch.profiler over a few module-level steps (GPU box).
    python tools/e2e_profile.py [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import workload as W
from categoricalnf_b200 import ops

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
B, S, V = W.LM["B"], W.LM["S"], W.LM["V"]
prm = W.lm_params(seed=0)
W.LMDevicePath(prm, dev).data_init(seed=0)
model, prior = W.build_lm_model(prm, dev)
host_tokens = [W.lm_tokens(B, S, V, seed=j).pin_memory() for j in range(2)]
host_ll = torch.empty(B, dtype=torch.float32).pin_memory()


def step(i):
    with torch.no_grad():
        tok = host_tokens[i % 2].to(dev, non_blocking=True)
        z, ldj = model(tok)
        logp, _ = ops.logistic_logprob(z)
        host_ll.copy_(ldj + logp, non_blocking=True)
        torch.cuda.current_stream().synchronize()


for i in range(3):
    step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    step(i)
e1.record()
torch.cuda.synchronize()
print("e2e step: %.3f ms" % (e0.elapsed_time(e1) / steps))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(steps):
        step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=15, max_name_column_width=60))
