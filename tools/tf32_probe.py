"""Does tcgen05.mma kind::tf32 truncate or round the 13 low mantissa bits of its fp32 operands?
x = 1 + 2^-11 + 2^-12 times 1: truncation gives 1.0, round-to-nearest 1 + 2^-10.  (Decides whether the 3xTF32
splitter has to write the rounded high part back or may leave the operand tile untouched.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops

for val, name in ((1 + 2 ** -11 + 2 ** -12, "1+2^-11+2^-12"), (1 + 2 ** -11, "1+2^-11 (tie)"), (1 + 2 ** -12, "1+2^-12"),
                  (-(1 + 2 ** -11 + 2 ** -12), "-(1+2^-11+2^-12)")):
    x = torch.zeros(128, 32, device="cuda")
    x[:, 0] = val
    w = torch.zeros(32, 32, device="cuda")
    w[0, 0] = 1.0
    y = ops.linear(x, w, None, precision="tf32")
    # operand B as well
    x2 = torch.zeros(128, 32, device="cuda")
    x2[:, 0] = 1.0
    w2 = torch.zeros(32, 32, device="cuda")
    w2[0, 0] = val
    y2 = ops.linear(x2, w2, None, precision="tf32")
    print("%-20s A-operand -> %.10f   B-operand -> %.10f   (trunc %.10f, rna %.10f)"
          % (name, y[0, 0].item(), y2[0, 0].item(), torch.tensor(val).item() and float(int(val * 1024) / 1024 if val > 0 else -int(-val * 1024) / 1024),
             round(val * 1024) / 1024))
