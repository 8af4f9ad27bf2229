"""cnf_linear_fwd (3xTF32, pre-split parameter weights) at the projection shapes of a small GraphCNF shard, for every N tile:
the data behind the automatic N-tile choice of launch_gemm (csrc/linear_tc.cu).
    python tools/bench_linear_bn.py [--reps 30]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from categoricalnf_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--shape", type=int, nargs=3, default=None, help="only this M N K (e.g. for an ncu capture)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
shapes = [(2432, 768, 384), (2432, 192, 384), (2432, 384, 768), (2432, 384, 384), (2432, 1152, 384), (2432, 1536, 384),
          (1944, 192, 192), (1944, 388, 192), (27656, 192, 192), (27656, 388, 192), (27656, 384, 192), (9, 12, 64),
          (2432, 384, 8), (9728, 768, 384), (9728, 192, 384)]
if a.shape:
    shapes = [tuple(a.shape)]
for (M, N, K) in shapes:
    x = torch.randn(M, K, device=dev)
    w = torch.nn.Parameter(torch.randn(N, K, device=dev) / K ** 0.5)
    b = torch.nn.Parameter(torch.randn(N, device=dev))
    row = {"M": M, "N": N, "K": K}
    for bn in (0, 256, 224, 192, 160, 128, 96, 64, 32):
        if bn and (bn > ((N + 31) // 32) * 32 and bn != 256):
            continue
        if bn == 256 and N <= 224:
            continue
        fn = lambda: ops.linear(x, w, b, block_n=bn)
        with torch.no_grad():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(a.reps):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
        row["auto" if bn == 0 else str(bn)] = round(e0.elapsed_time(e1) / a.reps * 1e3, 1)
    print(json.dumps(row), flush=True)
