"""tcgen05.mma dependency microbenchmark: cycles per MMA for chains on 1, 2 and 4 accumulators (N = 32 ... 256)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gigagan_pytorch_b200 import _lib
dev = torch.device("cuda:0")
out = torch.zeros(2, dtype=torch.int64, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
print("N    nacc  iters  issue cyc/MMA   total cyc/MMA   (tensor work = N/2 cycles per MMA)")
for N in (16, 32, 64, 128, 256):
    for nacc in (1, 2, 4, 8):
        if nacc * N > 512:
            continue
        for iters in (64,):
            for _ in range(2):
                _lib.call("gg_debug_mma_chain", N, nacc, iters, ctypes.c_void_p(out.data_ptr()), st)
            torch.cuda.synchronize()
            a, b = out.tolist()
            print(f"{N:4d} {nacc:5d} {iters:6d} {a / iters:14.1f} {b / iters:15.1f}")
