import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gi-gs_b200"))
import torch
from gigs import light as GL
dev = torch.device("cuda:0")
base = torch.rand(6, 256, 256, 3, device=dev) * 0.5 + 0.25
fl = GL.PrefilteredLight(base)
gb = torch.zeros_like(base)
for _ in range(3):
    fl.build(); fl.backward(gb, accumulate=False)
torch.cuda.synchronize()
print("ok")
