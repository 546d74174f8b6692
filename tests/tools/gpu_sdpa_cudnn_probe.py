#!/usr/bin/env python
"""One cuDNN-backend F.scaled_dot_product_attention call at the C5 attention shape between cudaProfilerStart/Stop, so
that `ncu --profile-from-start off --set full` captures exactly the library kernel we are compared against
(launch geometry, registers, pipe utilisation, opcode mix). Context for DESIGN.md section 3.2; not product code."""
import sys

import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

B, H, L = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (64, 20, 1500)
D = H * 64
qkv = torch.randn(B, L, 3 * D, device="cuda", dtype=torch.bfloat16)
q, k, v = (qkv[:, :, i * D:(i + 1) * D].unflatten(-1, (H, 64)).transpose(1, 2) for i in range(3))
with torch.no_grad(), sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
    for _ in range(3):
        torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    torch.nn.functional.scaled_dot_product_attention(q, k, v)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok")
