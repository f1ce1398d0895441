#!/usr/bin/env python3
"""Smallest whole-model run for ncu: 2 forwards of the B=32 512x512 model.  tools/ncu_one.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dlv3p_b200  # noqa: E402
from bench import random_weights  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
net = dlv3p_b200.DeepLabV3PlusXception((512, 512, 3), 21, 16, batch=B, device=0)
net.set_weights(random_weights(net.weight_specs()))
img = torch.randint(0, 256, (B, 512, 512, 3), device='cuda', dtype=torch.uint8)
out = torch.empty((B, 512, 512), device='cuda', dtype=torch.uint8)
for _ in range(2):
    net.model.forward(img.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
