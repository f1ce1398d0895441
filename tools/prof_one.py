#!/usr/bin/env python3
"""Run one micro-benchmark op a few times (for ncu captures): prof_one.py <op> <flags> <dims...>"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlv3p_b200 import ffi  # noqa: E402

op, flags = int(sys.argv[1]), int(sys.argv[2])
dims = [int(a) for a in sys.argv[3:]]
print(ffi.op_time(op, dims, 3, flags))
