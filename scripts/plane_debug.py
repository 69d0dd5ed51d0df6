#!/usr/bin/env python
"""Targeted experiments for the plane kernel: which kind of tap shift breaks?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tc_check as t
from elektronn2_b200 import _lib
h = _lib.get_handle(0)
print('E2_PLANE_BO', os.environ.get('E2_PLANE_BO'))
for k in [(1, 1, 2), (1, 1, 3), (1, 1, 9), (3, 3, 3)]:
    sp = (3 + k[0], 15 + k[1], 15 + k[2])   # output (4,16,16): no ragged edges
    t.check_conv(h, 1, 32, sp, 64, k)
