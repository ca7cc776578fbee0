"""Shared-memory bandwidth micro-benchmark of libtfin (tfin_smem_bandwidth), for ncu cross-checks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bayesianinferencedl_b200 import _cabi
h = _cabi.TfinHandle(0)
print("smem GB/s", h.smem_bandwidth(), "SMs", h.get_int("sm_count"))
