#!/usr/bin/env python
"""Device facts the design depends on: L2 size, persisting-L2 limits, SM count, peer access."""
import ctypes as C, torch
rt = C.CDLL("libcudart.so.12") if True else None
def attr(a, dev=0):
    v = C.c_int(0); rt.cudaDeviceGetAttribute(C.byref(v), a, dev); return v.value
p = torch.cuda.get_device_properties(0)
print("device", p.name, "SMs", p.multi_processor_count, "L2 bytes", p.L2_cache_size, "mem GB", p.total_memory / 2**30)
print("max persisting L2", attr(108), "max access policy window", attr(109), "smem/SM", attr(81), "smem/block optin", attr(97))
print("n devices", torch.cuda.device_count())
