#!/usr/bin/env python
"""Print the halo-kernel planner's choice for the stride-1 layer shapes of the BASELINE configs (CPU only)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_deepsort_b200 import _lib
L = ctypes.CDLL(_lib.LIB_PATH)
f = L.ydst_conv_tiling
f.restype = ctypes.c_int
f.argtypes = [ctypes.c_int] * 6 + [ctypes.POINTER(ctypes.c_int)] * 4 + [ctypes.POINTER(ctypes.c_double)]
def q(N, H, W, cin, cout, k):
    bn, ks, occ, ctas = [ctypes.c_int() for _ in range(4)]; us = ctypes.c_double()
    assert f(N, H, W, cin, cout, k, bn, ks, occ, ctas, us) == 0
    gf = 2 * N * H * W * cin * cout * k * k / 1e9
    print(f"N{N:<3} {H:>3}x{W:<3} {cin:>4}->{cout:<4} k{k}: bn {bn.value:<3} ks {ks.value:<2} occ {occ.value} ctas {ctas.value:<4} model {us.value:5.1f}us ({gf / us.value / 1e3 * 1e3:6.0f} TF/s)")
    return us.value
SH = [(1,152,152,64,128,3,2),(1,152,152,128,64,1,2),(1,76,76,128,256,3,11),(1,76,76,256,128,1,10),(1,38,38,256,512,3,11),(1,38,38,512,256,1,11),
      (1,19,19,512,1024,3,7),(1,19,19,1024,512,1,7),(1,19,19,1024,255,1,1),(1,38,38,768,256,1,1),(1,76,76,384,128,1,1),(1,76,76,256,255,1,1),
      (50,64,32,64,64,3,4),(50,32,16,128,128,3,3),(50,16,8,256,256,3,3),(50,8,4,512,512,3,3)]
tot = 0
for a in SH:
    tot += q(*a[:6]) * a[6]
print("sum over yolov3-608 + ReID(50) stride-1 tensor-core convs: %.0f us" % tot)
