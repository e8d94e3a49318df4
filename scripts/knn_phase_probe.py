"""Phase times (SM clock cycles) of scan_select_tile_kernel in the sweep: library built with -DNNGP_KNN_TIMING
(NNGPARA_LIB=...libnngpara_timing.so); stamps of the last launch."""
import ctypes, os, sys, runpy
sys.argv = [sys.argv[0], "40", "25"]
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_target.py"), run_name="__main__")
from nearest_neighbors_gparareal_b200 import _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_longlong * 16)()
lib.nngp_knn_stamps(out)
s = list(out)
print("block 0: setup->scan+sum", s[1] - s[0], "sort+store", s[2] - s[1], "ticket", s[3] - s[2])
print("level 2 parts: heads", s[8]-s[4], "head ranks", s[9]-s[8], "active lists", s[10]-s[9], "ranks", s[11]-s[10], "to barrier", s[5]-s[11])
print("last CTA: level 2", s[5] - s[4], "pair setup", s[6] - s[5], "r2 tiles", s[7] - s[6], "(stamps from different SMs are not comparable)")
