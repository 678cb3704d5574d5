"""Times the decimator / correlator with parts switched off (ltb_debug_set_flag) to separate
fill, FMA body and epilogue costs.  Profiling aid; run on a GPU box.  Needs the debug build of the
library (the release one has no switch): it is selected here through LTB200_LIB before the import."""
import os, sys, time
os.environ.setdefault("LTB200_LIB", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                 "gr-ltetrigger_b200", "lib", "libltetrigger_b200_debug.so"))
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))
import torch
import ltetrigger_b200 as lt
import ctypes
DBG = lt.lib()
DBG.ltb_debug_set_flag.argtypes = [ctypes.c_int, ctypes.c_int]

S, D = int(os.environ.get("S", 256)), int(os.environ.get("D", 16))
n = 192000 * D
x = torch.randn((S, n, 2), device="cuda", dtype=torch.float32)
stream = torch.cuda.current_stream()
for variant in (0,):
    name = "decimator"
    for val, what in ((0, "normal"), (1, "no fill"), (2, "no fma body"), (3, "neither")):
        DBG.ltb_debug_set_flag(0, val)
        trig = lt.Trigger(n_streams=S, decim=D, max_chunk=n, record_all=False, cuda_stream=stream.cuda_stream)
        ts = []
        for i in range(4):
            trig.process_device_ptr(x.data_ptr(), n * 8, n)
            ts.append(trig.last_kernel_times())
        t = np.array(ts[1:]).mean(axis=0)
        print("%-10s %-12s frontend %.3f ms  corr %.3f ms  track %.3f ms" % (name, what, t[0], t[1], t[2]), flush=True)
        trig.close()
    DBG.ltb_debug_set_flag(0, 0)
