"""Per-CTA timeline of the decimator (profiling aid): how long staging and arithmetic take and how
many CTAs of an SM are in each phase at a time."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ltetrigger_b200", "python"))
import torch
import ltetrigger_b200 as lt

S, D = int(os.environ.get("S", 128)), 16
n = 192000 * D
x = torch.randn((S, n, 2), device="cuda", dtype=torch.float32)
tiles = 375 * S
buf = torch.zeros((tiles, 4), dtype=torch.int64, device="cuda")
lt.lib().ltb_debug_set_trace(buf.data_ptr())
stream = torch.cuda.current_stream()
trig = lt.Trigger(n_streams=S, decim=D, max_chunk=n, record_all=False, cuda_stream=stream.cuda_stream)
trig.process_device_ptr(x.data_ptr(), n * 8, n)
lt.lib().ltb_debug_set_flag(0, 4)
trig.process_device_ptr(x.data_ptr(), n * 8, n)
lt.lib().ltb_debug_set_flag(0, 0)
t = buf.cpu().numpy()
t0 = t[:, 0].min()
start, fill, end, sm = t[:, 0] - t0, t[:, 1] - t0, t[:, 2] - t0, t[:, 3]
print("kernel span %.3f ms, tiles %d" % (end.max() / 1e6, tiles))
print("staging  ns: median %.0f  p10 %.0f  p90 %.0f" % tuple(np.percentile(fill - start, [50, 10, 90])))
print("compute  ns: median %.0f  p10 %.0f  p90 %.0f" % tuple(np.percentile(end - fill, [50, 10, 90])))
# occupancy of phases on one SM over time
for s_id in (0, 37, 100):
    m = sm == s_id
    ev = sorted([(a, 'S') for a in start[m]] + [(b, 'F') for b in fill[m]] + [(c, 'E') for c in end[m]])
    nf = nc = 0; last = 0; hist = {}
    for tt, k in ev:
        hist[(nf, nc)] = hist.get((nf, nc), 0) + (tt - last); last = tt
        if k == 'S': nf += 1
        elif k == 'F': nf -= 1; nc += 1
        else: nc -= 1
    tot = sum(hist.values())
    print("SM %d: tiles %d; time share by (n_staging, n_computing):" % (s_id, m.sum()),
          {k: round(v / tot, 3) for k, v in sorted(hist.items()) if v / tot > 0.01})
