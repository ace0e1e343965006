"""Dev tool: host-side profile of the drop-in route (it is CPU-launch bound: ~0.6 ms of kernels per step)."""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import bench  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "hrnet_w48_tl_620_b4"]
dev = torch.device("cuda", 0)
data = bench.synth_inputs(wl, wl["B"], seed=1, device=dev, pin=False)
st = bench.GpuStep(wl, data, dev)
for _ in range(5):
    st.step_dropin()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    st.step_dropin()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host ms/step (enqueue only): %.3f ; incl. drain: %.3f" % ((t1 - t0) / 50 * 1e3, (t2 - t0) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    st.step_dropin()
pr.disable()
torch.cuda.synchronize()
ps = pstats.Stats(pr)
ps.sort_stats("tottime").print_stats(28)
