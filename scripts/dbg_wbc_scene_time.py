"""Development aid: device time of the scene path (K6 + K3 + torques) beside K3 alone, on the handle's stream."""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from dfki_quad_b200 import wbc

B = 16384
b = wbc.random_wbc_batch(B, seed=20261023)
s = wbc.BatchedWBCSolver(max_batch=B)
s.scene_configure()
mb = b["mpc_batch"]
scene = wbc.pack_scene(mb["pos"], mb["quat"], mb["vel"], mb["omega"], b["q"], b["qd"], b["contacts"], b["a_body"], b["a_foot"], b["f_ref"])
H, g, A, lo, hi = s.abi_arrays(b["qp"])
ext = torch.cuda.ExternalStream(s.stream())
with torch.cuda.stream(ext):
    d = [torch.from_numpy(a).cuda() for a in (H, g, A, lo, hi)]
    dsc = torch.from_numpy(scene).cuda()
    dx = torch.zeros((B, 30), dtype=torch.float64, device="cuda"); dt = torch.zeros((B, 12), dtype=torch.float64, device="cuda")
    di = torch.zeros((B, 48), dtype=torch.uint8, device="cuda")
ext.synchronize()


def timed(fn):
    best = 1e9
    for _ in range(5):
        with torch.cuda.stream(ext):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext); fn(); e1.record(ext)
        ext.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


t_qp = timed(lambda: s.solve_device(B, *[t.data_ptr() for t in d], dx.data_ptr(), di.data_ptr(), sync=False))
t_sc = timed(lambda: s.scene_solve_device(B, dsc.data_ptr(), dx.data_ptr(), dt.data_ptr(), di.data_ptr(), sync=False))
print(f"K3 alone {t_qp:.3f} ms; K6 + K3 + torques {t_sc:.3f} ms; difference {t_sc - t_qp:.3f} ms")
