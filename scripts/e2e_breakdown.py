"""Where the end-to-end time of one host-buffer call goes (config 2): wall clock of each layer, 30 repetitions."""
import ctypes as C
import time

import numpy as np
import torch

from dfki_quad_b200 import BatchedMPC, _lib, sequencer

N, B = 10, 4096
cfg = sequencer.GO2_MPC
batch = sequencer.random_batch(B, N, seed=2)
mpc = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "CONDENSED_ADMM", horizon=N,
                 dt=cfg["dt"], max_batch=B)
mpc.set_model(**sequencer.go2_seq_model())
ext = torch.cuda.ExternalStream(mpc.stream())
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
rec_h, ct_h, seq_h = pin(batch["rec"]), pin(batch["contact"]), pin(batch["seq_in"])
f_h = torch.zeros((B, N, 12), dtype=torch.float64).pin_memory()
x_h = torch.zeros((B, N + 1, 13), dtype=torch.float64).pin_memory()
with torch.cuda.stream(ext):
    d_in, d_ct, d_seq = rec_h.cuda(), ct_h.cuda(), seq_h.cuda()
    d_f = torch.zeros((B, N, 12), dtype=torch.float64, device="cuda")
    d_x = torch.zeros((B, N + 1, 13), dtype=torch.float64, device="cuda")
    d_info = torch.zeros((B, 48), dtype=torch.uint8, device="cuda")
ext.synchronize()


def wall(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def ev(fn, reps=30):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        with torch.cuda.stream(ext):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            fn()
            e1.record(ext)
        ext.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.mean(ts))


dev = lambda: mpc.solve_device(B, d_in.data_ptr(), d_ct.data_ptr(), d_f.data_ptr(), d_x.data_ptr(), d_info.data_ptr(), sync=True)
dev_nosync = lambda: mpc.solve_device(B, d_in.data_ptr(), d_ct.data_ptr(), d_f.data_ptr(), d_x.data_ptr(), d_info.data_ptr(), sync=False)
seqdev = lambda: mpc.sequence_and_solve_device(B, d_seq.data_ptr(), None, d_f.data_ptr(), d_x.data_ptr(), d_info.data_ptr(), sync=False)
lib = _lib.load()
info = (_lib.SolverInfo * B)()
raw_rec = lambda: lib.b200qp_mpc_solve_batch(mpc._h, B, rec_h.data_ptr(), ct_h.data_ptr(), f_h.data_ptr(), x_h.data_ptr(), C.addressof(info))
raw_seq = lambda: lib.b200qp_mpc_sequence_solve_batch(mpc._h, B, seq_h.data_ptr(), None, f_h.data_ptr(), x_h.data_ptr(), C.addressof(info))
py_rec = lambda: mpc.solve_records(rec_h.numpy(), ct_h.numpy(), f_h.numpy(), x_h.numpy())
py_seq = lambda: mpc.sequence_and_solve(seq_h.numpy(), None, f_h.numpy(), x_h.numpy())
h2d = lambda: d_in.copy_(rec_h, non_blocking=True)
print("events, device-resident solve          %.3f ms" % ev(dev_nosync))
print("events, device-resident sequencer+solve %.3f ms" % ev(seqdev))
print("wall,   device-resident solve (sync)   %.3f ms" % wall(dev))
with torch.cuda.stream(ext):
    print("events, H2D of the records alone       %.3f ms" % ev(h2d))
print("wall,   C ABI host call, records       %.3f ms" % wall(raw_rec))
print("wall,   C ABI host call, sequencer in  %.3f ms" % wall(raw_seq))
print("wall,   python solve_records           %.3f ms" % wall(py_rec))
print("wall,   python sequence_and_solve      %.3f ms" % wall(py_seq))
