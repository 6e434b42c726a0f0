import os, subprocess, sys
sys.path.insert(0, ".")
if len(sys.argv) > 2 and sys.argv[1] == "--run":
    import numpy as np
    from dfki_quad_b200 import wbc
    import torch
    B = 16384
    b = wbc.random_wbc_batch(B, seed=20261023)
    s = wbc.BatchedWBCSolver(max_batch=B)
    H, g, A, lo, hi = s.abi_arrays(b["qp"])
    ext = torch.cuda.ExternalStream(s.stream())
    with torch.cuda.stream(ext):
        d = [torch.from_numpy(a).cuda() for a in (H, g, A, lo, hi)]
        dx = torch.zeros((B, 30), dtype=torch.float64, device="cuda"); di = torch.zeros((B, 48), dtype=torch.uint8, device="cuda")
    ext.synchronize()
    best = 1e9
    for _ in range(5):
        with torch.cuda.stream(ext):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ext)
            s.solve_device(B, *[t.data_ptr() for t in d], dx.data_ptr(), di.data_ptr(), sync=False)
            e1.record(ext)
        ext.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = np.frombuffer(di.cpu().numpy().tobytes(), dtype=np.dtype([("status", "i4"), ("it", "i4"), ("a", "i4"), ("b", "i4"), ("kkt", "f8", (4,))]))
    print(sys.argv[2], f"{B / best / 1e3 * 1e3:.0f} solves/s  {best:.2f} ms  ok {(st['status'] == 0).mean():.4f} it {st['it'].mean():.2f}")
else:
    # usage: python scripts/dbg_wbc_ab.py build_a.so build_b.so ...   (absolute paths of library builds to compare on this box)
    for lib in sys.argv[1:] or [""]:
        env = dict(os.environ)
        if lib:
            env["B200QP_LIB"] = lib
        subprocess.run([sys.executable, __file__, "--run", os.path.basename(lib) or "default"], env=env)
