"""Read-before-write hunt in the warp Riccati kernel: NaN-fill the scratch (B200QP_POISON_SCRATCH=1) or every SM's shared memory (B200QP_POISON_SMEM=1) and see whether a batch that is solved 100 % otherwise changes."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np
    from dfki_quad_b200 import BatchedMPC, sequencer
    cfg = sequencer.GO2_MPC
    out = []
    for N, gaits, seed in ((10, ("TRAVERSE_GALLOP", "BOUND"), 1024), (16, ("TROT", "PACE", "BOUND"), 3), (10, ("PRONK", "FLYING_TROT"), 5)):
        b = sequencer.random_batch(512, N, seed=seed, gaits=gaits)
        m = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], 0.6, 0.0, 400.0, "SPARSE_RICCATI", horizon=N, dt=cfg["dt"], max_batch=512)
        f, x, info = m.solve_records(b["rec"], b["contact"])
        out.append(f"{info.success.mean():.4f}/{np.bincount(info.return_code, minlength=4).tolist()}/{float(np.abs(f[info.success]).sum()):.9e}")
        m.close()
    print(sys.argv[1], " | ".join(out))
else:
    subprocess.run([sys.executable, __file__, "clean"], env=dict(os.environ))
    subprocess.run([sys.executable, __file__, "scratch NaN-filled"], env=dict(os.environ, B200QP_POISON_SCRATCH="1"))
    subprocess.run([sys.executable, __file__, "shared memory NaN-filled"], env=dict(os.environ, B200QP_POISON_SMEM="1"))
