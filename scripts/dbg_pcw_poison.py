"""Read-before-write hunt in the warp Riccati kernel: NaN-fill slices of the per-block scratch (B200QP_POISON_SCRATCH=lo:hi, doubles from
the end of the block's region) and see whether a batch that is solved 100 % otherwise changes."""
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
    ranges = ["0:0", "0:100000", "0:736", "736:940", "940:1228", "1228:1388", "1388:1484", "1484:1592", "1592:1796", "1796:1892", "1892:1988", "1988:2084",
              "2084:2400", "2400:100000"]
    for r in ranges:
        env = dict(os.environ, B200QP_POISON_SCRATCH=r)
        subprocess.run([sys.executable, __file__, r], env=env)
