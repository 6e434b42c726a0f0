"""Multi-GPU tests of the library-level sharded call (need >= 2 visible GPUs; skipped otherwise)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ngpu():
    from dfki_quad_b200 import _lib

    return _lib.load().b200qp_device_count()


def test_sharded_call_equals_single_gpu_call():
    """b200qp_mpc_solve_batch_sharded on all visible GPUs: one host array in, disjoint output ranges, bit-identical to the
    single-GPU call (same kernels, same problems), failure semantics kept per slice."""
    import torch
    from dfki_quad_b200 import BatchedMPC, ShardedMPC, sequencer

    G = _ngpu()
    if G < 2:
        pytest.skip("needs >= 2 GPUs")
    N, n = 10, 3000 * G + 7
    cfg = sequencer.GO2_MPC
    batch = sequencer.random_batch(n, N, seed=99, gaits=("TROT", "PACE", "STAND"))
    rec = batch["rec"].copy()
    rec[n // 2, 5] = np.nan  # one failing problem in the middle slice
    one = BatchedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "AUTO", horizon=N, dt=cfg["dt"], max_batch=n)
    f1, x1, i1 = one.solve_records(rec, batch["contact"])
    sh = ShardedMPC(cfg["alpha"], cfg["state_weights_move"], cfg["mu"], cfg["fmin"], cfg["fmax"], "AUTO", horizon=N, dt=cfg["dt"], max_batch=n,
                    devices=tuple(range(G)))
    for pinned in (False, True):
        if pinned:
            recp, ctp = torch.from_numpy(rec).pin_memory().numpy(), torch.from_numpy(batch["contact"]).pin_memory().numpy()
            f = torch.full((n, N, 12), 123.0, dtype=torch.float64).pin_memory().numpy()
            x = torch.full((n, N + 1, 13), 321.0, dtype=torch.float64).pin_memory().numpy()
        else:
            recp, ctp = rec, batch["contact"]
            f, x = np.full((n, N, 12), 123.0), np.full((n, N + 1, 13), 321.0)
        f, x, info = sh.solve_records(recp, ctp, f, x)
        ok = info["status"] == 0
        assert ok.sum() == n - 1 and not ok[n // 2]
        assert np.all(f[n // 2] == 123.0) and np.all(x[n // 2] == 321.0)
        assert np.array_equal(f[ok], f1[ok]) and np.array_equal(x[ok], x1[ok])
    sh.close()
