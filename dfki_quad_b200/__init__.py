"""dfki_quad_b200 - B200-native batched MPC / WBC QP engine behind the dfki-quad `MPCInterface` plugin boundary.

Everything numerical runs in libb200qp.so (hand-written sm_100a CUDA, C ABI in include/b200qp.h).  Importing this
package does not need a GPU; creating a solver does, and fails loudly otherwise (no CPU fallback)."""
from . import _lib  # noqa: F401
from .gait import GAIT_DATABASE, GAIT_NAMES, contact_schedule, gait_arrays, sequencer_phase  # noqa: F401
from .mpc import BatchedMPC, MPCPrediction, ShardedMPC, SolverInformation, in_doubles, pack_records, pack_seq2  # noqa: F401

__all__ = ["BatchedMPC", "ShardedMPC", "MPCPrediction", "SolverInformation", "contact_schedule", "gait_arrays", "GAIT_DATABASE",
           "GAIT_NAMES", "in_doubles", "pack_records", "pack_seq2", "sequencer_phase"]
