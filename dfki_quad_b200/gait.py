"""Gait definitions and the batched contact schedule (kernel K4).

Mirrors `Gait` / `GaitDatabase` (ws/src/controllers/src/mit_controller/gait.cpp:14-42,67-112,732-767) and the phase
update of `SimpleGaitSequencer::UpdateState` (simple_gait_sequencer.cpp:62-80).  The schedule itself is computed on the
GPU by `b200qp_gait_schedule` (bit-exact with the strict IEEE evaluation of the reference expression); this module
only holds the gait table and marshals arrays.
"""
from __future__ import annotations

import numpy as np

from . import _lib

# GaitDatabase::getGait (gait.cpp:732-767): name -> (period, duty factor, phase offsets fl fr bl br)
GAIT_DATABASE = {
    "STAND": (0.5, 1.0, (0.0, 0.0, 0.0, 0.0)),
    "STATIC_WALK": (1.25, 0.8, (0.0, 0.5, 0.75, 0.25)),
    "WALKING_TROT": (0.5, 0.6, (0.0, 0.5, 0.5, 0.0)),
    "TROT": (0.5, 0.5, (0.0, 0.5, 0.5, 0.0)),
    "FLYING_TROT": (0.4, 0.4, (0.0, 0.5, 0.5, 0.0)),
    "PACE": (0.35, 0.5, (0.0, 0.5, 0.0, 0.5)),
    "BOUND": (0.4, 0.4, (0.0, 0.0, 0.5, 0.5)),
    "ROTARY_GALLOP": (0.4, 0.2, (0.0, 0.8571, 0.3571, 0.5)),
    "TRAVERSE_GALLOP": (0.5, 0.2, (0.0, 0.8571, 0.3571, 0.5)),
    "PRONK": (0.5, 0.5, (0.0, 0.0, 0.0, 0.0)),
}
GAIT_NAMES = list(GAIT_DATABASE.keys())


def gait_arrays(gait_ids):
    """gait_ids: int array [n] indexing GAIT_NAMES -> (period [n], duty [n,4], offset [n,4])."""
    per = np.array([GAIT_DATABASE[g][0] for g in GAIT_NAMES])
    duty = np.array([[GAIT_DATABASE[g][1]] * 4 for g in GAIT_NAMES])
    off = np.array([GAIT_DATABASE[g][2] for g in GAIT_NAMES])
    gait_ids = np.asarray(gait_ids)
    return per[gait_ids].copy(), duty[gait_ids].copy(), off[gait_ids].copy()


def contact_schedule(period, duty, offset, phase, steps, dt, measured_contact=None, fmin=0.0, fmax=0.0,
                     want_bounds=False, device=0):
    """Gait::update_sequence for a batch -> (contact uint8 [n,steps,4], swing_time [n,steps,4][, lbu, ubu [n,steps,12]])."""
    lib = _lib.load()
    period = np.ascontiguousarray(period, dtype=np.float64)
    n = period.shape[0]
    duty = np.ascontiguousarray(np.broadcast_to(duty, (n, 4)), dtype=np.float64)
    offset = np.ascontiguousarray(np.broadcast_to(offset, (n, 4)), dtype=np.float64)
    phase = np.ascontiguousarray(phase, dtype=np.float64)
    meas = None if measured_contact is None else np.ascontiguousarray(measured_contact, dtype=np.uint8)
    contact = np.zeros((n, steps, 4), dtype=np.uint8)
    swing = np.zeros((n, steps, 4))
    lbu = np.zeros((n, steps, 12)) if want_bounds else None
    ubu = np.zeros((n, steps, 12)) if want_bounds else None
    rc = lib.b200qp_gait_schedule(device, n, steps, float(dt), period.ctypes.data, duty.ctypes.data, offset.ctypes.data,
                                  phase.ctypes.data, None if meas is None else meas.ctypes.data, float(fmin), float(fmax),
                                  contact.ctypes.data, swing.ctypes.data,
                                  None if lbu is None else lbu.ctypes.data, None if ubu is None else ubu.ctypes.data)
    _lib.check(rc, "b200qp_gait_schedule")
    return (contact, swing, lbu, ubu) if want_bounds else (contact, swing)


def sequencer_phase(time_in_phase, period, duty=None):
    """SimpleGaitSequencer::UpdateState phase (simple_gait_sequencer.cpp:62-80): fmod((t - t0) / T, 1) with C fmod; a pure
    standing gait (every duty factor 0 or 1, :66-69) keeps phase 0.  `duty` [n,4] enables the standing branch."""
    ph = np.fmod(np.asarray(time_in_phase, dtype=np.float64) / np.asarray(period, dtype=np.float64), 1.0)
    if duty is not None:
        d = np.asarray(duty, dtype=np.float64)
        moving = np.any((d != 0.0) & (d != 1.0), axis=-1)
        ph = np.where(moving, ph, 0.0)
    return ph
