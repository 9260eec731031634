"""Host-side flow table: the caller of the hot path ("next" scope).

Mirrors the behaviour of the reference's ``class Flow`` and ``run_ryu`` line handling
(reference ``traffic_classifier.py:29-96`` and ``:149-165``) with a columnar state instead of one Python
object per flow, so that one ``model.predict`` call classifies every flow of a poll (the reference
calls ``predict`` once per flow, ``:103-106``) and so that the same state array can be advanced on the
GPU by ``tcsdn_flow_update`` (csrc/flow.cu).

State row (float64 x 19), identical to include/tcsdn.h ``TCSDN_FLOW_STATE``:
forward  [0..8]  = packets, bytes, delta_packets, delta_bytes, inst_pps, avg_pps, inst_bps, avg_bps, last_time
reverse  [9..17] = same for the reverse direction
[18]             = time_start
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

STATE = 19
FWD, REV, T0 = 0, 9, 18
# column order handed to model.predict (reference traffic_classifier.py:104)
FEATURE_COLUMNS = (FWD + 2, FWD + 3, FWD + 4, FWD + 5, FWD + 6, FWD + 7,
                   REV + 2, REV + 3, REV + 4, REV + 5, REV + 6, REV + 7)
# column order of a training-data row (reference traffic_classifier.py:121-142, header at :217)
TRAINING_COLUMNS = (FWD + 0, FWD + 1, FWD + 2, FWD + 3, FWD + 4, FWD + 5, FWD + 6, FWD + 7,
                    REV + 0, REV + 1, REV + 2, REV + 3, REV + 4, REV + 5, REV + 6, REV + 7)
TRAINING_HEADER = ("Forward Packets\tForward Bytes\tDelta Forward Packets\tDelta Forward Bytes\t"
                   "Forward Instantaneous Packets per Second\tForward Average Packets per second\t"
                   "Forward Instantaneous Bytes per Second\tForward Average Bytes per second\t"
                   "Reverse Packets\tReverse Bytes\tDelta Reverse Packets\tDelta Reverse Bytes\t"
                   "DeltaReverse Instantaneous Packets per Second\tReverse Average Packets per second\t"
                   "Reverse Instantaneous Bytes per Second\tReverse Average Bytes per second\tTraffic Type\n")


def parse_monitor_line(line: bytes) -> Optional[Tuple[int, str, str, str, str, str, int, int]]:
    """One stdout line of simple_monitor_13.py (reference :66) ->
    (time, datapath, in_port, eth_src, eth_dst, out_port, packets, bytes), or None if it is not a data line.
    Field handling follows traffic_classifier.py:151-165 (split on tabs, utf-8, ints for 0, 6, 7)."""
    if not line.startswith(b"data"):
        return None
    f = [x.decode("utf-8", "strict") for x in line.split(b"\t")[1:]]
    if len(f) < 8:
        return None
    return int(f[0]), f[1], f[2], f[3], f[4], f[5], int(f[6]), int(f[7])


def update_direction(block: np.ndarray, time_start: float, packets: float, nbytes: float, curr_time: float) -> None:
    """Flow.updateforward / updatereverse (reference :63-96) on one 9-slot direction block, in place."""
    block[2] = packets - block[0]
    block[0] = packets
    if curr_time != time_start:
        block[5] = packets / float(curr_time - time_start)
    if curr_time != block[8]:
        block[4] = block[2] / float(curr_time - block[8])
    block[3] = nbytes - block[1]
    block[1] = nbytes
    if curr_time != time_start:
        block[7] = nbytes / float(curr_time - time_start)
    if curr_time != block[8]:
        block[6] = block[3] / float(curr_time - block[8])
    block[8] = curr_time


class FlowTable:
    """Insertion-ordered table of bidirectional flows keyed by (datapath, eth_src, eth_dst)."""

    def __init__(self):
        self._index: Dict[Tuple[str, str, str], int] = {}
        self._meta: List[Tuple[str, str, str, str, str]] = []  # datapath, inport, ethsrc, ethdst, outport
        self._state = np.zeros((64, STATE), np.float64)
        self._active = np.zeros((64, 2), bool)

    def __len__(self):
        return len(self._meta)

    @property
    def state(self) -> np.ndarray:
        return self._state[:len(self)]

    def ingest(self, rec) -> None:
        """One parsed monitor record; same branch order as reference :157-165."""
        t, dp, inport, src, dst, outport, packets, nbytes = rec
        i = self._index.get((dp, src, dst))
        if i is not None:
            self._update(i, 0, packets, nbytes, t)
            return
        i = self._index.get((dp, dst, src))
        if i is not None:
            self._update(i, 1, packets, nbytes, t)
            return
        i = len(self._meta)
        if i == self._state.shape[0]:
            self._state = np.concatenate([self._state, np.zeros_like(self._state)])
            self._active = np.concatenate([self._active, np.zeros_like(self._active)])
        self._index[(dp, src, dst)] = i
        self._meta.append((dp, inport, src, dst, outport))
        row = self._state[i]
        row[:] = 0.0
        row[FWD + 0], row[FWD + 1] = packets, nbytes   # Flow.__init__ :38-39
        row[FWD + 8] = row[REV + 8] = row[T0] = t       # :47, :59, :30
        self._active[i] = (True, False)                 # forward ACTIVE, reverse INACTIVE (:46, :58)

    def _update(self, i, direction, packets, nbytes, t):
        row = self._state[i]
        blk = row[REV:REV + 9] if direction else row[FWD:FWD + 9]
        update_direction(blk, row[T0], packets, nbytes, t)
        self._active[i, direction] = not (blk[3] == 0 or blk[2] == 0)  # :75-78 / :93-96

    def features(self, dtype=np.float64) -> np.ndarray:
        """[n_flows, 12] in the order of reference :104."""
        return np.ascontiguousarray(self.state[:, FEATURE_COLUMNS], dtype=dtype)

    def rows(self):
        """(flow id, src, dst, forward status, reverse status) per flow, table order."""
        for i, (dp, _inport, src, dst, _out) in enumerate(self._meta):
            yield i, src, dst, "ACTIVE" if self._active[i, 0] else "INACTIVE", \
                "ACTIVE" if self._active[i, 1] else "INACTIVE"

    def training_lines(self, traffic_type: str):
        """One TSV line per flow, as reference printflows :121-142 writes them."""
        ints = {0, 1, 2, 3, 8, 9, 10, 11}
        for row in self.state:
            vals = []
            for k, col in enumerate(TRAINING_COLUMNS):
                v = row[col]
                vals.append(str(int(v)) if k in ints else repr(float(v)))
            yield "\t".join(vals + [str(traffic_type)]) + "\n"
