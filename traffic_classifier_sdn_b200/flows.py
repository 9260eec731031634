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


class DeviceFlowTable(FlowTable):
    """The same table with its state rows in HBM: ``Flow.updateforward / updatereverse`` (reference
    ``traffic_classifier.py:63-96``) run in ``tcsdn_flow_update`` (csrc/flow.cu), the twelve features of ``:104`` are
    written by that kernel into a device matrix and ``model.predict_indices`` reads them there -- one report is
    update -> predict with the features never leaving the GPU.  Flow keying and line parsing stay on the host (row N2).

    The reference applies monitor lines one at a time.  Here samples are queued and applied by one kernel call per batch;
    a second sample for a flow that already has one queued flushes the queue first, so every flow still sees its samples
    in order and the state is bit-identical to the host table's (tests/test_parity_gpu.py).  Device memory comes from
    torch (plumbing); the arithmetic is libtcsdn's."""

    def __init__(self, device=None, feature_dtype=np.float64):
        import torch
        from . import _lib
        super().__init__()
        self._torch, self._lib = torch, _lib
        self._dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._fdtype = torch.float32 if np.dtype(feature_dtype) == np.float32 else torch.float64
        self._dstate = torch.zeros((64, STATE), dtype=torch.float64, device=self._dev)
        self._feat = None
        self._pend = {}          # flow index -> (direction, packets, bytes, time)
        self._new_rows = {}      # flow index -> initial state row (Flow.__init__)
        self._updated = {}       # flow index -> bit 0 / 1: the forward / reverse direction has had an update
        self._dirty = True       # features out of date

    # ---- queueing (host, same branch order as FlowTable.ingest)
    def ingest(self, rec) -> None:
        t, dp, inport, src, dst, outport, packets, nbytes = rec
        i = self._index.get((dp, src, dst))
        d = 0
        if i is None:
            i = self._index.get((dp, dst, src))
            d = 1
        if i is not None:
            if i in self._pend or i in self._new_rows:
                self.flush()
            self._pend[i] = (d, float(packets), float(nbytes), float(t))
            self._updated[i] = self._updated.get(i, 0) | (1 << d)
            self._dirty = True
            return
        i = len(self._meta)
        self._index[(dp, src, dst)] = i
        self._meta.append((dp, inport, src, dst, outport))
        row = np.zeros(STATE)
        row[FWD + 0], row[FWD + 1] = packets, nbytes
        row[FWD + 8] = row[REV + 8] = row[T0] = t
        self._new_rows[i] = row
        self._dirty = True

    def flush(self):
        """apply the queued samples: new flows' initial rows are copied in, then ONE tcsdn_flow_update over all flows"""
        torch = self._torch
        n = len(self._meta)
        if n > self._dstate.shape[0]:
            grown = torch.zeros((max(n, 2 * self._dstate.shape[0]), STATE), dtype=torch.float64, device=self._dev)
            grown[: self._dstate.shape[0]] = self._dstate
            self._dstate = grown
        if self._new_rows:
            idx = torch.tensor(sorted(self._new_rows), dtype=torch.int64, device=self._dev)
            rows = torch.from_numpy(np.stack([self._new_rows[i] for i in sorted(self._new_rows)])).to(self._dev)
            self._dstate[idx] = rows
            self._new_rows.clear()
        if n == 0:
            return
        direction = np.full(n, 2, np.uint8)
        packets, nbytes, now = np.zeros(n), np.zeros(n), np.zeros(n)
        for i, (d, p, b, t) in self._pend.items():
            direction[i], packets[i], nbytes[i], now[i] = d, p, b, t
        self._pend.clear()
        if self._feat is None or self._feat.shape[0] < n:
            self._feat = torch.empty((max(n, self._dstate.shape[0]), 12), dtype=self._fdtype, device=self._dev)
        args = [torch.from_numpy(a).to(self._dev) for a in (packets, nbytes, now)]
        dd = torch.from_numpy(direction).to(self._dev)
        with torch.cuda.device(self._dev):
            self._lib.check(self._lib.load().tcsdn_flow_update(
                self._dstate.data_ptr(), args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), dd.data_ptr(), n,
                self._feat.data_ptr(), self._lib.F32 if self._fdtype == torch.float32 else self._lib.F64,
                torch.cuda.current_stream().cuda_stream))
        self._dirty = False

    # ---- views
    def features_device(self):
        """[n_flows, 12] CUDA tensor in the order of reference :104 (valid until the next ingest)"""
        if self._dirty or self._pend or self._new_rows:
            self.flush()
        return self._feat[: len(self._meta)]

    @property
    def state(self) -> np.ndarray:
        self.features_device()
        return self._dstate[: len(self._meta)].cpu().numpy()

    def features(self, dtype=np.float64) -> np.ndarray:
        return self.features_device().cpu().numpy().astype(dtype, copy=False)

    def rows(self):
        st = self.state
        for i, (dp, _inport, src, dst, _out) in enumerate(self._meta):
            # a direction keeps its initial status (forward ACTIVE :46, reverse INACTIVE :58) until its first update (:75-78, :93-96)
            mask = self._updated.get(i, 0)
            fwd = not (st[i, FWD + 3] == 0 or st[i, FWD + 2] == 0) if mask & 1 else True
            rev = not (st[i, REV + 3] == 0 or st[i, REV + 2] == 0) if mask & 2 else False
            yield i, src, dst, "ACTIVE" if fwd else "INACTIVE", "ACTIVE" if rev else "INACTIVE"
