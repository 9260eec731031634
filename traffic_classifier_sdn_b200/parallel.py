"""Row-sharded data parallelism over the GPUs of one box.

Rows are independent and the models are tiny, so the path shards by rows only: rank r of W classifies the
contiguous block [r*ceil(n/W), (r+1)*ceil(n/W)) with its own replica of the model, and there is NO collective
on the data path.  The only exchange the reference's use could want is the full label vector on every rank:
one all-gather of int32 class indices (NCCL over NVLink when the tensors are CUDA tensors, gloo on CPU).

One process per GPU (torchrun); ``torch.distributed`` is plumbing, the kernels never see it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block of rank `rank`: (start, stop).  Blocks differ by at most ceil-rounding; empty blocks allowed."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    per = -(-n // world)
    start = min(n, rank * per)
    return start, min(n, start + per)


class Communicator:
    """The library's own NCCL communicator (include/tcsdn.h: tcsdn_comm_init / tcsdn_allgather_labels).

    Collective constructor: every rank of the ``torch.distributed`` group calls it; rank 0 creates the NCCL unique
    id and the group's object broadcast carries it to the others (any transport would do -- the C ABI only wants
    the 128 bytes).  The current CUDA device must already be this rank's GPU."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._lib = _lib.load()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        buf = C.create_string_buffer(_lib.COMM_ID_BYTES)
        if self.rank == 0:
            _lib.check(self._lib.tcsdn_comm_unique_id(buf))
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._h = C.c_void_p()
        _lib.check(self._lib.tcsdn_comm_init(self.rank, self.world, C.c_char_p(box[0]), C.byref(self._h)))

    def allgather_labels(self, local, n_block: int, n_classes: "int | None" = None, out=None):
        """local: CUDA int32 tensor with <= n_block entries -> CUDA int32 tensor [world * n_block] (short shards padded -1).
        With ``n_classes`` <= 255 the labels cross the wire as bytes (tcsdn_allgather_labels_u8); ``out`` lets the
        caller keep one result buffer (needed when the call is captured into a CUDA graph)."""
        import torch
        if out is None:
            out = torch.empty(self.world * n_block, dtype=torch.int32, device=local.device)
        st = torch.cuda.current_stream(local.device).cuda_stream
        lp = C.c_void_p(local.data_ptr() if local.numel() else 0)
        if n_classes is not None:
            _lib.check(self._lib.tcsdn_allgather_labels_u8(self._h, lp, local.numel(), n_block, C.c_void_p(out.data_ptr()),
                                                           int(n_classes), C.c_void_p(st)))
        else:
            _lib.check(self._lib.tcsdn_allgather_labels(self._h, lp, local.numel(), n_block, C.c_void_p(out.data_ptr()),
                                                        C.c_void_p(st)))
        return out

    # ---- the exchange fused into the classification (include/tcsdn.h: tcsdn_comm_gather_buffer / tcsdn_predict_gathered)
    def gather_buffer(self, n_block: int) -> int:
        """Collective: size the peer-memory label buffers for blocks of up to n_block rows.  -> bytes per rank slot."""
        ptr, slot = C.c_void_p(), C.c_int64(0)
        _lib.check(self._lib.tcsdn_comm_gather_buffer(self._h, int(n_block), C.byref(ptr), C.byref(slot)))
        self._slot = int(slot.value)
        return self._slot

    def predict_gathered(self, model, X):
        """Classify this rank's block X (CUDA tensor [n_local, d]) with the labels of ALL ranks as the result: a CUDA uint8
        tensor [world, slot_bytes] (255 = padding) that aliases the library's buffer -- valid until the call after the next.
        Every rank calls it; LogisticRegression / GaussianNB / KMeans store into the peers from inside their kernel.
        Blocks of different lengths (a short last shard): call ``gather_buffer(n_block)`` with the common block size first."""
        import torch
        model._check_fitted()
        if not getattr(self, "_slot", 0) or X.shape[0] > self._slot:
            self.gather_buffer(X.shape[0])
        X = X.contiguous()
        out = C.c_void_p()
        with torch.cuda.device(X.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(self._lib.tcsdn_predict_gathered(model._handle, self._h, C.c_void_p(X.data_ptr()), X.shape[0], X.shape[1],
                                                        _lib.F32 if X.dtype == torch.float32 else _lib.F64, C.byref(out), C.c_void_p(st)))

        class _View:   # zero-copy view of the library's device buffer
            __cuda_array_interface__ = {"shape": (self.world, self._slot), "typestr": "|u1", "data": (int(out.value), False), "version": 2}
        return torch.as_tensor(_View(), device=X.device)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.tcsdn_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def predict_sharded(model, X, gather: bool = True, group=None, comm: "Communicator | None" = None):
    """Classify this rank's block of X (every rank holds the same X, or at least its own block's rows).

    Returns the int32 class indices of the whole batch on every rank when ``gather`` (one all-gather of the
    per-rank label vectors, padded to equal length), else only this rank's block.  Works with numpy arrays
    (gloo) and CUDA tensors (NCCL: through ``comm`` -- the library's own communicator -- when one is passed,
    else through torch.distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return model.predict_indices(X)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n = len(X)
    a, b = shard_bounds(n, rank, world)
    local = model.predict_indices(X[a:b]) if b > a else (X.new_empty(0, dtype=torch.int32) if torch.is_tensor(X)
                                                         else np.empty(0, np.int32))
    if not gather:
        return local
    per = -(-n // world)
    if comm is not None and torch.is_tensor(local) and local.is_cuda:
        return comm.allgather_labels(local.contiguous(), per, n_classes=len(model.classes_))[:n]
    is_t = torch.is_tensor(local)
    t = local if is_t else torch.from_numpy(np.ascontiguousarray(local))
    pad = torch.zeros(per, dtype=torch.int32, device=t.device)   # equal-sized contributions for the collective
    pad[: t.numel()] = t
    out = torch.empty(per * world, dtype=torch.int32, device=t.device)
    if t.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:
        dist.all_gather(list(out.split(per)), pad, group=group)
    # blocks are contiguous and only the trailing ranks can be short, so dropping the tail removes all padding
    res = out[:n]
    return res if is_t else res.numpy()
