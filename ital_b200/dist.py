"""Host-side plumbing for the multi-GPU learner: one process per GPU, rows sharded in contiguous blocks.

Nothing crosses GPUs except fixed-size point records (a few KB): per labelled point one summed export, per
greedy step one all-gather of every shard's best candidate.  torch.distributed is used only
to move those records; with the NCCL backend they travel through device tensors over NVLink, with gloo (CPU
tests) through host tensors.
"""
import numpy as np


def partition_rows(n_rows, world_size):
    """Contiguous row blocks, sizes differing by at most one: returns world_size + 1 offsets."""
    base, extra = divmod(int(n_rows), int(world_size))
    sizes = [base + (1 if r < extra else 0) for r in range(world_size)]
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)


def pick_winner(records):
    """Index of the best record: highest score, ties to the lowest global row (np.argmax on the ascending
    candidate list, ital/ital.py:98,130).  records[:, 0] = global row (-1 = empty), records[:, 1] = score.
    Returns -1 if every record is empty."""
    best = -1
    for r in range(len(records)):
        idx, score = records[r][0], records[r][1]
        if idx < 0 or score != score:
            continue
        if best < 0 or score > records[best][1] or (score == records[best][1] and idx < records[best][0]):
            best = r
    return best


def merge_top(comm, idx, val, width, want):
    """Merge the shards' descending (row, mean) lists into the global top list: mean descending, ties by ascending
    row (ActiveRetrievalBase.top_results, ital/retrieval_base.py:64-75).  `width` = longest local list any shard may
    send, `want` = entries wanted (None = all)."""
    pad = np.full((2, width), -1.0)
    pad[0, :len(idx)] = idx
    pad[1, :len(idx)] = val
    allv = comm.gather_records(pad)
    gi, gv = allv[:, 0, :].reshape(-1), allv[:, 1, :].reshape(-1)
    keep = gi >= 0
    gi, gv = gi[keep].astype(np.int64), gv[keep]
    order = np.lexsort((gi, -gv))
    return gi[order] if want is None else gi[order[:want]]


class LocalComm(object):
    """Single process: collectives are identities."""
    rank, world_size = 0, 1

    def sum_records(self, rec):
        return rec

    def gather_records(self, rec):
        return rec[None, :]

    def gather_rows(self, local, offsets):
        return local

    def barrier(self):
        pass


class TorchComm(object):
    """torch.distributed process group; NCCL moves CUDA tensors, gloo moves CPU tensors."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self._torch, self._dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        backend = dist.get_backend(group)
        self.device = torch.device('cuda', device if device is not None else torch.cuda.current_device()) \
            if backend == 'nccl' else torch.device('cpu')
        self.on_device = self.device.type == 'cuda'
        self._bufs = {}

    def device_buffers(self, rec_len):
        """(one record, world_size records) as float64 CUDA tensors, reused across fetches."""
        if rec_len not in self._bufs:
            self._bufs[rec_len] = (self._torch.zeros(rec_len, dtype=self._torch.float64, device=self.device),
                                   self._torch.zeros(rec_len * self.world_size, dtype=self._torch.float64,
                                                     device=self.device))
        return self._bufs[rec_len]

    def all_gather_device(self, out, rec):
        """NCCL all-gather of device-resident records, ordered after the kernels that wrote `rec` and before
        the kernels that read `out` (same CUDA stream semantics as any torch collective)."""
        self._dist.all_gather_into_tensor(out, rec, group=self.group)

    def _t(self, a):
        return self._torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def sum_records(self, rec):
        """Complete records from per-shard exports (non-owners contribute zeros)."""
        t = self._t(rec)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def gather_records(self, rec):
        t = self._t(rec).reshape(-1)
        out = self._torch.empty(self.world_size * t.numel(), dtype=t.dtype, device=self.device)
        self._dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().numpy().reshape((self.world_size,) + tuple(np.shape(rec)))

    def gather_bytes(self, mine):
        """All ranks' fixed-size byte strings (numpy uint8), by rank: (world_size, len) array."""
        t = self._torch.from_numpy(np.ascontiguousarray(mine, dtype=np.uint8)).to(self.device)
        out = self._torch.empty(self.world_size * t.numel(), dtype=t.dtype, device=self.device)
        self._dist.all_gather_into_tensor(out, t, group=self.group)
        return out.cpu().numpy().reshape(self.world_size, -1)

    def all_agree(self, ok):
        """True iff `ok` is true on every rank."""
        t = self._torch.tensor([1 if ok else 0], dtype=self._torch.int32, device=self.device)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MIN, group=self.group)
        return bool(int(t.item()))

    def barrier(self):
        self._dist.barrier(group=self.group)

    def gather_rows(self, local, offsets):
        """Concatenate per-shard row vectors (ragged by at most one row) into the global vector."""
        width = int(np.max(np.diff(offsets)))
        pad = np.zeros(width, dtype=np.float64)
        pad[:len(local)] = local
        allv = self.gather_records(pad)
        return np.concatenate([allv[r, :offsets[r + 1] - offsets[r]] for r in range(self.world_size)])
