"""Host-side sharding of independent games / trees over ranks (one process per GPU) and the two collectives of the path.

Games and search trees are independent units: rank r of R plays the global ids
    game_id0(step, r, R, n) ... + n - 1   =   (step * R + r) * n ...
and every random draw is keyed by the global id, so the union of all ranks' results for a step is exactly what one process
would produce for the same ids — results never depend on R.  The only exchanges are sum all-reduces of (a) int64 result /
win counters at report time and (b) the fp32 vector [gradient | loss numerator | position count] once per REINFORCE update.
On GPUs both run over NCCL through the library's own entry points (iago_comm_*: the all-reduce is enqueued on the engine's stream, the
Adam step reads the reduced count on the device — no host round trip; `Communicator` below).  torch.distributed launches the processes
and carries the 128-byte NCCL id; it is also the fallback transport (gloo) of the CPU tests of the host logic.
"""
import ctypes as C

import torch
import torch.distributed as dist


def world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def game_id0(step, rank, world_size, n_per_rank):
    """First global game id of `rank` in `step` when every rank plays n_per_rank games per step (weak scaling)."""
    return (step * world_size + rank) * n_per_rank


def shard_range(n_total, rank, world_size):
    """[lo, hi) of a fixed total of n_total units split over the ranks (strong scaling), remainder to the low ranks."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(t, group=None):
    """In-place sum over ranks; a no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def all_reduce_max_(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def reduce_counters(counters, device=None, group=None):
    """dict name -> int, summed over ranks (wins / draws / losses / plies / games / playouts)."""
    keys = sorted(counters)
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=device)
    all_reduce_sum_(t, group)
    return dict(zip(keys, t.tolist()))


def mean_gradient_(flat, n_params, group=None):
    """flat = [sum-gradient (n_params) | loss numerator | position count] of this rank. All-reduces it and returns
    (mean loss, total count); flat[:n_params] / count is then the gradient of the reference's mean loss over ALL ranks' positions."""
    all_reduce_sum_(flat, group)
    num, count = flat[n_params:n_params + 2].tolist()
    return (num / count if count > 0 else 0.0), count


def root_parallel_moves(visits, group=None):
    """One game searched on several GPUs at once (the optional exchange for a single tree): every rank searches the
    same root with its own global tree id — the rollouts draw from different Philox streams — and the root visit counts
    [n_trees, 65] (index 64 = the pass child) are summed over the ranks.  Returns (summed visits, best action per tree) with the
    reference's tie rule (MCTS.py:147: the first maximum, i.e. the lowest action; pass = -1 only when it is the only child)."""
    v = torch.as_tensor(visits).to(torch.int64).clone()
    all_reduce_sum_(v, group)
    best = torch.argmax(v[:, :64], dim=1)
    only_pass = (v[:, :64].sum(dim=1) == 0) & (v[:, 64] > 0)
    best = torch.where(only_pass, torch.full_like(best, -1), best)
    return v, best


def barrier(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=group)


def broadcast_object(obj, group=None, src=0):
    """`obj` of rank `src` on every rank (host-side control decisions such as the opponent file of a REINFORCE set)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


class Communicator:
    """An NCCL communicator owned by libiago_b200.so (include/iago_b200.h iago_comm_*).  Collectives are enqueued on the engine's
    current stream and never synchronise the host."""

    def __init__(self, engine, rank, world, unique_id):
        from ._lib import check
        self.eng, self.lib, self.rank, self.world = engine, engine.lib, int(rank), int(world)
        h = C.c_void_p()
        check(self.lib.iago_comm_create(engine.ctx, C.c_char_p(unique_id), self.rank, self.world, C.byref(h)))
        self.h = h

    @staticmethod
    def unique_id(engine):
        from ._lib import check
        buf = C.create_string_buffer(128)
        check(engine.lib.iago_comm_unique_id(buf, 128))
        return buf.raw

    @classmethod
    def from_process_group(cls, engine, group=None):
        """One communicator per process of an initialised torch.distributed group: rank 0 makes the id, the group carries it."""
        rank, size = world(group)
        uid = broadcast_object(cls.unique_id(engine) if rank == 0 else None, group)
        return cls(engine, rank, size, uid)

    def all_reduce_sum_(self, t):
        from ._lib import check
        assert t.is_cuda and t.is_contiguous()
        fn = {torch.float32: self.lib.iago_comm_allreduce_sum_f32, torch.int64: self.lib.iago_comm_allreduce_sum_i64}[t.dtype]
        check(fn(self.h, C.c_void_p(t.data_ptr()), t.numel(), self.eng._stream(None)))
        return t

    def close(self):
        if getattr(self, "h", None):
            self.lib.iago_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
