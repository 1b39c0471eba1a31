"""Data-parallel plumbing of the learners (one process per GPU, torch.distributed).

The reference has no data parallelism; the B200 design shards envs and replay per
rank and exchanges exactly one thing per update: the flat gradient arena, summed over ranks
before the global-norm clip and AdamW (which then apply grad_scale = 1/world, so every rank takes
the step of the mean gradient of the concatenated batch and parameters stay identical).
"""
import torch
import torch.distributed as dist


def world_size(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def allreduce_sum_(flat_grad, group=None):
    """In-place sum of the flat gradient arena over the ranks (NCCL over NVLink on GPUs, gloo in
    the CPU tests).  The mean is taken inside the optimiser kernel (grad_scale)."""
    if world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def broadcast_(flat_params, src=0, group=None):
    """Make rank ``src``'s parameters everyone's (initialisation)."""
    if world_size(group) > 1:
        dist.broadcast(flat_params, src=src, group=group)
    return flat_params


def params_in_sync(flat_params, group=None, atol=0.0):
    """True when every rank holds the same parameters (cheap periodic assertion)."""
    if world_size(group) <= 1:
        return True
    lo, hi = flat_params.clone(), flat_params.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool((hi - lo).abs().max().item() <= atol)


def merge_moments(mean, var, count, group=None):
    """One (mean, var, count) for the whole job out of every rank's own: the ranks' statistics are
    folded in rank order with the reference's parallel-variance update
    (RunningMeanStd.update_from_moments, pql/utils/torch_util.py:91-103), so every rank ends with the
    same normaliser.  For statistics that are merged once (checkpoints, static
    normalisers); the per-env-step path is ``RunningMeanStd(process_group=...)``, which all-reduces
    the batch sums before the update."""
    w = world_size(group)
    if w <= 1:
        return mean, var, count
    n = mean.numel()
    mine = torch.cat([mean.reshape(-1).float(), var.reshape(-1).float(),
                      torch.tensor([float(count)], dtype=torch.float32, device=mean.device)])
    parts = [torch.empty_like(mine) for _ in range(w)]
    dist.all_gather(parts, mine, group=group)
    m, v, c = parts[0][:n].clone(), parts[0][n:2 * n].clone(), float(parts[0][2 * n])
    for t in parts[1:]:
        bm, bv, bc = t[:n], t[n:2 * n], float(t[2 * n])
        delta = bm - m
        tot = c + bc
        m_2 = v * c + bv * bc + delta ** 2 * c * bc / tot
        m, v, c = m + delta * bc / tot, m_2 / tot, tot
    return m.reshape(mean.shape), v.reshape(var.shape), c


class FusedExchange:
    """Symmetric-memory state of the fused all-reduce + optimiser kernel (pqlb_adamw_polyak_dp): this
    rank's gradient arena, the receive buffer of the reduced gradient and the control block (flags +
    per-block sums of squares) are allocated in peer-mapped memory and exchanged once, at plan
    construction (a collective: every rank builds its plans in the same order).

    GRID is the launch width on every rank.  The kernel's blocks spin on flags written by the peers,
    so they must not be able to starve the OTHER learner's kernels of SMs (two learners on two
    streams, ranks at different points of their schedules: a full-width spinning grid could block the
    kernel its peer is waiting for): 64 one-block-per-SM blocks leave 84 SMs to everything else."""

    GRID = 64
    GRID_SPLIT = 32        # split mode: the exchange is a launch of its own (pqlb_grad_exchange_dp), the optimiser follows
    FLAG_WORDS = 32

    def __init__(self, n, device, group=None, split=False):
        self.split = bool(split)
        if self.split:
            self.GRID = self.GRID_SPLIT
        import torch.distributed._symmetric_memory as symm_mem
        name = (group if group is not None else dist.group.WORLD).group_name
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError("fused exchange: at most 8 ranks (one NVSwitch domain)")
        self.grad = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.red = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.ctl = symm_mem.empty(self.FLAG_WORDS + self.world * self.GRID, dtype=torch.float32, device=device)
        for t in (self.grad, self.red, self.ctl):
            t.zero_()
        self.handles = [symm_mem.rendezvous(t, name) for t in (self.grad, self.red, self.ctl)]
        self.local = torch.zeros(8, dtype=torch.int64, device=device)   # [epoch, blocks done, 4 phase clocks (ns), -, -]
        torch.cuda.synchronize(device)
        dist.barrier(group)              # every rank's flags are zero before anybody can signal

    def phase_ns(self):
        """(exchanges, [wait for gradients, reduce + deliver, wait for slices, optimiser] mean ns of block 0)."""
        v = self.local.tolist()
        n = max(1, v[0])
        return v[0], [x / n for x in v[2:6]]

    def desc(self):
        from .. import _lib
        d = _lib.DpDesc()
        for r in range(self.world):
            d.grad_peers[r] = self.handles[0].buffer_ptrs[r]
            d.red_peers[r] = self.handles[1].buffer_ptrs[r]
            d.ctl_peers[r] = self.handles[2].buffer_ptrs[r]
        d.local = self.local.data_ptr()
        d.rank, d.world, d.grid = self.rank, self.world, self.GRID
        # NVLS (in-switch reduction) when the symmetric-memory handles carry multicast mappings; PQLB_DP_NVLS=0: per-peer loads
        import os
        mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.handles[:2]]
        mc = [p + int(getattr(h, "offset", 0) or 0) if p else 0 for p, h in zip(mc, self.handles[:2])]
        self.nvls = all(mc) and os.environ.get("PQLB_DP_NVLS", "1") != "0"
        d.grad_mc, d.red_mc = (mc[0], mc[1]) if self.nvls else (None, None)
        return d
