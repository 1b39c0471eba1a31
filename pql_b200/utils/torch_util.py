"""Device-resident observation normaliser: drop-in for RunningMeanStd of pql/utils/torch_util.py:69-114.

``update`` is one kernel launch (pqlb_rms_update: column statistics of the batch accumulated in
fp64, merged with the reference's update_from_moments arithmetic); ``mean`` / ``var`` are updated in
place (the reference rebinds new tensors every step), ``count`` is tracked on the host exactly like
the reference's Python float and mirrored in an fp64 device scalar for the kernel."""
import torch

from .. import _lib


class RunningMeanStd:
    def __init__(self, epsilon=1e-4, shape=(), device='cuda', process_group=None, data_parallel=False):
        """``process_group`` / ``data_parallel`` (B200 extension): with more than one rank, every
        ``update`` all-reduces the batch's column sums (2 * cols doubles) before merging, so all ranks apply
        the reference's update on the CONCATENATED batch and hold bit-identical statistics."""
        self.device = torch.device(device)
        self._group, self._world = process_group, 1
        if process_group is not None or (data_parallel and torch.distributed.is_available() and torch.distributed.is_initialized()):
            self._world = torch.distributed.get_world_size(process_group)
        if self.device.type != "cuda":
            raise RuntimeError("pql_b200.RunningMeanStd lives in GPU memory; there is no CPU path")
        _lib.load()
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        if len(shape) != 1:
            raise NotImplementedError("only flat observations are on the PQL path")
        self.mean = torch.zeros(shape, dtype=torch.float32, device=self.device)
        self.var = torch.ones(shape, dtype=torch.float32, device=self.device)
        self.epsilon = epsilon
        self.count = epsilon
        self._count_dev = torch.full((1,), float(epsilon), dtype=torch.float64, device=self.device)
        self._ws = None

    def _workspace(self, rows, cols):
        need = int(_lib.load().pqlb_rms_workspace_bytes(rows, cols))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
        return self._ws

    @torch.no_grad()
    def update(self, x):
        """torch_util.py:77-81 + update_from_moments (:91-103)."""
        cols = self.mean.numel()
        x = x.reshape(-1, cols)
        if x.device != self.device or x.dtype != torch.float32:
            x = x.to(device=self.device, dtype=torch.float32)
        if x.stride(1) != 1:
            x = x.contiguous()
        ws = self._workspace(x.shape[0], cols)
        with torch.cuda.device(self.device):
            if self._world > 1:
                if getattr(self, "_sums", None) is None:
                    self._sums = torch.zeros(2 * cols, dtype=torch.float64, device=self.device)
                _lib.call("pqlb_rms_moments", _lib.ptr(x), x.shape[0], cols, x.stride(0), _lib.ptr(self._sums), _lib.ptr(ws), ws.numel())
                torch.distributed.all_reduce(self._sums, group=self._group)      # every rank contributes the same row count
                rows = x.shape[0] * self._world
                _lib.call("pqlb_rms_apply", _lib.ptr(self._sums), rows, cols, _lib.ptr(self.mean), _lib.ptr(self.var),
                          _lib.ptr(self._count_dev))
                self.count = self.count + rows
                return
            _lib.call("pqlb_rms_update", _lib.ptr(x), x.shape[0], cols, x.stride(0), _lib.ptr(self.mean), _lib.ptr(self.var),
                      _lib.ptr(self._count_dev), _lib.ptr(ws), ws.numel())
        self.count = self.count + x.shape[0]

    @torch.no_grad()
    def normalize(self, x):
        """(x - mean) / sqrt(var + epsilon), fp32 (torch_util.py:83-85)."""
        cols = self.mean.numel()
        x2 = x.reshape(-1, cols)
        if x2.device != self.device or x2.dtype != torch.float32:
            x2 = x2.to(device=self.device, dtype=torch.float32)
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        out = torch.empty((x2.shape[0], cols), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("pqlb_actor_inputs", _lib.ptr(x2), x2.shape[0], cols, x2.stride(0), _lib.ptr(self.mean),
                      _lib.ptr(self.var), float(self.epsilon), 0, 0, _lib.ptr(out), cols, None, 0, None, 0.0, 0, 0)
        return out.reshape(x.shape)

    def unnormalize(self, x):
        raise NotImplementedError("RunningMeanStd.unnormalize is not on the PQL path")

    def update_from_moments(self, batch_mean, batch_var, batch_count):
        raise NotImplementedError("moments are merged inside pqlb_rms_update: call update(x)")

    def get_states(self, device=None):
        """Snapshot of (mean, var, epsilon).  ``update`` writes mean / var in place, so - unlike the
        reference, which rebinds fresh tensors every update (torch_util.py:91-103) - the live tensors
        must not be handed to a consumer that reads them later on another stream: clones are."""
        if device is not None and torch.device(device) != self.device:
            return self.mean.to(device), self.var.to(device), self.epsilon
        return self.mean.clone(), self.var.clone(), self.epsilon

    def load_state_dict(self, info):
        """(mean, var, count) as stored in the reference's checkpoints ('obs_rms', model_util.py:24-41;
        evaluator.py:115-117 saves get_states(), whose third entry is epsilon - kept as given)."""
        self.mean.copy_(torch.as_tensor(info[0], dtype=torch.float32).reshape(self.mean.shape))
        self.var.copy_(torch.as_tensor(info[1], dtype=torch.float32).reshape(self.var.shape))
        self.count = float(info[2])
        self._count_dev.fill_(self.count)
