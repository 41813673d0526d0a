"""CUDA-graph replay of the streaming step (used by ``OnlineSKIRegression`` and ``ShardedOnlineSKIRegression``).

One streaming step (``experiments/regression.py:49-54``: ``evaluate`` + ``update``) issues ~300 kernels, most of them
tiny r x r / scalar ops.  Their launch cost, and the stall after every host read, is what bounds the step once the
panel kernels are sharded over several GPUs.  In graph mode ``evaluate`` and ``update`` each replay ONE captured
CUDA graph over static input buffers and read their scalar results back with a single device->host copy.

The captured sequence is exactly the eager code path (capture runs the same Python once; nothing executes during
capture, the first replay does the work), so everything that changes from step to step must live in device memory:
panels, caches, Adam state (``capturable=True``) and the observation counter.  Host reads are not capturable, hence
the interpolation bounds flags are collected (``ops._BOUNDS_SINK``) and returned with the results.
The two graphs share one private memory pool and are always replayed in capture order (evaluate, update).
"""
import warnings

import torch

from . import _lib, ops


class StepGraphs:
    def __init__(self, warmup_calls=2):
        self.warm0 = self.warm = int(warmup_calls)
        self.eval = self.upd = None
        self.q = None
        self.phase = None          # "evaluated" between a replayed evaluate() and its update()
        self.failed = False
        self.replays = 0
        self.launches = 0          # kernels of libwiski_b200 executed through replays
        self.x = self.y = self.stream = self.pool = None

    def setup(self, inputs, targets):
        self.q = inputs.shape[0]
        self.x = torch.empty_like(inputs).contiguous()
        self.y = torch.empty_like(targets).contiguous()
        self.stream = torch.cuda.Stream(device=inputs.device)
        self.pool = torch.cuda.graph_pool_handle()

    def load(self, inputs, targets):
        self.x.copy_(inputs)
        self.y.copy_(targets)

    def capture(self, fn):
        """Capture ``fn()`` (returns a tensor of scalar results) on the private stream / pool."""
        graph = torch.cuda.CUDAGraph()
        sink = []
        ops._BOUNDS_SINK = sink
        lib = _lib.load()
        l0 = lib.wiski_launch_count()
        try:
            with torch.cuda.graph(graph, pool=self.pool, stream=self.stream, capture_error_mode="thread_local"):
                res = fn()
                out = torch.cat([res.reshape(-1)] + [f.to(res.dtype).reshape(-1) for f, _, _ in sink])
        finally:
            ops._BOUNDS_SINK = None
        return {"graph": graph, "out": out, "n_res": res.numel(), "checks": [(x, spec) for _, x, spec in sink],
                "launches": int(lib.wiski_launch_count() - l0)}

    def replay(self, cap):
        cap["graph"].replay()
        self.replays += 1
        self.launches += cap["launches"]
        vals = cap["out"].tolist()                      # the one device->host read of the call
        for k, (x, spec) in enumerate(cap["checks"]):
            if vals[cap["n_res"] + k] != 0:
                ops._raise_out_of_bounds(x, spec)
        return vals[:cap["n_res"]]

    def fail(self, err):
        warnings.warn(f"CUDA-graph capture failed ({type(err).__name__}: {err}); continuing eagerly", RuntimeWarning)
        self.failed = True
        self.eval = self.upd = None


def make_adam_capturable(opt):
    """Device-side step counters, moments kept (torch.optim.Adam(capturable=True) semantics)."""
    for group in opt.param_groups:
        group["capturable"] = True
    for p, st in opt.state.items():
        if "step" in st and torch.is_tensor(st["step"]) and not st["step"].is_cuda:
            # same scalar dtype torch.optim uses for capturable state (the bias corrections are then evaluated on the
            # device in this precision: hyper-parameters follow the eager trajectory to ~1e-7 relative, not bit-exactly)
            sdt = torch.float64 if torch.get_default_dtype() == torch.float64 else torch.float32
            st["step"] = st["step"].to(device=p.device, dtype=sdt)
