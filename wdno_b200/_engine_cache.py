"""Engine lifetime shared by the two denoisers.

An engine (`Unet3DEngine` / `Unet2DEngine`) is a frozen snapshot of the parameters re-packed into kernel tile order.
Parameters can change behind the module's back -- `parent.load_state_dict(...)` (torch calls the child's
`_load_from_state_dict`, never its `load_state_dict`), `ema.copy_params_...`, an optimiser step -- and all of them are
in-place writes, which bump `Tensor._version`.  `engine()` therefore compares a fingerprint (sum of parameter versions
and storage addresses) and rebuilds the plans when it moved, so a stale snapshot can never run.
"""


class EngineOwner:
    _engine = None
    _engine_fp = None

    def _make_engine(self):  # pragma: no cover - provided by the model class
        raise NotImplementedError

    def _fingerprint(self):
        v, a = 0, 0
        for p in self.parameters():
            v += p._version
            a ^= p.data_ptr()
        return v, a

    def invalidate(self):
        self._engine = None
        self._engine_fp = None

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self.invalidate()
        return super().load_state_dict(*a, **k)

    def engine(self):
        fp = self._fingerprint()
        if self._engine is not None and self._engine_fp != fp and self._engine_fp[1] == fp[1] and hasattr(self._engine, "refresh"):
            # same storage, new values (optimiser step, EMA copy_, nested load_state_dict): re-pack the tiles in place --
            # plans and captured graphs stay valid
            self._engine.refresh()
            self._engine_fp = fp
        if self._engine is None or self._engine_fp != fp:
            self._engine = self._make_engine()
            self._engine_fp = fp
        return self._engine


class GraphedRefresh:
    """Runs `fn` (a fixed sequence of device ops on static addresses) eagerly the first time and replays it from a CUDA graph
    afterwards.  `signature()` describes the set of buffers `fn` touches (e.g. how many packed-weight entries exist): when it
    changes (a plan for a new geometry was built) the graph is dropped and re-captured.  Falls back to eager execution for
    good if capture fails (e.g. an op that synchronises)."""

    def __init__(self, fn, signature=None):
        self.fn, self.signature = fn, signature
        self.graph, self.failed, self.sig, self.eager_runs = None, False, None, 0

    def __call__(self):
        import torch
        sig = self.signature() if self.signature is not None else None
        if sig != self.sig:
            self.graph, self.sig, self.eager_runs = None, sig, 0
        if self.failed or self.eager_runs == 0 or not torch.cuda.is_available() or torch.cuda.is_current_stream_capturing():
            self.eager_runs += 1
            return self.fn()
        if self.graph is None:
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self.fn()
                self.graph = g
            except Exception:  # noqa: BLE001
                self.failed = True
                torch.cuda.synchronize()
                return self.fn()
        self.graph.replay()
