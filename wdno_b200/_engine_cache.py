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
