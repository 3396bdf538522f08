"""TEST INFRASTRUCTURE ONLY -- stand-in for `tuneconfig` (progress bar + experiment
fan-out), which the reference imports at module level in tfmpc/solvers/ilqr.py."""
import contextlib as _contextlib
import sys as _sys
import types as _types


class _Bar:
    def __init__(self, n):
        self._n = n

    def __iter__(self):
        return iter(range(self._n))

    def set_postfix(self, **kw):
        pass


class Experiment:
    @staticmethod
    @_contextlib.contextmanager
    def trange(epochs, run_id=0, num_workers=1, desc="", show_progress=True, **kw):
        yield _Bar(int(epochs))


experiment = _types.ModuleType("tuneconfig.experiment")
experiment.Experiment = Experiment
_sys.modules["tuneconfig.experiment"] = experiment
