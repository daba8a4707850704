"""Import helper for the vendored, UNMODIFIED reference modules under oracle/_ref (built by oracle/build_ref.py).
TEST / BASELINE INFRASTRUCTURE ONLY: used by tests/, bench.py's cpu_baseline / `--impl reference` / eager-GPU legs.

`load()` puts oracle/_ref first on sys.path and imports the reference's `train`, `transformer`, `bar_distribution`,
`utils`, `encoders`, `priors` under their own top-level names (they import each other by those names).  matplotlib is
not installed; priors/utils.py imports it at module level for a plotting helper (priors/utils.py:10-11), so empty
stand-in modules are registered first -- the reference files themselves are untouched.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_NAMES = ("utils", "bar_distribution", "transformer", "encoders", "positional_encodings", "decoders", "priors", "train")


def available():
    return os.path.isfile(os.path.join(REF_DIR, "train.py"))


def load():
    """-> dict name -> reference module.  Raises if oracle/_ref has not been built."""
    if not available():
        raise RuntimeError("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    # the reference modules use bare top-level names; make sure THEY are found first and not this repo's drop-ins
    stale = [n for n in list(sys.modules) if n.split(".")[0] in _NAMES
             and not getattr(sys.modules[n], "__file__", "").startswith(REF_DIR)]
    for n in stale:
        del sys.modules[n]
    if sys.path[0] != REF_DIR:
        if REF_DIR in sys.path:
            sys.path.remove(REF_DIR)
        sys.path.insert(0, REF_DIR)
    return {n: importlib.import_module(n) for n in _NAMES}


class StepTimer:
    """Wraps a reference DataLoader class so that the wall-clock time of `steps` full training steps (data + forward +
    backward + optimizer, exactly the loop of train.py:64-99) can be read after `train()` returns: a timestamp is taken
    (after a device sync when CUDA is in use) each time the loop asks for the next batch."""

    def __init__(self, dl_class, sync=None):
        import time
        stamps = self.stamps = []

        class Timed(dl_class):
            def __iter__(inner):
                def gen():
                    for item in dl_class.__iter__(inner):
                        if sync is not None:
                            sync()
                        stamps.append(time.perf_counter())
                        yield item
                    if sync is not None:
                        sync()
                    stamps.append(time.perf_counter())
                return gen()
        Timed.__name__ = dl_class.__name__
        self.cls = Timed

    def seconds(self, warmup, steps):
        """stamps[i] is taken when batch i is handed out (i.e. after it was sampled), so stamps[i+1] - stamps[i] =
        compute of step i + sampling of batch i+1 = one full step.  Needs steps_per_epoch >= warmup + steps + 1."""
        return self.stamps[warmup + steps] - self.stamps[warmup]
