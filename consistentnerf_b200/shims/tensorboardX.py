"""tensorboardX stand-in: SummaryWriter appending scalars to <logdir>/scalars.jsonl.

add_scalar(tag, tensor, step) converts the value to a Python float -- one host synchronisation per call, exactly the cost the
original has (NP/run_nerf_view.py:1908-1937 makes ten-odd such calls per step; pipeline.StepLog is the sync-free alternative)."""
import json
import os


class SummaryWriter:
    def __init__(self, logdir=None, *_a, **_k):
        self.logdir = logdir or "runs"
        os.makedirs(self.logdir, exist_ok=True)
        self._f = open(os.path.join(self.logdir, "scalars.jsonl"), "a")

    def add_scalar(self, tag, value, global_step=None, *_a, **_k):
        try:
            v = float(value)
        except Exception:
            v = float("nan")
        self._f.write(json.dumps({"tag": tag, "value": v, "step": None if global_step is None else int(global_step)}) + "\n")

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()

    def __getattr__(self, name):            # add_image, add_histogram, ...: accepted and dropped
        if name.startswith("add_"):
            return lambda *a, **k: None
        raise AttributeError(name)
