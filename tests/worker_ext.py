"""Worker extension used by the shutdown / failure-propagation tests (``--worker-extension-cls``): reached on every
rank through ``collective_rpc``; arms a failure inside the model runner so the error starts where a real one would."""
import os


class FaultInjection:
    def whoami(self):
        return {"rank": self.rank, "pid": os.getpid(), "cls": type(self).__name__}

    def arm_failure(self, after_steps: int = 0, kind: str = "raise", only_rank=None):
        if only_rank is not None and self.rank != only_rank:
            return False
        runner, state = self.runner, {"n": 0}
        orig = runner.execute_model

        def wrapped(*a, **k):
            state["n"] += 1
            if state["n"] > after_steps:
                if kind == "exit":
                    os._exit(17)
                raise RuntimeError("injected forward failure")
            return orig(*a, **k)
        runner.execute_model = wrapped
        return True


class Clashing:
    def execute_model(self, *a, **k):      # collides with the worker's own method: must be rejected
        return None
