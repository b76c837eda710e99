"""Engine exceptions (reference: vllm/v1/engine/exceptions.py)."""


class EngineGenerateError(Exception):
    """A request-level failure; the engine is still healthy."""


class EngineDeadError(Exception):
    """The engine core process died or hit an unrecoverable error."""

    def __init__(self, *args, suppress_context: bool = False, **kwargs):
        super().__init__(*(args or ("EngineCore encountered an issue. See stack trace (above) for the root cause.",)),
                         **kwargs)
        self.__suppress_context__ = suppress_context
