"""Engine lifecycle and engine resolution: context engine, global engine, global conf.

The part of the reference's ``ExecutionEngine`` base class that is neither compute nor IO
(fugue/execution/execution_engine.py:50-90, 351-439, 1194-1212) and of its factory
(fugue/execution/factory.py:224-339, fugue/constants.py:37-70):

* an engine can be THE global engine (``set_global``) and / or a context engine (``as_context``, nestable,
  ContextVar based: safe across threads and async tasks); the global engine counts as being in a context;
* ``on_enter_context`` / ``on_exit_context`` hooks fire around every entry and exit; when the count of
  contexts an engine is in drops to 0 the engine is stopped; ``stop`` calls ``stop_engine`` at most once;
* resolution order for ``engine=None``: context engine, global engine, inference from the inputs, default;
* ``register_global_conf``: base configs every engine picks up.

Pure Python (no torch, no CUDA): the device engine inherits ``EngineLifecycle``; the CPU tests drive it with a
stand-in engine class.
"""
import contextvars
import threading
from contextlib import contextmanager
from typing import Any, Dict, Iterator, Optional

_CONTEXT_ENGINE: "contextvars.ContextVar[Optional[EngineLifecycle]]" = contextvars.ContextVar(
    "fugue_b200_engine", default=None)
_LOCK = threading.RLock()

FUGUE_GLOBAL_CONF: Dict[str, Any] = {
    "fugue.workflow.concurrency": 1,
    "fugue.default.partitions": -1,
}


class FugueInvalidOperation(Exception):
    """fugue/exceptions.py ``FugueInvalidOperation``."""


def register_global_conf(conf: Dict[str, Any], on_dup: str = "overwrite") -> None:
    """Base configs for engines created from now on (fugue/constants.py:56-70).  ``on_dup``: "overwrite",
    "ignore" (keep the old value) or "throw" - a ValueError when a key exists with a DIFFERENT value, and then
    nothing of ``conf`` is registered (tests/fugue/execution/test_execution_engine.py:64-85)."""
    if on_dup == "throw":
        clash = {k: v for k, v in conf.items() if k in FUGUE_GLOBAL_CONF and FUGUE_GLOBAL_CONF[k] != v}
        if clash:
            raise ValueError(f"global conf already set with different values: {sorted(clash)}")
    for k, v in conf.items():
        if k in FUGUE_GLOBAL_CONF and on_dup == "ignore":
            continue
        FUGUE_GLOBAL_CONF[k] = v


class _GlobalEngine:
    """The one global engine; replacing it takes the old one out of its context (which may stop it)."""

    def __init__(self) -> None:
        self._engine: Optional["EngineLifecycle"] = None

    def set(self, engine: Optional["EngineLifecycle"]) -> None:
        with _LOCK:
            old = self._engine
            if old is not None:
                old._lc()["is_global"] = False
                old._exit_context()
            self._engine = engine
            if engine is not None:
                engine._enter_context()
                engine._lc()["is_global"] = True

    def get(self) -> Optional["EngineLifecycle"]:
        return self._engine


_GLOBAL_ENGINE = _GlobalEngine()


class EngineLifecycle:
    """Mixin with the context / global / stop protocol.  Keeps its state in one lazily created dict so that a
    subclass does not have to call an ``__init__`` of this class."""

    def _lc(self) -> Dict[str, Any]:
        st = self.__dict__.get("_lifecycle_state")
        if st is None:
            st = dict(count=0, is_global=False, stopped=False, lock=threading.RLock())
            self.__dict__["_lifecycle_state"] = st
        return st

    # ---- hooks (override) -------------------------------------------------------------------
    def on_enter_context(self) -> None:
        return

    def on_exit_context(self) -> None:
        return

    def stop_engine(self) -> None:
        return

    # ---- protocol ---------------------------------------------------------------------------
    @property
    def in_context(self) -> bool:
        with _LOCK:
            return self._lc()["count"] > 0

    @property
    def is_global(self) -> bool:
        return bool(self._lc()["is_global"])

    def stop(self) -> None:
        st = self._lc()
        with st["lock"]:
            if not st["stopped"]:
                self.stop_engine()
                st["stopped"] = True

    def set_global(self) -> "EngineLifecycle":
        _GLOBAL_ENGINE.set(self)
        return self

    @contextmanager
    def as_context(self) -> Iterator["EngineLifecycle"]:
        with _LOCK:
            self._enter_context()
            token = _CONTEXT_ENGINE.set(self)
        try:
            yield self
        finally:
            with _LOCK:
                _CONTEXT_ENGINE.reset(token)
                self._exit_context()

    def _enter_context(self) -> None:
        self.on_enter_context()
        self._lc()["count"] += 1

    def _exit_context(self) -> None:
        st = self._lc()
        st["count"] -= 1
        self.on_exit_context()
        if st["count"] == 0:
            self.stop()

    def __copy__(self) -> "EngineLifecycle":   # engines are shared, never copied (execution_engine.py:1176-1180)
        return self

    def __deepcopy__(self, memo: Any) -> "EngineLifecycle":
        return self


def try_get_context_engine() -> Optional[EngineLifecycle]:
    """Context engine if any, else the global engine, else None (factory.py:224-234)."""
    engine = _CONTEXT_ENGINE.get()
    return engine if engine is not None else _GLOBAL_ENGINE.get()


def clear_global_engine() -> None:
    _GLOBAL_ENGINE.set(None)


def get_current_conf() -> Dict[str, Any]:
    """Conf of the context / global engine, else the global conf (fugue/execution/api.py:104-111)."""
    engine = try_get_context_engine()
    return engine.conf if engine is not None else FUGUE_GLOBAL_CONF  # type: ignore
