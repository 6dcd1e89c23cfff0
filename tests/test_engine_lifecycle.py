"""Context engine / global engine / stop protocol and engine resolution (``fugue_b200/lifecycle.py`` +
the factory functions of ``fugue_b200.api``), driven with a stand-in engine on CPU.

Behaviours pinned by the reference: tests/fugue/execution/test_api.py:8-69 (hook order, ``in_context`` /
``is_global`` states, stop exactly once when an engine leaves its last context, conf visibility) and
tests/fugue/execution/test_factory.py:252-283 (nesting, resolution order, inference).
"""
import copy
import threading

import pytest

from fugue_b200 import api as fa
from fugue_b200 import lifecycle as L
from fugue_b200.lifecycle import EngineLifecycle


class StandIn(EngineLifecycle):
    """An engine without a device: records what the protocol does to it."""

    def __init__(self, conf=None, **kwargs):
        self.conf = {**L.FUGUE_GLOBAL_CONF, **dict(conf or {}), **kwargs}
        self.entered, self.exited, self.stops = [], [], 0

    def on_enter_context(self):
        self.entered.append(self.in_context)

    def on_exit_context(self):
        self.exited.append(self.in_context)

    def stop_engine(self):
        self.stops += 1

    def create_default_sql_engine(self):
        return ("default-sql", self)


class OtherStandIn(StandIn):
    pass


@pytest.fixture(autouse=True)
def _clean_registry():
    saved_e, saved_s, saved_c = dict(fa._ENGINE_FACTORIES), dict(fa._SQL_ENGINE_FACTORIES), dict(L.FUGUE_GLOBAL_CONF)
    fa.register_execution_engine("standin", lambda conf, **kw: StandIn(conf, **kw))
    fa.register_execution_engine("other", lambda conf, **kw: OtherStandIn(conf, **kw))
    yield
    fa.clear_global_engine()
    for table, saved in ((fa._ENGINE_FACTORIES, saved_e), (fa._SQL_ENGINE_FACTORIES, saved_s),
                         (L.FUGUE_GLOBAL_CONF, saved_c)):
        table.clear()
        table.update(saved)


def test_global_and_context_engines_step_by_step():
    assert fa.get_current_conf().get("fugue.x", 0) == 0
    fa.register_global_conf({"fugue.x": 1})
    assert fa.get_current_conf().get("fugue.x", 0) == 1
    with pytest.raises(L.FugueInvalidOperation):
        fa.get_context_engine()

    e = fa.set_global_engine(StandIn(), {"fugue.x": 2})
    assert (e.entered, e.exited) == ([False], []) and e.in_context and e.is_global
    assert fa.get_current_conf()["fugue.x"] == 2 and fa.get_context_engine() is e

    with fa.engine_context("other", {"fugue.x": 3}) as e2:
        assert fa.get_current_conf()["fugue.x"] == 3 and fa.get_context_engine() is e2
        assert e2.in_context and not e2.is_global
        with e.as_context():                                    # the global engine as an inner context
            assert (e.entered, e.exited, e.stops) == ([False, True], [], 0)
            assert fa.get_context_engine() is e and fa.get_current_conf()["fugue.x"] == 2
            assert e2.in_context and e.in_context and e.is_global
        assert (e.entered, e.exited) == ([False, True], [True])
        assert fa.get_context_engine() is e2 and e.in_context and e.is_global
    assert (e2.in_context, e2.is_global, e2.stops) == (False, False, 1)   # left its last context: stopped
    assert e.stops == 0 and fa.get_current_conf()["fugue.x"] == 2

    e3 = fa.set_global_engine("standin", {"fugue.x": 4})        # replacing the global engine stops the old one
    assert (e.stops, e.entered, e.exited) == (1, [False, True], [True, False])
    assert not e.in_context and not e.is_global and e3.in_context and e3.is_global
    assert fa.get_current_conf()["fugue.x"] == 4
    fa.clear_global_engine()
    assert not e3.in_context and not e3.is_global and e3.stops == 1
    assert fa.get_current_conf().get("fugue.x", 0) == 1
    with pytest.raises(L.FugueInvalidOperation):
        fa.get_context_engine()
    with pytest.raises(ValueError):
        fa.set_global_engine(None)


def test_stop_runs_stop_engine_once():
    e = StandIn()
    for _ in range(3):
        e.stop()
    with e.as_context():
        pass
    assert e.stops == 1 and copy.copy(e) is e and copy.deepcopy(e) is e


def test_nested_contexts_and_resolution_order():
    e1, e2 = StandIn(), OtherStandIn()
    with e2.as_context():
        assert (e1.in_context, e2.in_context) == (False, True)
        with e1.as_context() as ex:
            assert ex is e1 and e1.in_context and e2.in_context
            got = fa.make_execution_engine(None, conf={"x": False})
            assert got is e1 and got.conf["x"] is False                  # context engine first, conf applied
        assert (e1.in_context, e2.in_context) == (False, True)
        assert fa.make_execution_engine(None, conf={"x": True}) is e2
    assert not e1.in_context and not e2.in_context and (e1.stops, e2.stops) == (1, 1)
    g = StandIn().set_global()
    assert fa.make_execution_engine() is g                               # then the global engine
    with OtherStandIn().as_context() as c:
        assert fa.make_execution_engine() is c
    assert fa.make_execution_engine("other") is not g                    # an explicit engine wins over both


def test_engine_specs():
    by_name = fa.make_execution_engine("standin", {"a": 1}, b=2)
    assert type(by_name) is StandIn and by_name.conf["a"] == 1 and by_name.conf["b"] == 2
    by_type = fa.make_execution_engine(OtherStandIn, {"a": 3})
    assert type(by_type) is OtherStandIn and by_type.conf["a"] == 3
    inst = StandIn({"a": 1})
    assert fa.make_execution_engine(inst, {"a": 5}, c=6) is inst and inst.conf["a"] == 5 and inst.conf["c"] == 6
    with pytest.raises(ValueError):
        fa.make_execution_engine("spark")
    with pytest.raises(TypeError):
        fa.make_execution_engine(123)
    with pytest.raises(TypeError):
        fa.make_execution_engine(dict)                                   # a type that is not an engine

    fa.register_sql_engine("s", lambda engine, **kw: ("s-sql", engine))
    pair = fa.make_execution_engine(("standin", "s"), {"a": 7})
    assert type(pair) is StandIn and pair.conf["a"] == 7 and pair._sql_engine == ("s-sql", pair)
    dflt = fa.make_execution_engine((inst, None))
    assert dflt is inst and inst._sql_engine == ("default-sql", inst)
    with pytest.raises(ValueError):
        fa.make_execution_engine(("standin", "nope"))
    with pytest.raises(KeyError):
        fa.register_sql_engine("s", lambda engine: None, on_dup="throw")
    fa.register_sql_engine("s", lambda engine, **kw: "ignored", on_dup="ignore")
    assert fa.make_sql_engine("s", inst) == ("s-sql", inst)


def test_context_engine_is_per_thread():
    seen = {}
    e = StandIn()

    def worker():
        seen["inside"] = L.try_get_context_engine()

    with e.as_context():
        t = threading.Thread(target=worker)   # a new thread starts from an empty context
        t.start()
        t.join()
        assert L.try_get_context_engine() is e
    assert seen["inside"] is None


def test_global_conf_is_the_base_of_every_engine_conf():
    """tests/fugue/execution/test_execution_engine.py:64-85."""
    fa.register_global_conf({"ftest.a": 1})
    assert StandIn().conf["ftest.a"] == 1 and StandIn({"ftest.a": 2}).conf["ftest.a"] == 2
    fa.register_global_conf({"ftest.a": 1, "ftest.b": 2}, on_dup="throw")     # same value: not a clash
    assert StandIn().conf["ftest.b"] == 2
    with pytest.raises(ValueError):
        fa.register_global_conf({"ftest.a": 2, "ftest.c": 3}, on_dup="throw")
    assert "ftest.c" not in L.FUGUE_GLOBAL_CONF and L.FUGUE_GLOBAL_CONF["ftest.a"] == 1   # all or nothing
    fa.register_global_conf({"ftest.a": 5, "ftest.d": 6}, on_dup="ignore")
    assert (L.FUGUE_GLOBAL_CONF["ftest.a"], L.FUGUE_GLOBAL_CONF["ftest.d"]) == (1, 6)


def test_sql_facet_encodes_table_names():
    from fugue_b200.sql import B200SQLEngine, StructuredRawSQL

    eng = StandIn()
    eng.log = "the-log"
    f1, f2 = B200SQLEngine(eng), B200SQLEngine(eng)
    assert f1.execution_engine is eng and f1.conf is eng.conf and f1.log == "the-log" and f1.dialect is None
    assert f1.encode_name("t") != f2.encode_name("t") and f1.encode_name("t").endswith("t")
    dfs, text = f1.encode({"a": 1, "b": 2}, StructuredRawSQL([(False, "SELECT * FROM"), (True, "a"), (False, "JOIN"),
                                                              (True, "b")]))
    assert set(dfs) == {f1.encode_name("a"), f1.encode_name("b")}
    assert text == f"SELECT * FROM {f1.encode_name('a')} JOIN {f1.encode_name('b')}"
    for call in (lambda: f1.table_exists("x"), lambda: f1.load_table("x"), lambda: f1.save_table(None, "x")):
        with pytest.raises(NotImplementedError):
            call()
