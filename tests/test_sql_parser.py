"""The SELECT parser of B200SQLEngine (host logic): SQL text -> column expressions (fugue_b200.column)."""
from pytest import raises

from fugue_b200.sql import StructuredRawSQL, _parse_select


def parse(items, rest):
    return _parse_select(items, rest, f"SELECT {items} FROM {rest}")


def test_group_by_having_order_limit():
    st = parse("key, SUM(v0) AS s, COUNT(*) AS c", "t GROUP BY key HAVING SUM(v0) >= 7 OR key = 1 ORDER BY s DESC, key LIMIT 3")
    assert [str(c) for c in st.columns] == ["key", "SUM(v0) AS s", "COUNT(*) AS c"]
    assert st.table == "t" and [str(g) for g in st.group_by] == ["key"]
    assert str(st.having) == "|(>=(SUM(v0),7),==(key,1))"
    assert st.order_by == [("s", False), ("key", True)] and st.limit == 3 and not st.distinct


def test_expressions_and_aliases():
    st = parse("DISTINCT a.key k, (v0 + 1.5e2) * -2 AS w, CAST(v1 AS long) AS q, COALESCE(x, 0) c, 'it''s' AS s, "
               "`odd name`, COUNT(*), MAX(v0), COUNT(DISTINCT key) AS d",
               "`my table` AS a WHERE NOT (key > 3 AND v0 IS NOT NULL) OR key IN (1,2) OR v1 NOT BETWEEN 0 AND 1")
    assert st.distinct and st.table == "my table"
    assert [str(c) for c in st.columns] == [
        "key AS k", "*(+(v0,150.0),-2) AS w", "CAST(v1 AS long) AS q", "COALESCE(x,0) AS c", "'it\\'s' AS s",
        "odd name", "COUNT(*) AS count", "MAX(v0) AS v0", "COUNT(DISTINCT key) AS d"]
    assert str(st.where) == "|(|(~(&(>(key,3),NOT_NULL(v0))),|(==(key,1),==(key,2))),~(&(>=(v1,0),<=(v1,1))))"


def test_precedence():
    st = parse("a + b * c - d / 2 AS x, -a * 2 AS y, NOT a = 1 AND b <> 2 OR c AS z, a - -1 AS w", "t")
    assert [str(c) for c in st.columns] == [
        "-(+(a,*(b,c)),/(d,2)) AS x", "*(-(a),2) AS y", "|(&(~(==(a,1)),!=(b,2)),c) AS z", "-(a,-1) AS w"]


def test_rejections():
    with raises(NotImplementedError):
        parse("key", "(SELECT * FROM t)")
    with raises(NotImplementedError):
        parse("key", "t UNION SELECT key FROM u")
    with raises(NotImplementedError):
        parse("key", "t LIMIT x")
    with raises(ValueError):
        parse("key", "t WHERE SUM(v) > 1")
    with raises(NotImplementedError):
        parse("key ; DROP", "t")


def test_structured_raw_sql_pieces():
    st = StructuredRawSQL([(False, "SELECT * FROM"), (True, "a"), (False, " WHERE x>1 ")])
    assert st.construct() == "SELECT * FROM a WHERE x>1"
    assert st.construct({"a": "tbl"}) == "SELECT * FROM tbl WHERE x>1"
    st = StructuredRawSQL.from_expr("SELECT * FROM <tmpdf:abc> WHERE x<3")
    assert st.construct(lambda n: n.upper()) == "SELECT * FROM ABC WHERE x<3"
