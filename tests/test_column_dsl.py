"""Column-expression DSL: the observable behaviour the reference's own tests pin
(tests/fugue/column/test_expressions.py, test_functions.py, test_sql.py - expected strings and types are
copied as literals; fugue itself is not importable here)."""
import pyarrow as pa
from pytest import raises

from fugue_b200.column import (SelectColumns, SQLExpressionGenerator, _BinaryOpExpr, _get_column_mentions,
                               all_cols, col, function, functions as f, lit, null, to_uuid)
from fugue_b200.schema import Schema


def test_named_col():
    assert "*" == str(all_cols())
    raises(NotImplementedError, lambda: all_cols().output_name)
    raises(NotImplementedError, lambda: all_cols().alias("x"))
    raises(NotImplementedError, lambda: all_cols().cast("long"))
    assert "a" == str(col("a"))
    assert "a" == str(col(col("a")))
    assert "ab AS xx" == str(col("ab").alias("xx"))
    assert "ab AS xx" == str(col("ab", "xx").cast(None))
    assert "CAST(ab AS long) AS xx" == str(col("ab", "xx").cast("long"))
    assert "CAST(ab AS long) AS xx" == str(col("ab").alias("xx").cast(int))
    raises(NotImplementedError, lambda: col([1, 2]))
    assert to_uuid(col("a")) != to_uuid(col("b"))
    assert to_uuid(col("a")) != to_uuid(col("a").alias("v"))
    assert to_uuid(col("a")) != to_uuid(col("a").cast(int))
    assert to_uuid(col("a").cast(int).alias("v")) == to_uuid(col("a").alias("v").cast(int))
    assert "" == col("a").infer_alias().as_name
    assert "a" == col("a").cast(int).infer_alias().as_name
    assert "CAST(a AS long) AS a" == str(col("a").cast(int).infer_alias())
    assert "CAST(a AS long) AS x" == str(col("a").cast(int).alias("x").infer_alias())


def test_lit_col():
    assert "NULL" == str(lit(None))
    assert "TRUE" == str(null().is_null())
    assert "FALSE" == str(null().not_null())
    assert "'a'" == str(lit("a"))
    assert "'a\"\\'\\\\'" == str(lit("a\"'\\"))
    assert "'a' AS x" == str(lit("a", "x"))
    assert "TRUE" == str(lit("a").not_null())
    assert "1.1" == str(lit(1.1))
    assert "11" == str(lit(11))
    assert "TRUE" == str(lit(True))
    assert "1 AS xx" == str(lit(1).alias("xx"))
    raises(NotImplementedError, lambda: lit([1, 2]))
    assert to_uuid(lit("a")) != to_uuid(col("a"))
    assert to_uuid(lit(1)) != to_uuid(lit("1"))
    assert to_uuid(null()) == to_uuid(null())
    assert to_uuid(lit("a").cast(int).alias("v")) == to_uuid(lit("a").alias("v").cast(int))


def test_operators():
    assert "-(a)" == str(-col("a"))
    assert "a" == (-col("a")).infer_alias().output_name
    assert "a" == str(+col("a"))
    assert "~(a)" == str(~col("a"))
    assert "IS_NULL(a)" == str(col("a").is_null())
    assert "NOT_NULL(a) AS a" == str(col("a").not_null().infer_alias())
    assert "+(ab,1)" == str(col("ab") + 1)
    assert "+('x',a)" == str("x" + col("a"))
    assert "-(1.1,a)" == str(1.1 - col("a"))
    assert "*(1.1,a)" == str(1.1 * col("a"))
    assert "/(1.1,a)" == str(1.1 / col("a"))
    assert "+(ab,1) AS xx" == str((col("ab") + 1).alias("xx"))
    assert "&(TRUE,a)" == str(True & col("a"))
    assert "|(a,FALSE)" == str(col("a") | False)
    assert ">(a,1.1)" == str(1.1 < col("a"))
    assert "<(1.1,a)" == str(lit(1.1) < col("a"))
    assert ">=(a,1.1)" == str(1.1 <= col("a"))
    assert "==(a,1.1)" == str(1.1 == col("a"))
    assert "!=(a,1)" == str(col("a") != 1)
    assert "-(+(a,*(10,b)),/(c,d))" == str((col("a") + 10 * col("b")) - col("c") / col("d"))
    assert "|(==(a,1.1),&(&(b,~(c)),TRUE))" == str((1.1 == col("a")) | col("b") & ~col("c") & True)
    expr = function("f", col("x") + col("z"), col("y"), 1, 1.1, False, "t")
    assert "f(+(x,z),y,1,1.1,FALSE,'t') AS x" == str(expr.alias("x"))
    expr = f.coalesce(col("x") + col("z"), col("y"), 1, 1.1, False, "t")
    assert "COALESCE(+(x,z),y,1,1.1,FALSE,'t')" == str(expr)
    expr = (col("a") + col("b")) * function("x", col("b"), a=col("c"), b=lit(1))
    assert {"a", "b", "c"} == set(_get_column_mentions(expr))


def test_schema_inference():
    schema = Schema("a:int,b:str,c:bool,d:double")
    assert pa.int32() == col("a").infer_type(schema)
    assert pa.int32() == (-col("a")).infer_type(schema)
    assert pa.int64() == (-col("a")).cast(int).infer_type(schema)
    assert pa.int64() == (-col("a").cast(int)).infer_type(schema)
    assert pa.string() == col("b").infer_type(schema)
    assert (-col("b")).infer_type(schema) is None
    assert (~col("b")).infer_type(schema) is None
    assert pa.bool_() == (~col("c")).alias("x").infer_type(schema)
    assert pa.float64() == (-col("d").alias("x")).infer_type(schema)
    assert col("x").infer_type(schema) is None
    assert pa.string() == col("x").cast(str).infer_type(schema)
    assert all_cols().infer_type(schema) is None
    assert pa.bool_() == (col("a") < col("d")).infer_type(schema)
    assert pa.bool_() == (~(col("a") != col("d"))).infer_type(schema)
    assert pa.int64() == (~(col("a") != col("d"))).cast(int).infer_type(schema)
    assert (col("a") - col("d")).infer_type(schema) is None
    assert pa.int64() == lit(1).infer_type(schema)
    assert pa.string() == lit("a").infer_type(schema)
    assert pa.bool_() == lit(False).infer_type(schema)
    assert pa.string() == lit(False).cast(str).infer_type(schema)
    assert pa.float64() == lit(2.2).infer_type(schema)
    assert null().infer_type(schema) is None
    assert pa.string() == null().cast(str).infer_type(schema)


def test_functions():
    schema = Schema("a:int,b:str,c:bool,d:double")
    assert f.is_agg(f.first(col("a")))
    assert f.is_agg(f.count_distinct(col("a")).alias("x"))
    assert f.is_agg(f.first(col("a")) + 1)
    assert f.is_agg((f.first(col("a")) < 1).alias("x"))
    assert f.is_agg(col("a") * f.first(col("a")) + 1)
    assert not f.is_agg(col("a") + col("b"))
    assert not f.is_agg(null())
    expr = f.coalesce(col("a"), 1, None, col("b") + col("c"))
    assert "COALESCE(a,1,NULL,+(b,c))" == str(expr)
    assert expr.infer_type(schema) is None
    expr = f.min(col("a"))
    assert "MIN(a)" == str(expr)
    assert pa.int32() == expr.infer_type(schema)
    assert "MIN(a) AS a" == str(expr.infer_alias())
    assert "CAST(MIN(a) AS long) AS a" == str(expr.cast(int).infer_alias())
    assert "MIN(a) AS b" == str(expr.alias("b").infer_alias())
    assert "MIN(-(a)) AS a" == str(f.min(-col("a")).infer_alias())
    assert pa.float64() == f.min(lit(1.1)).infer_type(schema)
    assert pa.int32() == f.max(col("a")).infer_type(schema)
    assert "FIRST(a)" == str(f.first(col("a")))
    assert pa.int32() == f.last(col("a")).infer_type(schema)
    assert f.avg(col("a")).infer_type(schema) is None
    assert f.sum(col("a")).infer_type(schema) is None
    assert "COUNT(a)" == str(f.count(col("a")))
    expr = f.count_distinct(col("a"))
    assert "COUNT(DISTINCT a)" == str(expr)
    assert "COUNT(DISTINCT a) AS a" == str(expr.infer_alias())
    expr = f.count_distinct(all_cols())
    assert "COUNT(DISTINCT *)" == str(expr)
    raises(NotImplementedError, lambda: expr.infer_alias())


def test_select_columns():
    cols = SelectColumns(col("a"), lit(1, "b"), col("bb") + col("cc"), f.first(col("c")))
    assert to_uuid(cols) == to_uuid(cols)
    raises(ValueError, lambda: cols.assert_all_with_names())
    cols2 = SelectColumns(col("a"), lit(1, "b"), col("bb") + col("cc"), f.first(col("c")), arg_distinct=True)
    assert to_uuid(cols) != to_uuid(cols2)
    cols = SelectColumns(col("a").alias("b"), lit(1, "b"))
    raises(ValueError, lambda: cols.assert_all_with_names())
    cols = SelectColumns(all_cols(), col("a")).assert_no_agg()
    raises(ValueError, lambda: cols.assert_all_with_names())
    raises(ValueError, lambda: SelectColumns(all_cols(), all_cols(), col("a").alias("p")))
    raises(ValueError, lambda: SelectColumns(all_cols(), f.first(col("a")).alias("x")))
    cols = SelectColumns(col("aa").alias("a").cast(int), lit(1, "b"), (col("bb") + col("cc")).alias("c"),
                         f.first(col("c")).alias("d")).assert_all_with_names()
    raises(AssertionError, lambda: cols.assert_no_agg())
    assert not cols.simple
    assert "CAST(aa AS long) AS a" == str(cols.simple_cols[0])
    assert cols.has_literals and "1 AS b" == str(cols.literals[0])
    assert cols.has_agg
    assert "+(bb,cc) AS c" == str(cols.non_agg_funcs[0])
    assert "FIRST(c) AS d" == str(cols.agg_funcs[0])
    assert 2 == len(cols.group_keys)
    assert "aa" == cols.group_keys[0].output_name
    assert "" == cols.group_keys[1].output_name
    assert isinstance(cols.group_keys[1], _BinaryOpExpr)
    cols = SelectColumns(col("a")).assert_no_wildcard()
    assert cols.simple and not cols.has_literals and not cols.has_agg
    cols = SelectColumns(col("x"), all_cols(), col("y") + col("z")).replace_wildcard(Schema("a:int,b:int"))
    assert ["x", "a", "b", "+(y,z)"] == [str(c) for c in cols.all_cols]


def test_sql_generator():
    gen = SQLExpressionGenerator()
    assert "a AS bc" == gen.generate(col("a").alias("bc"))
    assert "'a' AS bc" == gen.generate(lit("a").alias("bc"))
    assert "CAST(a AS long) AS a" == gen.generate(col("a").cast(int))
    assert "(a+2)*3" == gen.generate((col("a") + 2) * 3)
    assert "(-a+2)*3" == gen.generate((-col("a") + 2) * 3)
    assert "(a*2)/3 AS x" == gen.generate(((col("a") * 2) / 3).alias("x"))
    assert "COUNT(DISTINCT a) AS x" == gen.generate((f.count_distinct(col("a"))).alias("x"))
    assert "(a=-1) AND (b>=c)" == gen.generate((col("a") == -1) & (col("b") >= col("c")))
    assert "TRUE AND NOT (b>=c)" == gen.generate(True & ~(col("b") >= col("c")))
    assert "TRUE OR (b>=c) IS NOT NULL" == gen.generate(True | (col("b") >= col("c")).not_null())
    assert "COALESCE(a,b+c,(d+e)-1,NULL) IS NULL" == gen.generate(
        f.coalesce(col("a"), col("b") + col("c"), col("d") + col("e") - 1, null()).is_null())
    assert ("MY(MIN(`x y`),MAX(y+1),AVG(z),2,aa=FIRST(a),bb=LAST('b'),cc=COUNT(DISTINCT *)) AS `x z`"
            == gen.generate(function("MY", f.min(col("x y")), f.max(col("y") + 1), f.avg(col("z")), 2,
                                     aa=f.first(col("a")), bb=f.last(lit("b")),
                                     cc=f.count_distinct(all_cols())).alias("x z")))

    def sql(pieces):
        return " ".join(p[1] for p in pieces if p[1] != "").strip()

    assert "SELECT * FROM t WHERE (a>1) AND b IS NULL" == sql(gen.where((col("a") > 1) & col("b").is_null(), "t"))
    raises(ValueError, lambda: list(gen.where(f.max(col("a")) > 1, "t")))
    cols = SelectColumns(col("a"), f.max(col("b") + 1).alias("x"))
    assert "SELECT a, MAX(b+1) AS x FROM t WHERE b<2 GROUP BY a HAVING MAX(b+1)>0" == sql(
        gen.select(cols, "t", where=col("b") < 2, having=f.max(col("b") + 1) > 0))
    cols = SelectColumns(col("a"), lit(1, "o"), f.sum(col("b")).alias("c"))
    assert "SELECT a, 1 AS o, c FROM ( SELECT a, SUM(b) AS c FROM t GROUP BY a )" == sql(gen.select(cols, "t"))
    gen2 = SQLExpressionGenerator(enable_cast=False)
    cols = SelectColumns(col("a").cast(int), (col("b") + 1).alias("c").cast(str))
    assert "SELECT a AS a, b+1 AS c FROM t" == sql(gen2.select(cols, "t"))
    diff = gen2.correct_select_schema(Schema("a:int,b:int"), cols, Schema("a:int,c:long"))
    assert diff == Schema("a:long,c:str")
