"""A catalogue of column expressions / SELECT statements written against the API that the reference's
``fugue.column`` and ``fugue_b200.column`` share.  ``describe_all(ns)`` records what the DSL in ``ns`` says
about each of them; tests/golden/make_column_golden.py runs it on the reference's code, the test on ours."""
from typing import Any, Dict, List


def _expressions(ns: Any) -> Dict[str, Any]:
    col, lit, null, all_cols, function, f = ns.col, ns.lit, ns.null, ns.all_cols, ns.function, ns.f
    a, b, c, d = col("a"), col("b"), col("c"), col("d")
    return {
        "named": a, "named_alias": a.alias("x"), "named_cast": a.cast("double"), "named_cast_alias": col("a", "x").cast(int),
        "odd_name": col("x y"), "odd_alias": a.alias("a b"),
        "lit_int": lit(1), "lit_float": lit(1.5), "lit_str": lit("it's"), "lit_bool": lit(True), "lit_null": null(),
        "lit_alias": lit(1, "one"), "lit_cast": lit(1).cast(str).alias("s"), "null_cast": null().cast("double").alias("n"),
        "neg": -a, "neg_alias": (-a).alias("na"), "pos": +a, "not": ~c, "is_null": a.is_null(), "not_null": (a + b).not_null(),
        "add": a + b, "radd": 1 + a, "sub": a - 1, "rsub": 1.5 - a, "mul": a * b, "rmul": 2 * a, "div": a / b, "rdiv": 1 / a,
        "lt": a < b, "le": a <= 1, "gt": a > 1.5, "ge": 1 >= a, "eq": a == b, "ne": "x" != col("s"), "rlt": 1 < a,
        "and": (a > 1) & c, "rand": True & c, "or": c | (b < 2), "ror": False | c,
        "nested": ((a + b) * (a - b) / 2 - 1).alias("n"), "logic": ((a < b) & ~(b > 1)) | a.is_null(),
        "cast_inner": (a.cast(float) + b).alias("ci"), "cast_outer": (a + b).cast("int").alias("co"),
        "neg_cast": (-d).cast(int), "not_cmp": ~(a != d),
        "coalesce": f.coalesce(a, b + 1, 0, None), "func": function("my_f", a, 1, "s", False, x=b, y=2),
        "min": f.min(a), "max_expr": f.max(a + 1), "sum": f.sum(a), "avg": f.avg(d), "count": f.count(a),
        "count_star": f.count(all_cols()).alias("n"), "count_distinct": f.count_distinct(b), "first": f.first(a),
        "last_neg": f.last(-a), "agg_alias": f.sum(a).alias("s"), "agg_cast": f.max(a).cast(float),
        "agg_arith": (f.max(a) * 2 - f.min(b)).alias("r"), "agg_cmp": f.sum(a) >= 7, "agg_in_func": f.coalesce(f.max(a), 0),
        "min_lit": f.min(lit(1.1)), "max_cast_arg": f.max(a.cast("long")),
    }


def _selects(ns: Any) -> Dict[str, Any]:
    col, lit, all_cols, f, S = ns.col, ns.lit, ns.all_cols, ns.f, ns.SelectColumns
    a, b = col("a"), col("b")
    return {
        "plain": (S(a, b.alias("bb")), None, None),
        "star_where": (S(all_cols()), (a > 1) & b.is_null(), None),
        "exprs": (S(a, (b + 1).alias("c").cast(str), lit(1, "one")), a + b == 3, None),
        "distinct": (S(a, (b * 2).alias("d"), arg_distinct=True), None, None),
        "agg": (S(a, f.sum(b).cast(float).alias("s")), None, None),
        "agg_having": (S(a, f.max(b + 1).alias("x")), (b < 2) & (a > 1), f.max(b + 1) > 0),
        "agg_no_keys": (S(f.count(all_cols()).alias("n"), f.avg(b).alias("m")), None, None),
        "agg_literal": (S(a, lit(1, "o").cast(str), f.sum(b).alias("c")), None, (f.sum(b) >= 7) | (a == 1)),
        "agg_expr_key": (S((a + 1).alias("k"), f.min(b).alias("lo")), None, None),
    }


def _describe_expr(ns: Any, e: Any) -> Dict[str, Any]:
    schema = ns.Schema("a:int,b:long,c:bool,d:double,s:str")
    gen, gen_nc = ns.SQLExpressionGenerator(), ns.SQLExpressionGenerator(enable_cast=False)
    out: Dict[str, Any] = {"str": str(e), "is_agg": bool(ns.f.is_agg(e))}
    for key, fn in (("sql", lambda: gen.generate(e)), ("sql_nocast", lambda: gen_nc.generate(e)),
                    ("output_name", lambda: e.output_name),
                    ("inferred_alias", lambda: e.infer_alias().output_name),
                    ("inferred_type", lambda: None if e.infer_type(schema) is None else str(e.infer_type(schema)))):
        try:
            out[key] = fn()
        except Exception as ex:  # the kind of error is part of the behaviour
            out[key] = "!" + type(ex).__name__
    return out


def _describe_select(ns: Any, item: Any) -> Dict[str, Any]:
    cols, where, having = item
    gen = ns.SQLExpressionGenerator(enable_cast=False)
    out: Dict[str, Any] = {
        "str": str(cols), "has_agg": cols.has_agg, "has_literals": cols.has_literals, "simple": cols.simple,
        "group_keys": [str(k) for k in cols.group_keys], "agg_funcs": [str(k) for k in cols.agg_funcs],
        "non_agg_funcs": [str(k) for k in cols.non_agg_funcs], "literals": [str(k) for k in cols.literals],
    }
    try:
        out["sql"] = " ".join(t for _, t in gen.select(cols, "t", where=where, having=having) if t != "")
    except Exception as ex:
        out["sql"] = "!" + type(ex).__name__
    return out


def describe_all(ns: Any) -> Dict[str, Any]:
    return {"expressions": {k: _describe_expr(ns, e) for k, e in _expressions(ns).items()},
            "selects": {k: _describe_select(ns, s) for k, s in _selects(ns).items()}}
