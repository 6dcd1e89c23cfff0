"""Host file IO (``fugue_b200.io``): the behaviours the reference pins for ``fugue/_utils/io.py`` in
tests/fugue/utils/test_io.py:14-262 (path analysis, parquet / csv / json round trips, folders, patterns,
save modes), on local paths.  CPU only - the engine adds one ``to_df`` (H2D) behind ``load_df`` and one
``as_local`` (D2H) in front of ``save_df``."""
import gzip
import os

import pytest

from fugue_b200.dataframe import ArrayDataFrame, PandasDataFrame, df_eq
from fugue_b200.io import FilePath, load_df, save_df

ROW = [["1", 2, 3]]
SCHEMA = "a:str,b:int,c:long"


@pytest.mark.parametrize("name,hint,suffix,fmt", [
    ("/a/b/c.parquet", None, ".parquet", "parquet"), ("c.csv", None, ".csv", "csv"),
    ("/a/b/c.csv.gz", None, ".csv.gz", "csv"), ("/a/b/c.json", None, ".json", "json"),
    ("/a/b/c.json.gz", None, ".json.gz", "json"), ("/a/b/c.test.parquet", None, ".test.parquet", "parquet"),
    ("/a/b/c.ppp.gz", "csv", ".ppp.gz", "csv"), ("/a/b/c", "csv", "", "csv"),
    ("/a/b/*.parquet", None, ".parquet", "parquet"), ("/a/b/*123.parquet", None, ".parquet", "parquet"),
])
def test_path_analysis(name, hint, suffix, fmt):
    fp = FilePath(name, hint)
    assert (fp.suffix, fp.file_format, fp.has_glob, fp.raw_path) == (suffix, fmt, "*" in name, name)
    assert fp.path == os.path.abspath(name)


def test_unknown_formats():
    for name, hint in [("/a/b/c.ppp", None), ("/a/b/c.parquet", "csvv"), ("/a/b/c", None)]:
        with pytest.raises(NotImplementedError):
            FilePath(name, hint)
    with pytest.raises(NotImplementedError):
        FilePath("s3://bucket/x.parquet")


def test_parquet(tmp_path):
    path = str(tmp_path / "deep" / "er" / "a.parquet")                     # parents are created
    for df in (PandasDataFrame(ROW, SCHEMA), ArrayDataFrame([[[1, 2]]], "a:[int]"),
               ArrayDataFrame([[dict(a=1)]], "a:{a:long}")):
        save_df(df, path)
        df_eq(df, load_df(path), throw=True, check_order=True)
    save_df(PandasDataFrame(ROW, SCHEMA), path)
    df_eq(load_df(path, columns=["b", "a"]), [[2, "1"]], "b:int,a:str", throw=True)
    df_eq(load_df(path, columns="b:str,a:int"), [["2", 1]], "b:str,a:int", throw=True)
    with pytest.raises(Exception):
        load_df(path, columns="bb:str,a:int")


def test_parquet_folders_patterns_and_modes(tmp_path):
    df = PandasDataFrame(ROW, SCHEMA)
    twice = [ROW[0], ROW[0]]
    for name in ("folder.parquet", "folder"):
        folder = tmp_path / name
        folder.mkdir()
        (folder / "_SUCCESS").touch()
        save_df(df, str(folder / "1.parquet"))
        save_df(df, str(folder / "3.parquet"))
    f1, f2 = str(tmp_path / "folder" / "1.parquet"), str(tmp_path / "folder" / "3.parquet")
    for source, hint in [(str(tmp_path / "folder"), "parquet"), ([f1, f2], "parquet"),
                         (str(tmp_path / "folder.parquet"), None), (str(tmp_path / "folder" / "*.parquet"), None)]:
        df_eq(load_df(source, hint), twice, SCHEMA, throw=True)
    # a folder is replaced by a single file
    target = str(tmp_path / "folder.parquet")
    save_df(load_df(target), target, mode="overwrite")
    assert os.path.isfile(target)
    df_eq(load_df(target), twice, SCHEMA, throw=True)
    for existing in (f1, target):
        with pytest.raises(FileExistsError):
            save_df(df, existing, mode="error")
    with pytest.raises(NotImplementedError):
        save_df(df, f1, mode="dummy")
    with pytest.raises(FileNotFoundError):
        load_df(str(tmp_path / "nothing" / "*.parquet"))


def test_csv(tmp_path):
    df = PandasDataFrame(ROW, SCHEMA)
    path = str(tmp_path / "a.csv")
    save_df(df, path)                                                        # no header by default
    assert open(path).read().startswith("1,2,3")
    with pytest.raises(ValueError):
        load_df(path, header=False)                                          # names have to come from somewhere
    got = load_df(path, columns=["a", "b", "c"], header=False, infer_schema=True)
    assert got.as_array() == [[1, 2, 3]] and got.schema == "a:long,b:long,c:long"
    got = load_df(path, columns="a:double,b:str,c:str", header=False)
    assert got.as_array() == [[1.0, "2", "3"]] and got.schema == "a:double,b:str,c:str"
    save_df(df, path, header=True)
    assert open(path).read().startswith("a,b,c") or open(path).read().startswith('"a","b","c"')
    assert load_df(path, header=True).as_array() == [["1", "2", "3"]]        # text unless asked to infer
    assert load_df(path, header=True, infer_schema=True).as_array() == [[1, 2, 3]]
    assert load_df(path, columns=["b", "a"], header=True, infer_schema=True).as_array() == [[2, 1]]
    assert load_df(path, columns="b:str,a:double", header=True).as_array() == [["2", 1.0]]
    with pytest.raises(KeyError):
        load_df(path, columns="b:str,x:double", header=True)
    with pytest.raises(NotImplementedError):
        load_df(path, columns="b:str,x:double", header=2)
    with pytest.raises(ValueError):
        load_df(path, columns="b:str,a:double", header=True, infer_schema=True)


def test_csv_keeps_text_as_written_and_reads_folders(tmp_path):
    rows = [["007", None, "x,y"], ["1e3", "", "q\"uote"]]
    path = str(tmp_path / "t.csv")
    save_df(ArrayDataFrame(rows, "a:str,b:str,c:str"), path, header=True)
    got = load_df(path, header=True).as_array()
    assert got[0] == ["007", None, "x,y"] and got[1][0] == "1e3" and got[1][2] == "q\"uote"   # no number parsing
    folder = tmp_path / "parts"
    folder.mkdir()
    for i in range(3):
        save_df(ArrayDataFrame([[i, i * 1.5]], "k:long,v:double"), str(folder / f"{i}.csv"))
    got = load_df(str(folder), "csv", columns="k:long,v:double", header=False)
    assert sorted(got.as_array()) == [[0, 0.0], [1, 1.5], [2, 3.0]]
    gz = str(tmp_path / "z.csv.gz")
    save_df(ArrayDataFrame([[5, 6]], "a:long,b:long"), gz, header=True)
    assert gzip.open(gz, "rt").read().splitlines()[1] == "5,6"
    assert load_df(gz, header=True, infer_schema=True).as_array() == [[5, 6]]


def test_json(tmp_path):
    path = str(tmp_path / "a.json")
    save_df(PandasDataFrame(ROW, SCHEMA), path)
    assert open(path).read().strip() == '{"a": "1", "b": 2, "c": 3}'          # one record per line
    df_eq(load_df(path), [["1", 2, 3]], "a:str,b:long,c:long", throw=True)
    df_eq(load_df(path, columns=["b", "a"]), [[2, "1"]], "b:long,a:str", throw=True)
    df_eq(load_df(path, columns="b:str,a:int"), [["2", 1]], "b:str,a:int", throw=True)
    with pytest.raises(KeyError):
        load_df(path, columns="bb:str,a:int")
    gz = str(tmp_path / "b.json.gz")
    save_df(ArrayDataFrame([[1, None], [2, "x"]], "k:long,s:str"), gz)
    df_eq(load_df(gz), [[1, None], [2, "x"]], "k:long,s:str", throw=True)
