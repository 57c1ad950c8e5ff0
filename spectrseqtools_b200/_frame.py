"""Tiny column-store frame with the sliver of the polars surface the mass-explanation path touches.

The reference shapes its alphabet with polars (masses.py:53-88, mass_table.py:154-204,
mass_explanation.py:17-42) and its unit test builds a one-column frame to sum nucleoside masses
(tests/test_explain_masses.py:16-31).  polars is not in this image, so `masses.py` builds
`EXPLANATION_MASSES` as a real `polars.DataFrame` when polars imports and as this `DataFrame`
otherwise.  `install_polars_shim()` registers a module named ``polars`` backed by this file ONLY when
the real one is missing; it never shadows a real install.

This is host-side glue for ~100-row tables: plain Python lists, no vectorisation on purpose.
"""
from __future__ import annotations

import sys
import types
from typing import Any, Callable, Dict, Iterable, List, Sequence


class Float64:  # dtype tags (only identity matters)
    pass


class Int64:
    pass


class Utf8:
    pass


class Series:
    def __init__(self, name: Any = None, values: Any = None):
        # polars allows Series(values) and Series(name, values) and Series(one_column_frame)
        if values is None and not isinstance(name, str):
            name, values = "", name
        if isinstance(values, DataFrame):
            if len(values.columns) != 1:
                raise ValueError("Series(DataFrame) needs exactly one column")
            name = values.columns[0]
            values = values._data[name]
        elif isinstance(values, Series):
            values = values._values
        self.name = name or ""
        self._values = list(values if values is not None else [])

    def to_list(self) -> list:
        return list(self._values)

    def __len__(self):
        return len(self._values)

    def __iter__(self):
        return iter(self._values)

    def __getitem__(self, i):
        return self._values[i]

    def sum(self):
        return sum(self._values)

    def max(self):
        return max(self._values)

    def min(self):
        return min(self._values)

    def alias(self, name: str) -> "Series":
        return Series(name, self._values)

    def __repr__(self):
        return f"Series({self.name!r}, {self._values!r})"


class Expr:
    """Deferred column expression: a function frame -> list, plus an output name."""

    def __init__(self, fn: Callable[["DataFrame"], list], name: str):
        self._fn = fn
        self._name = name

    def _eval(self, df: "DataFrame") -> list:
        return self._fn(df)

    def alias(self, name: str) -> "Expr":
        return Expr(self._fn, name)

    def _binary(self, other, op) -> "Expr":
        if isinstance(other, Expr):
            return Expr(lambda df: [op(a, b) for a, b in zip(self._fn(df), other._fn(df))], self._name)
        return Expr(lambda df: [op(a, other) for a in self._fn(df)], self._name)

    def __eq__(self, other):  # type: ignore[override]
        return self._binary(other, lambda a, b: a == b)

    def __ne__(self, other):  # type: ignore[override]
        return self._binary(other, lambda a, b: a != b)

    def __lt__(self, other):
        return self._binary(other, lambda a, b: a < b)

    def __le__(self, other):
        return self._binary(other, lambda a, b: a <= b)

    def __gt__(self, other):
        return self._binary(other, lambda a, b: a > b)

    def __ge__(self, other):
        return self._binary(other, lambda a, b: a >= b)

    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    def add(self, other):
        return self.__add__(other)

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def __mul__(self, other):
        return self._binary(other, lambda a, b: a * b)

    def __truediv__(self, other):
        return self._binary(other, lambda a, b: a / b)

    def __and__(self, other):
        return self._binary(other, lambda a, b: bool(a) and bool(b))

    def __or__(self, other):
        return self._binary(other, lambda a, b: bool(a) or bool(b))

    def __invert__(self):
        return Expr(lambda df: [not a for a in self._fn(df)], self._name)

    __hash__ = None  # type: ignore[assignment]

    def is_in(self, values: Iterable) -> "Expr":
        pool = list(values)
        return Expr(lambda df: [a in pool for a in self._fn(df)], self._name)

    def map_elements(self, fn: Callable, return_dtype: Any = None, **_kw) -> "Expr":
        return Expr(lambda df: [fn(a) for a in self._fn(df)], self._name)

    def round(self, decimals: int = 0) -> "Expr":
        return Expr(lambda df: [round(a, decimals) for a in self._fn(df)], self._name)

    def cast(self, dtype: Any) -> "Expr":
        conv = int if dtype is Int64 else float if dtype is Float64 else (lambda a: a)
        return Expr(lambda df: [conv(a) for a in self._fn(df)], self._name)

    def sum(self) -> "Expr":
        return Expr(lambda df: [sum(self._fn(df))], self._name)

    def max(self) -> "Expr":
        return Expr(lambda df: [max(self._fn(df))], self._name)

    def min(self) -> "Expr":
        return Expr(lambda df: [min(self._fn(df))], self._name)

    @property
    def str(self) -> "_StrNamespace":
        return _StrNamespace(self)


class _StrNamespace:
    """``pl.col(x).str``: only the literal-substring test the prediction path uses."""

    def __init__(self, expr: Expr):
        self._expr = expr

    def contains(self, pattern: str, literal: bool = False, **_kw) -> Expr:
        import re as _re

        rx = None if literal else _re.compile(pattern)
        inner = self._expr
        return Expr(lambda df: [(pattern in a) if rx is None else bool(rx.search(a)) for a in inner._fn(df)], inner._name)


def struct(*names: Any) -> Expr:
    """``pl.struct(a, b, ...)``: one dict per row (what ``map_elements`` callbacks index by column name)."""
    flat: List[str] = []
    for n in names:
        flat.extend(n if isinstance(n, (list, tuple)) else [n])
    return Expr(lambda df: [dict(zip(flat, row)) for row in zip(*(df._data[n] for n in flat))] if flat else [], flat[0] if flat else "struct")


def concat(frames: Iterable["DataFrame"], **_kw) -> "DataFrame":
    """Vertical concatenation of frames with the same columns."""
    frames = list(frames)
    if not frames:
        return DataFrame()
    cols = frames[0].columns
    return DataFrame({c: [x for f in frames for x in f._data[c]] for c in cols})


def col(name: str) -> Expr:
    return Expr(lambda df: list(df._data[name]), name)


def lit(value: Any, dtype: Any = None, **_kw) -> Expr:
    return Expr(lambda df: [value] * df.height, "literal")


class DataFrame:
    def __init__(self, data: Any = None, schema: Any = None, orient: Any = None, **_kw):
        self._data: Dict[str, list] = {}
        names = list(schema) if schema is not None else None
        if data is None:
            for n in names or []:
                self._data[n] = []
        elif isinstance(data, DataFrame):
            self._data = {k: list(v) for k, v in data._data.items()}
        elif isinstance(data, dict):
            for k, v in data.items():
                if isinstance(v, Series):
                    v = v.to_list()
                elif isinstance(v, (str, bytes)) or not hasattr(v, "__iter__"):
                    v = [v]  # polars broadcasts scalars to a one-row column
                self._data[k] = list(v)
            width = max((len(v) for v in self._data.values()), default=0)
            for k, v in self._data.items():
                if len(v) == 1 and width > 1:
                    self._data[k] = v * width
        else:
            seq = list(data)
            if seq and isinstance(seq[0], Series):
                for s in seq:
                    self._data[s.name] = s.to_list()
            elif seq and isinstance(seq[0], (list, tuple)) and orient != "col" and names and len(names) == len(seq[0]) and (orient == "row" or len(names) != len(seq)):
                for j, n in enumerate(names):
                    self._data[n] = [row[j] for row in seq]
            elif seq and isinstance(seq[0], (list, tuple)):
                names = names or [f"column_{j}" for j in range(len(seq))]
                for n, colvals in zip(names, seq):
                    self._data[n] = list(colvals)
            else:  # a flat sequence of scalars is ONE column (tests/test_explain_masses.py:17)
                names = names or ["column_0"]
                if len(names) != 1:
                    raise ValueError("flat data needs a one-name schema")
                self._data[names[0]] = seq
        lens = {len(v) for v in self._data.values()}
        if len(lens) > 1:
            raise ValueError(f"ragged columns: { {k: len(v) for k, v in self._data.items()} }")

    # ---- shape / introspection
    @property
    def columns(self) -> List[str]:
        return list(self._data.keys())

    @property
    def height(self) -> int:
        return len(next(iter(self._data.values()))) if self._data else 0

    @property
    def width(self) -> int:
        return len(self._data)

    @property
    def shape(self):
        return (self.height, self.width)

    def __len__(self):
        return self.height

    def is_empty(self) -> bool:
        return self.height == 0

    def get_column_index(self, name: str) -> int:
        return self.columns.index(name)

    def get_column(self, name: str) -> Series:
        return Series(name, self._data[name])

    def __getitem__(self, name: str) -> Series:
        return self.get_column(name)

    def to_dict(self, as_series: bool = False) -> Dict[str, list]:
        return {k: list(v) for k, v in self._data.items()}

    def rows(self) -> List[tuple]:
        return list(zip(*self._data.values())) if self._data else []

    def iter_rows(self, named: bool = False):
        if named:
            for row in self.rows():
                yield dict(zip(self.columns, row))
        else:
            yield from self.rows()

    def item(self, row: int | None = None, column: Any = None):
        if row is None and column is None:
            if self.shape != (1, 1):
                raise ValueError(f"item() needs a 1x1 frame, got {self.shape}")
            return next(iter(self._data.values()))[0]
        name = column if isinstance(column, str) else self.columns[column]
        return self._data[name][row]

    # ---- verbs
    def _resolve(self, exprs: Sequence[Any]) -> List[tuple]:
        flat: List[Any] = []
        for e in exprs:
            if isinstance(e, (list, tuple)):
                flat.extend(e)
            else:
                flat.append(e)
        out = []
        for e in flat:
            if isinstance(e, str):
                out.append((e, list(self._data[e])))
            elif isinstance(e, Series):
                out.append((e.name, e.to_list()))
            else:
                out.append((e._name, e._eval(self)))
        return out

    def select(self, *exprs: Any) -> "DataFrame":
        return DataFrame({n: v for n, v in self._resolve(exprs)})

    def with_columns(self, *exprs: Any, **named: Any) -> "DataFrame":
        data = {k: list(v) for k, v in self._data.items()}
        for n, v in self._resolve(exprs):
            data[n] = v
        for n, e in named.items():
            data[n] = e._eval(self) if isinstance(e, Expr) else list(e)
        return DataFrame(data)

    def filter(self, *preds: Any) -> "DataFrame":
        keep = [True] * self.height
        for p in preds:
            vals = p._eval(self) if isinstance(p, Expr) else list(p)
            keep = [a and bool(b) for a, b in zip(keep, vals)]
        return DataFrame({k: [x for x, f in zip(v, keep) if f] for k, v in self._data.items()})

    def sort(self, by: Any, descending: bool = False) -> "DataFrame":
        keys = [by] if isinstance(by, (str, Expr)) else list(by)
        cols = [self._data[k] if isinstance(k, str) else k._eval(self) for k in keys]
        order = sorted(range(self.height), key=lambda i: tuple(c[i] for c in cols), reverse=descending)  # stable, like polars
        return DataFrame({k: [v[i] for i in order] for k, v in self._data.items()})

    def drop(self, *names: Any) -> "DataFrame":
        gone = set()
        for n in names:
            gone.update(n if isinstance(n, (list, tuple)) else [n])
        return DataFrame({k: list(v) for k, v in self._data.items() if k not in gone})

    def rename(self, mapping: Dict[str, str]) -> "DataFrame":
        return DataFrame({mapping.get(k, k): list(v) for k, v in self._data.items()})

    def with_row_index(self, name: str = "index", offset: int = 0) -> "DataFrame":
        data = {name: list(range(offset, offset + self.height))}
        data.update({k: list(v) for k, v in self._data.items()})
        return DataFrame(data)

    def write_csv(self, file: Any = None, separator: str = ",", **_kw):
        def cell(x):
            if x is None:
                return ""
            if isinstance(x, bool):
                return "true" if x else "false"
            return str(x)

        lines = [separator.join(self.columns)] + [separator.join(cell(x) for x in row) for row in self.rows()]
        text = "\n".join(lines) + "\n"
        if file is None:
            return text
        with open(file, "w") as fh:
            fh.write(text)
        return None

    def sum(self) -> "DataFrame":
        return DataFrame({k: [sum(v)] for k, v in self._data.items()})

    def replace_column(self, index: int, column: Series) -> "DataFrame":
        names = self.columns
        data = {}
        for j, n in enumerate(names):
            if j == index:
                data[column.name] = column.to_list()
            else:
                data[n] = list(self._data[n])
        return DataFrame(data)

    def join(self, other: "DataFrame", on: str, how: str = "inner") -> "DataFrame":
        if how not in ("left", "inner"):
            raise NotImplementedError(how)
        right_cols = [c for c in other.columns if c != on]
        out: Dict[str, list] = {c: [] for c in self.columns + right_cols}
        for i in range(self.height):
            key = self._data[on][i]
            hits = [j for j in range(other.height) if other._data[on][j] == key]
            if not hits and how == "left":
                for c in self.columns:
                    out[c].append(self._data[c][i])
                for c in right_cols:
                    out[c].append(None)
            for j in hits:
                for c in self.columns:
                    out[c].append(self._data[c][i])
                for c in right_cols:
                    out[c].append(other._data[c][j])
        return DataFrame(out)

    def __repr__(self):
        head = " | ".join(self.columns)
        body = "\n".join(" | ".join(str(x) for x in row) for row in self.rows())
        return f"shape: {self.shape}\n{head}\n{body}"


def read_csv(source: Any, separator: str = ",", **_kw) -> DataFrame:
    """Typed read of a small delimited file: int, then float, then true/false, else string (per column)."""
    with open(source) as fh:
        rows = [line.rstrip("\n").split(separator) for line in fh if line.strip()]
    header, body = rows[0], rows[1:]

    def convert(vals: List[str]) -> list:
        for conv in (int, float):
            try:
                return [conv(v) for v in vals]
            except ValueError:
                pass
        if all(v in ("true", "false") for v in vals):
            return [v == "true" for v in vals]
        return vals

    return DataFrame({h: convert([r[j] for r in body]) for j, h in enumerate(header)})


def install_polars_shim() -> bool:
    """Make ``import polars`` work for the reference's unit test when polars is absent.

    Returns True if the shim was installed, False if a real polars is importable (left alone).
    """
    if "polars" in sys.modules:
        return getattr(sys.modules["polars"], "__spectrseq_shim__", False)
    try:
        import polars  # noqa: F401

        return False
    except ImportError:
        pass
    mod = types.ModuleType("polars")
    mod.__spectrseq_shim__ = True
    mod.__doc__ = "spectrseqtools_b200 minimal stand-in for polars (real polars not installed)"
    for name, obj in dict(DataFrame=DataFrame, Series=Series, Expr=Expr, col=col, lit=lit, struct=struct, concat=concat,
                          read_csv=read_csv, Float64=Float64, Int64=Int64, Utf8=Utf8, String=Utf8, Boolean=bool).items():
        setattr(mod, name, obj)
    sys.modules["polars"] = mod
    return True
