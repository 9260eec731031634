"""Training-data files on either side of the path.

Writer side: the reference's ``train`` word appends one line per flow per report to
``<TypeOfData>_training_data.csv`` (reference ``traffic_classifier.py:121-142``, header ``:217``):
17 tab-separated fields -- 16 numbers and the traffic type.  ``flows.FlowTable.training_lines`` produces those
lines; ``write_training_file`` wraps it.

Reader side: the notebooks' recipe (every ``*.ipynb`` cells 2-4): four of the bundled files are
tab-separated, ``game`` is comma-separated; the files are concatenated in the order given; rows with a missing
field are dropped (``dropna()`` -- it removes exactly one row of the bundled data, the truncated last line of
``ping_training_data.csv``); the four cumulative counters ``Forward/Reverse Packets/Bytes`` are dropped and the
remaining 12 columns, in file order, are the features in the order of ``traffic_classifier.py:104``.
No pandas here: the format is 17 fields per line.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

from . import flows as _flows

COLUMNS: Tuple[str, ...] = tuple(_flows.TRAINING_HEADER.rstrip("\n").split("\t"))
LABEL_COLUMN = "Traffic Type"
DROPPED = ("Forward Packets", "Forward Bytes", "Reverse Packets", "Reverse Bytes")
FEATURE_NAMES: Tuple[str, ...] = tuple(c for c in COLUMNS if c != LABEL_COLUMN and c not in DROPPED)
# the notebooks read the files in this order (quake is referenced there but not bundled)
NOTEBOOK_ORDER = ("ping", "voice", "dns", "telnet", "game", "quake")


_POW10 = [float("1e%d" % k) for k in range(309)]   # correctly rounded, like the C table of literals


def parse_float(text: str) -> float:
    """Decimal text -> float64 the way the notebooks' ``pd.read_csv`` does it by default (pandas' C tokenizer,
    ``precise_xstrtod``, the default since pandas 1.2): accumulate up to 17 significant digits in a double, then ONE
    multiplication or division by a correctly rounded power of ten.  Exact for short numbers, but NOT correctly rounded
    for the 17-digit reprs the capture writes (the 17th digit no longer fits 2^53): 2 455 values of the bundled rows sit
    one ulp away from ``float(text)``.  The golden rows (and any model fitted through pandas) hold exactly these
    values, so the reader reproduces them.
    Raises ValueError for anything that is not a plain decimal number (inf/nan spellings included: such a row is
    dropped like a missing field)."""
    t = text.strip()
    i, n = 0, len(t)
    neg = False
    if i < n and t[i] in "+-":
        neg = t[i] == "-"
        i += 1
    number, num_digits, num_decimals, exponent = 0.0, 0, 0, 0
    seen = False
    while i < n and t[i].isdigit():
        seen = True
        if num_digits < 17:
            number = number * 10.0 + (ord(t[i]) - 48)
            num_digits += 1
        else:
            exponent += 1
        i += 1
    if i < n and t[i] == ".":
        i += 1
        while i < n and t[i].isdigit():
            seen = True
            if num_digits < 17:
                number = number * 10.0 + (ord(t[i]) - 48)
                num_digits += 1
                num_decimals += 1
            i += 1
        exponent -= num_decimals
    if not seen:
        raise ValueError(f"not a number: {text!r}")
    if neg:
        number = -number
    if i < n and t[i] in "eE":
        i += 1
        eneg = False
        if i < n and t[i] in "+-":
            eneg = t[i] == "-"
            i += 1
        if i >= n or not t[i].isdigit():
            raise ValueError(f"not a number: {text!r}")
        e = 0
        while i < n and t[i].isdigit():
            e = e * 10 + (ord(t[i]) - 48)
            i += 1
        exponent += -e if eneg else e
    if i != n:
        raise ValueError(f"not a number: {text!r}")
    if exponent > 308:
        return float("-inf") if neg else float("inf")
    if exponent > 0:
        return number * _POW10[exponent]
    if exponent < -308:
        if exponent < -616:
            return -0.0 if neg else 0.0
        return number / _POW10[-308 - exponent] / _POW10[308]
    return number / _POW10[-exponent]


def _split(line: str, delim: str) -> List[str]:
    return line.rstrip("\r\n").split(delim)


def read_training_file(path: str) -> Tuple[np.ndarray, np.ndarray]:
    """One training file -> (features [n, 12] float64, labels [n] str).

    The delimiter is whatever separates the header's fields (tab or comma).  A row is dropped when it has fewer
    fields than the header, an empty field or a field that is not a number -- what ``read_csv(...).dropna()`` drops."""
    with open(path, "r", newline="") as f:
        header = f.readline()
        if not header:
            raise ValueError(f"{path}: empty file")
        delim = "\t" if "\t" in header else ","
        names = [h.strip() for h in _split(header, delim)]
        if LABEL_COLUMN not in names:
            raise ValueError(f"{path}: no '{LABEL_COLUMN}' column")
        missing = [c for c in FEATURE_NAMES if c not in names]
        if missing:
            raise ValueError(f"{path}: missing columns {missing}")
        fcols = [names.index(c) for c in FEATURE_NAMES]
        ncols = [i for i, c in enumerate(names) if c != LABEL_COLUMN]   # every numeric column must parse (dropna is row-wide)
        lcol = names.index(LABEL_COLUMN)
        feats, labels = [], []
        for line in f:
            if not line.strip():
                continue
            parts = _split(line, delim)
            if len(parts) < len(names):
                continue                       # truncated line -> NaN fields -> dropped
            try:
                vals = [parse_float(parts[i]) for i in ncols]
            except ValueError:
                continue
            lab = parts[lcol].strip()
            if not lab or any(v != v for v in vals):
                continue
            row = dict(zip(ncols, vals))
            feats.append([row[i] for i in fcols])
            labels.append(lab)
    X = np.asarray(feats, dtype=np.float64).reshape(len(feats), len(FEATURE_NAMES))
    return X, np.asarray(labels, dtype=str)


def load_training_set(paths: Sequence[str]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate training files in the given order (the notebooks' ``pd.concat``)."""
    if not paths:
        raise ValueError("no training files")
    parts = [read_training_file(p) for p in paths]
    return np.concatenate([p[0] for p in parts], axis=0), np.concatenate([p[1] for p in parts], axis=0)


def write_training_file(path: str, tables: Iterable["_flows.FlowTable"], traffic_type: str) -> int:
    """Header + one line per flow per table snapshot, as the reference's ``train`` capture writes them.
    Returns the number of data lines written."""
    n = 0
    with open(path, "w") as f:
        f.write(_flows.TRAINING_HEADER)
        for t in tables:
            for ln in t.training_lines(traffic_type):
                f.write(ln)
                n += 1
    return n
