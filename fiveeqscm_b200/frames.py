"""Tabular front-end ("DataFrame/CSV front-end"): emission scenarios in from CSV /
DataFrames, ensemble summaries out as DataFrames.  Pure host convenience around
:func:`fiveeqscm_b200.concentrations.run_ensemble`; nothing here touches the GPU.

Scenario table layout (long format), one row per (scenario, year):

    year, scenario, co2, ch4, n2o          # `scenario` optional (one scenario if absent)

Years must be equally spaced and identical across scenarios; units are the gas's own emission
units per year (GtC, MtCH4, MtN2O-N for the illustrative defaults in :mod:`.params`).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from . import stats as _stats
from .params import GASES


def scenarios_from_frame(df, gases: Sequence[str] = GASES, year_col: str = "year", scenario_col: str = "scenario"):
    """(years [n_t], names [S], emissions [G][n_t][S] float64, dt) from a long-format table."""
    missing = [c for c in (year_col, *gases) if c not in df.columns]
    if missing:
        raise ValueError(f"scenario table lacks column(s) {missing}")
    if scenario_col in df.columns:
        names = list(dict.fromkeys(df[scenario_col].tolist()))
        groups = [df[df[scenario_col] == n] for n in names]
    else:
        names, groups = ["scenario"], [df]
    years = None
    cols = []
    for name, g in zip(names, groups):
        g = g.sort_values(year_col)
        y = g[year_col].to_numpy(dtype=np.float64)
        if years is None:
            years = y
        elif y.shape != years.shape or not np.array_equal(y, years):
            raise ValueError(f"scenario {name!r} does not cover the same years as {names[0]!r}")
        cols.append(np.stack([g[gas].to_numpy(dtype=np.float64) for gas in gases]))      # [G][n_t]
    if years.size == 0:
        raise ValueError("scenario table is empty")
    dt = 1.0
    if years.size > 1:
        steps = np.diff(years)
        dt = float(steps[0])
        if dt <= 0 or not np.allclose(steps, dt, rtol=0, atol=1e-9 * max(1.0, abs(dt))):
            raise ValueError("years must be strictly increasing and equally spaced")
    E = np.ascontiguousarray(np.stack(cols, axis=2))                                          # [G][n_t][S]
    if not np.all(np.isfinite(E)):
        raise ValueError("scenario table contains non-finite emissions")
    return years, names, E, dt


def scenarios_from_csv(path, gases: Sequence[str] = GASES, **kw):
    """:func:`scenarios_from_frame` on ``pandas.read_csv(path)``."""
    import pandas as pd
    return scenarios_from_frame(pd.read_csv(path), gases, **kw)


def summary_frame(result, years, pcts: Sequence[float] = (5.0, 17.0, 50.0, 83.0, 95.0)):
    """Per-year ensemble summary of temperature from a result computed with ``stats=HistSpec(...)``
    (after the cross-GPU all-reduce, if any): mean, std, min, max and percentiles off the histogram."""
    import pandas as pd
    if result.hist is None or result.moments is None or result.spec is None:
        raise ValueError("summary_frame needs a result computed with stats=HistSpec(...)")
    mom = _stats._np(result.moments)
    hist = _stats._np(result.hist)
    n = hist.sum(axis=1)
    if not np.all(n == n[0]):
        raise ValueError("histogram rows count different numbers of members")
    mean, std = _stats.mean_std(mom, int(n[0]))
    p = _stats.percentiles(hist, result.spec.lo, result.spec.hi, pcts)
    data = {"year": np.asarray(years), "members": n, "T_mean": mean, "T_std": std, "T_min": mom[:, 2], "T_max": mom[:, 3]}
    for j, q in enumerate(pcts):
        data[f"T_p{q:g}"] = p[:, j]
    return pd.DataFrame(data)


def member_frame(result, years, member: int = 0, gases: Sequence[str] = GASES):
    """One member's trajectory as a table: C_<gas>, RF_<gas>, (E_<gas>,) T per year."""
    import pandas as pd
    data = {"year": np.asarray(years)}
    for field, prefix in (("C", "C_"), ("RF", "RF_"), ("E", "E_"), ("alpha", "alpha_")):
        arr = getattr(result, field, None)
        if arr is not None:
            a = _stats._np(arr[:, :, member])
            for g, gas in enumerate(gases):
                data[prefix + gas] = a[g]
    if result.T is not None:
        data["T"] = _stats._np(result.T[:, member])
    return pd.DataFrame(data)
