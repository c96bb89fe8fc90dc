"""Wire formats and normalisers either side of the hot path.  Host-side pandas.

* ``write_indexed_csv``: the ``<biopsy>_cell2cell_assignment_indexed.csv`` the analysis scripts exchange
  (written at random_assignment_test.py:305 / run_loo_experiment.py:330 with ``DataFrame.to_csv``; read back by
  BE_data_analysis/aggregate_macrodna.py:24-41 with ``index_col=0``) -- header ``cell,predict_cell,step``.
* ``normalize_dna_counts``: BE_data_analysis/cna_filterer.py:30-40 (coverage filter, pseudocount, median-ratio
  copy number x 2, log1p).
* ``normalize_rna_counts``: BE_data_analysis/rna_filterer.py:20-36 (coverage filter, gene filter, pseudocount,
  RPM, log1p).
"""
from __future__ import annotations

import numpy as np
import pandas as pd


def write_indexed_csv(tagged: pd.DataFrame, path) -> None:
    """``tagged``: the second frame of ``cell2cell_assignment()`` (index ``cell``; columns ``predict_cell``, ``step``)."""
    if tagged.index.name is None or list(tagged.columns)[:1] not in (["predict_cell"], ["predicted_dna_cell"]):
        raise ValueError("expected the tagged assignment frame (index = RNA cell, first column = predicted DNA cell)")
    tagged.to_csv(path)


def read_indexed_csv(path) -> pd.DataFrame:
    """aggregate_macrodna.py:24: ``pd.read_csv(path, index_col=0)``."""
    return pd.read_csv(path, index_col=0)


def normalize_dna_counts(counts: pd.DataFrame, min_coverage: float = 3000) -> pd.DataFrame:
    """genes/bins x cells read counts -> log1p(2 * (count + 1) / median over bins), cells with total count
    <= min_coverage dropped (cna_filterer.py:27-40)."""
    df = counts.fillna(0)                                   # :27-29
    df = df[df.columns[df.sum() > min_coverage]]            # :31
    df = df + 1                                             # :34
    df = df.div(df.median())                                # :36
    df = df.mul(2)                                          # :37
    return np.log1p(df)                                     # :39


def normalize_rna_counts(counts: pd.DataFrame, min_coverage: float = 3000, min_transcripts: float = 3) -> pd.DataFrame:
    """genes x cells transcript counts -> log1p(RPM(count + 1)); cells with total <= min_coverage and genes never
    reaching min_transcripts in any cell dropped; the ``__chr*`` suffix of gene ids removed (rna_filterer.py:17-42)."""
    df = counts.fillna(0)                                   # :17-19
    df = df[df.columns[df.sum() > min_coverage]]            # :21
    df = df.T
    df = df[df.columns[(df >= min_transcripts).any()]]      # :26
    df = df.T
    df = df + 1                                             # :30
    df = df.div(df.sum())                                   # :32
    df = df.mul(1e6)                                        # :33
    df = np.log1p(df)                                       # :35
    return df.rename(index={g: str(g).split("__chr")[0] for g in df.index})  # :38-41
