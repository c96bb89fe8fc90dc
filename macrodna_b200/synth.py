"""Synthetic scRNA / scDNA instances with planted clonal structure.

The reference ships no inputs for its own test (``test/rna_data.csv`` and
``test/dna_data.csv`` are missing, ``.MISSING_LARGE_BLOBS:14-15``), so every
config of BASELINE.json is synthesised at the reference's shapes with the value
distributions its preprocessing produces:

* DNA: ``log1p`` of noisy integer copy numbers (``BE_data_analysis/cna_filterer.py:32-40``);
* RNA: ``log1p(RPM(counts + 1))`` of Poisson counts whose rate follows the clone's
  copy number (``BE_data_analysis/rna_filterer.py:22-36``).

Host-side NumPy only; nothing here touches the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd

# name -> (M rna cells, N dna cells, G genes, clones)
CONFIG_SHAPES = {
    "C2": (192, 249, 2000, 2),
    "C3": (2000, 200, 10000, 4),
    "C4": (5000, 1000, 15000, 8),
    "C5": (50000, 10000, 20000, 16),
}


@dataclass
class Instance:
    rna: np.ndarray  # [M, G] float64, cells x genes (what ``rna_df.T.to_numpy()`` yields)
    dna: np.ndarray  # [N, G] float64
    rna_clone: np.ndarray  # [M] planted clone of each RNA cell
    dna_clone: np.ndarray  # [N]


def make_arrays(n_rna, n_dna, n_genes, n_clones, seed, constant_dna_cell=False, dtype=np.float64) -> Instance:
    """``constant_dna_cell=True`` plants one zero-variance DNA cell (an all-copy-number-2 profile, as the
    reference's CRC data contains): its correlations are exactly 0.0, i.e. an exact tie by construction --
    which RNA cell it takes in a step is arbitrary and cascades into the later steps' sub-problems."""
    rng = np.random.default_rng(seed)
    # piecewise-constant copy-number profiles over ~50-gene segments, mostly 2
    n_seg = max(1, -(-n_genes // 50))
    seg_cn = rng.choice([1, 2, 3, 4], size=(n_clones, n_seg), p=[0.12, 0.64, 0.16, 0.08])
    cn = np.repeat(seg_cn, 50, axis=1)[:, :n_genes].astype(np.float64)
    dna_clone = rng.integers(0, n_clones, size=n_dna)
    rna_clone = rng.integers(0, n_clones, size=n_rna)
    dna = np.empty((n_dna, n_genes), dtype=dtype)
    blk = max(1, (1 << 24) // max(1, n_genes))
    for s in range(0, n_dna, blk):
        e = min(n_dna, s + blk)
        noise = rng.standard_normal((e - s, n_genes))
        dna[s:e] = np.log1p(np.maximum(cn[dna_clone[s:e]] * (1.0 + 0.05 * noise), 0.0))
    if constant_dna_cell and n_dna > 1:
        dna[n_dna // 2] = np.log1p(2.0)  # all-copy-number-2 cell: zero variance
    base = rng.lognormal(0.0, 1.0, size=n_genes)
    lib = rng.lognormal(0.0, 0.3, size=n_rna)
    rna = np.empty((n_rna, n_genes), dtype=dtype)
    for s in range(0, n_rna, blk):
        e = min(n_rna, s + blk)
        lam = base[None, :] * (cn[rna_clone[s:e]] * 0.5) * lib[s:e, None]
        counts = rng.poisson(lam).astype(np.float64) + 1.0
        rpm = counts / counts.sum(axis=1, keepdims=True) * 1e6
        rna[s:e] = np.log1p(rpm)
    return Instance(rna=rna, dna=dna, rna_clone=rna_clone, dna_clone=dna_clone)


def make_config_arrays(name: str, seed: int | None = None, scale: float = 1.0, ties: bool = False) -> Instance:
    m, n, g, k = CONFIG_SHAPES[name]
    if scale != 1.0:
        m, n, g = max(2, int(m * scale)), max(2, int(n * scale)), max(8, int(g * scale))
    if seed is None:
        seed = 1234 + int(name[1:])
    return make_arrays(m, n, g, k, seed, constant_dna_cell=ties)


def make_frames(inst: Instance, extra_rna_genes: float = 0.03, seed: int = 0, rna_ids=None, dna_ids=None):
    """Wrap an instance as the genes x cells DataFrames the API takes.

    The RNA frame gets ``extra_rna_genes`` extra genes and a shuffled gene order
    so the gene-intersection path (``macrodna.py:89-91``) is exercised.
    """
    rng = np.random.default_rng(seed + 99)
    m, g = inst.rna.shape
    n = inst.dna.shape[0]
    genes = ["g%06d" % i for i in range(g)]
    rna_ids = rna_ids if rna_ids is not None else ["R%06d" % i for i in range(m)]
    dna_ids = dna_ids if dna_ids is not None else ["D%06d" % i for i in range(n)]
    dna_df = pd.DataFrame(inst.dna.T, index=genes, columns=dna_ids)
    n_extra = int(round(g * extra_rna_genes))
    extra = rng.random((n_extra, m)) * 3.0
    rna_vals = np.concatenate([inst.rna.T, extra], axis=0)
    rna_genes = genes + ["x%06d" % i for i in range(n_extra)]
    perm = rng.permutation(len(rna_genes))
    rna_df = pd.DataFrame(rna_vals[perm], index=[rna_genes[i] for i in perm], columns=rna_ids)
    label = pd.DataFrame({"clone": inst.dna_clone, "cell": dna_ids})
    return rna_df, dna_df, label


def resample_dna_columns(dna_clone: np.ndarray, seed: int) -> np.ndarray:
    """Replicate generator of the resampling-stability sweep.

    Follows ``Resampling_stability_analyses/CRC_data_analyses/clonal_proportions_resampling.py:174-187``:
    new clone sizes ~ Multinomial(n, Dirichlet(1)), then cells of each clone are
    drawn WITH replacement -> duplicate DNA columns (exact ties by construction).
    Returns the column indices of the replicate's DNA cells.
    """
    rng = np.random.default_rng(seed)
    clones = np.unique(dna_clone)
    n = dna_clone.shape[0]
    props = rng.multinomial(n, rng.dirichlet(np.ones(len(clones))))
    cols = []
    for k, cnt in zip(clones, props):
        members = np.flatnonzero(dna_clone == k)
        if cnt > 0:
            cols.append(rng.choice(members, size=cnt, replace=True))
    return np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
