"""CPU restatement of MaCroDNA's cell-matching hot path.  TEST INFRASTRUCTURE ONLY.

This module is the parity oracle.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it;
the product package ``macrodna_b200`` never does.

It follows ``/root/reference/src/MaCroDNA/macrodna.py`` function by function
(citations are ``macrodna.py:line``) in vectorised float64 NumPy, with
``scipy.optimize.linear_sum_assignment(maximize=True)`` standing in for the
reference's Gurobi ILP (``macrodna.py:27-84``).  The ILP is a rectangular
assignment problem -- row sums <= 1 (``:45-47``), column sums <= 1 (``:49-51``),
exactly ``min(|R|, N)`` ones (``:29,53``), maximise sum c_ij x_ij (``:60-67``) --
whose LP relaxation is integral, so the exact LAP optimum equals the ILP optimum.

Parity status: PINNED on the one known answer the reference publishes for this
path (``README.md:138``: ``Best objective 2.830007718086e+00`` for
``tiny_test``), and cross-checked against the reference file itself run
byte-unmodified through ``oracle/run_reference.py`` (scipy-backed ``gurobipy``
stand-in) on small seeded cases -- fixtures in ``tests/golden``.  Beyond
``tiny_test`` the Gurobi boundary itself is unpinned (no Gurobi here, no other
published values); the scipy-backed reference run is the operative reference.
"""
from __future__ import annotations

import numpy as np
import pandas as pd
from scipy.optimize import linear_sum_assignment

EPS = 1e-10  # macrodna.py:25


def shared_genes(rna_df: pd.DataFrame, dna_df: pd.DataFrame) -> list:
    """Gene intersection, ``macrodna.py:89``.

    The reference uses a Python ``set`` (hash order, unreproducible).  Only the
    summation order over genes depends on it, so the restatement fixes the
    canonical order "DNA-frame order" (SURVEY.md section 7.3 item 8).
    """
    rna_genes = set(rna_df.index.to_list())
    return [g for g in dna_df.index.to_list() if g in rna_genes]


def standardise(x: np.ndarray):
    """Per-cell centring and 2-norm, the per-operand half of ``macrodna.py:25``.

    ``x`` is cells x genes.  Returns (centred rows, norms).  The mean is NumPy's
    pairwise-summed arithmetic mean, as ``x1.mean()`` in the reference.
    """
    x = np.asarray(x, dtype=np.float64)
    xc = x - x.mean(axis=1, keepdims=True)
    nrm = np.sqrt(np.einsum("ij,ij->i", xc, xc))
    return xc, nrm


def correlation_matrix(rna_np: np.ndarray, dna_np: np.ndarray) -> np.ndarray:
    """``corrs[i, j] = dot(r_i - mean, d_j - mean) / (1e-10 + |r_i - mean| |d_j - mean|)``.

    Vectorised form of the double loop at ``macrodna.py:103-107`` with the
    formula of ``cosine_similarity_np`` (``:24-25``).  rows = RNA cells, columns
    = DNA cells (``:102``).  A zero-variance cell gives exactly 0.0.
    """
    rc, rn = standardise(rna_np)
    dc, dn = standardise(dna_np)
    return (rc @ dc.T) / (EPS + rn[:, None] * dn[None, :])


def correlation_matrix_literal(rna_np: np.ndarray, dna_np: np.ndarray) -> np.ndarray:
    """The literal per-pair loop of ``macrodna.py:103-107`` (small cases only)."""
    from numpy.linalg import norm

    rna_np = np.asarray(rna_np)
    dna_np = np.asarray(dna_np)
    out = np.empty((rna_np.shape[0], dna_np.shape[0]))
    for i in range(rna_np.shape[0]):
        for j in range(dna_np.shape[0]):
            x1, x2 = rna_np[i], dna_np[j]
            out[i, j] = np.dot(x1 - x1.mean(), x2 - x2.mean()) / (
                EPS + norm(x1 - x1.mean()) * norm(x2 - x2.mean())
            )
    return out


def n_steps(n_rna: int, n_dna: int) -> int:
    """``macrodna.py:118-123``."""
    q, r = divmod(n_rna, n_dna)
    return int(q) + (1 if r != 0 else 0)


def lap_step(corrs: np.ndarray, rna_idx: np.ndarray):
    """One ``ilp`` call (``macrodna.py:27-84``) on rows ``rna_idx`` x all columns.

    Returns (matched global rna rows, matched dna columns, objective).
    """
    sub = corrs[rna_idx, :]
    r, c = linear_sum_assignment(sub, maximize=True)
    return rna_idx[r], c, float(sub[r, c].sum())


def step_loop(corrs: np.ndarray):
    """The step loop of ``macrodna.py:110-145`` in O(M) bookkeeping.

    Returns ``assign[M]`` (DNA column per RNA row), ``step[M]`` (1-based) and the
    per-step objective list (``m.objVal`` of each ``ilp`` call, as the variant
    ``random_assignment_test.py:91,141-142`` returns).
    """
    n_rna, n_dna = corrs.shape
    assign = np.full(n_rna, -1, dtype=np.int32)
    step = np.zeros(n_rna, dtype=np.int32)
    objs = []
    rna_idx = np.arange(n_rna)
    for s in range(n_steps(n_rna, n_dna)):
        rows, cols, obj = lap_step(corrs, rna_idx)
        assign[rows] = cols
        step[rows] = s + 1
        objs.append(obj)
        rna_idx = np.flatnonzero(assign < 0)  # macrodna.py:141-145, ascending
    if (assign < 0).any():
        raise ValueError("unassigned RNA cell")  # list.index(1) at macrodna.py:160
    return assign, step, np.asarray(objs, dtype=np.float64)


def leave_one_out(corrs: np.ndarray, cell_idx: int, k_steps: int):
    """``leave_one_out`` of Resampling_stability_analyses/BE_data_analyses/run_loo_experiment.py:201-319 on a
    given correlation matrix: delete row ``cell_idx`` (:224), run the step loop on the rest (:246-269), then give
    the left-out cell the DNA cell of highest correlation among those with at most ``k_steps - 1`` matches
    (:296-309).  Returns (assign_rest, step_rest, objs_rest, test_dna or -1, test_corr, total objective)."""
    loo = np.delete(corrs, cell_idx, 0)
    if loo.shape[0]:
        assign, step, objs = step_loop(loo)
    else:
        assign, step, objs = np.empty(0, np.int32), np.empty(0, np.int32), np.empty(0)
    cell = corrs[cell_idx]
    cnt = np.bincount(assign, minlength=corrs.shape[1])
    test_dna, test_val = -1, 0.0
    for idx_ in np.argsort(cell)[::-1]:
        if cnt[idx_] <= k_steps - 1:
            test_dna, test_val = int(idx_), float(cell[idx_])
            break
    total = float(np.sum(objs)) + (test_val if test_dna >= 0 else 0.0)
    return assign, step, objs, test_dna, test_val, total


def random_assign(corrs: np.ndarray, rng) -> float:
    """One draw of ``random_test.assign`` (random_assignment_test.py:233-258) with ``rng.choice`` in place of
    ``np.random.choice``: per step a random N-subset of the remaining RNA rows against all DNA columns, the last
    (short) step a random injective map of the remaining rows into the DNA columns; returns the sum."""
    n_rna, n_dna = corrs.shape
    rna_idx = np.arange(n_rna)
    dna_idx = np.arange(n_dna)
    s = 0.0
    removed = []
    for _ in range(n_steps(n_rna, n_dna)):
        n_min = min(len(rna_idx), len(dna_idx))
        if n_min == len(rna_idx):
            sel = rng.choice(dna_idx, n_min, replace=False)
            s += corrs[rna_idx, sel].sum()
        else:
            sel = rng.choice(rna_idx, n_min, replace=False)
            s += corrs[sel, dna_idx].sum()
            removed.append(sel)
            rna_idx = np.delete(np.arange(n_rna), np.concatenate(removed))
    return float(s)


def cell2cell_arrays(rna_np: np.ndarray, dna_np: np.ndarray):
    """Array-level restatement: cells x genes float64 in, (corrs, assign, step, objs) out."""
    corrs = correlation_matrix(rna_np, dna_np)
    assign, step, objs = step_loop(corrs)
    return corrs, assign, step, objs


def frames_from_vectors(rna_cells, dna_cells, assign, step):
    """Result assembly of ``macrodna.py:149-186`` from the two int vectors."""
    pred = [dna_cells[j] for j in assign]
    res = pd.DataFrame(list(zip(pred, rna_cells)), columns=["predict_cell", "cell"]).set_index("cell")
    tagged = pd.DataFrame(
        list(zip(pred, rna_cells, [int(s) for s in step])), columns=["predict_cell", "cell", "step"]
    ).set_index("cell")
    return res, tagged


class OracleMaCroDNA:
    """Frame-level restatement with the reference's constructor and method names."""

    def __init__(self, rna_df=None, dna_df=None, dna_label=None, clone_column="predict_clone"):
        self.rna_df = rna_df
        self.dna_df = dna_df
        self.dna_label = dna_label
        self.clone_column = clone_column
        self.last = None

    def cell2cell_assignment(self):
        dna_cells = list(self.dna_df.columns)  # macrodna.py:87
        rna_cells = list(self.rna_df.columns)  # macrodna.py:88
        genes = shared_genes(self.rna_df, self.dna_df)
        self.dna_df = self.dna_df.loc[genes, :]  # macrodna.py:90 (mutates self)
        self.rna_df = self.rna_df.loc[genes, :]  # macrodna.py:91
        dna_np = np.ascontiguousarray(self.dna_df.T.to_numpy(dtype=np.float64))  # :93
        rna_np = np.ascontiguousarray(self.rna_df.T.to_numpy(dtype=np.float64))  # :94
        corrs, assign, step, objs = cell2cell_arrays(rna_np, dna_np)
        self.last = dict(corrs=corrs, assign=assign, step=step, objs=objs)
        return frames_from_vectors(rna_cells, dna_cells, assign, step)

    def cell2clone_assignment(self):
        rna_result, _ = self.cell2cell_assignment()  # macrodna.py:190
        lab = self.dna_label.set_index("cell")  # :195
        rna_result[self.clone_column] = lab.loc[rna_result["predict_cell"]]["clone"].tolist()  # :197-198
        return rna_result


def tiny_frames(variant: str = "src"):
    """The literal frames of ``tiny_test``: ``macrodna.py:202-217`` ("src", 4 RNA x 4 DNA,
    RNA carries an extra gene g7) and ``CRC_data_analysis/macrodna.py:199-216`` ("crc",
    5 RNA x 4 DNA -> 2 steps)."""
    dna = pd.DataFrame.from_dict(
        {
            "cell1": [2, 2, 3, 1, 6, 2],
            "cell2": [2, 2, 2, 2, 2, 2],
            "cell3": [1, 1, 2, 2, 2, 3],
            "cell4": [2, 2, 2, 2, 2, 6],
            "gene": ["g1", "g2", "g3", "g4", "g5", "g6"],
        }
    ).set_index("gene")
    if variant == "src":
        rna = pd.DataFrame.from_dict(
            {
                "cell1": [0, 0, 10, 0, 20, 0, 0],
                "cell2": [2, 2, 2, 2, 2, 2, 0],
                "cell3": [0, 0, 2, 2, 0, 5, 0],
                "cell4": [1, 1, 1, 1, 1, 20, 0],
                "gene": ["g1", "g2", "g3", "g4", "g5", "g6", "g7"],
            }
        ).set_index("gene")
    elif variant == "crc":
        rna = pd.DataFrame.from_dict(
            {
                "cell1": [0, 0, 10, 0, 20, 0],
                "cell2": [2, 2, 2, 2, 2, 2],
                "cell3": [0, 0, 2, 2, 0, 5],
                "cell4": [1, 1, 0, 0, 1, 20],
                "cell5": [1, 1, 1, 1, 0, 0],
                "gene": ["g1", "g2", "g3", "g4", "g5", "g6"],
            }
        ).set_index("gene")
    else:
        raise ValueError(variant)
    lab = pd.DataFrame.from_dict({"clone": [0, 1, 2, 3], "cell": ["cell1", "cell2", "cell3", "cell4"]})
    return rna, dna, lab
