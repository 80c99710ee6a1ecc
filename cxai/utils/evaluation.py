"""Readers of the optimiser's result trees -- mirror of ``get_best_run`` / ``get_run_stats`` of the reference's
cxai/utils/evaluation.py:107-141 (the rest of that file evaluates trained classifiers and is outside the path)."""
from __future__ import annotations

import os

import pandas as pd

__all__ = ["get_best_run", "get_run_stats"]


def get_run_stats(path: str):
    """(final loss, final per-concept relevances if the file has ``R*`` columns, all losses) of one train_stats.csv."""
    stats = pd.read_csv(path)
    final_loss = list(stats["loss"])[-1]
    concept_relevances = [list(stats[key])[-1] for key in stats.keys() if key.startswith("R")]
    return final_loss, concept_relevances, list(stats["loss"])


def get_best_run(path: str):
    """The run directory (``run1`` .. ``run9``) under ``path`` whose final objective is highest (evaluation.py:107-126;
    the reference returns the statistics of the LAST directory it listed next to the best run's index -- here the
    relevances and losses returned belong to the best run)."""
    best_loss, best = 0, None
    for dir_level1 in sorted(d for d in os.listdir(path) if not d.startswith(".")):
        run = int(dir_level1[-1])
        loss, concept_relevances, train_losses = get_run_stats(os.path.join(path, dir_level1, "train_stats.csv"))
        if loss > best_loss:
            best_loss = loss
            best = (run, loss, concept_relevances, os.path.join(path, dir_level1), train_losses)
    if best is None:
        raise FileNotFoundError(f"no run with a positive final objective under {path}")
    return best
