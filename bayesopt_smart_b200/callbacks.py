"""Observer callbacks with the reference's class names and ``callback(state)`` contract
(reference bayesopt/callbacks.py).  These are host-side conveniences, not part of the hot path;
``state`` keys are those built at bayesian_optimization.py:226-243.
"""
from __future__ import annotations

from datetime import datetime
from typing import Optional

import numpy as np


class PlotterCallback:
    """Forward 2-D states to a plotter object exposing ``plot(...)`` (callbacks.py:19-41)."""

    def __init__(self, plotter):
        self.plotter = plotter

    def __call__(self, state):
        if state["x_vector"].shape[1] == 2:
            self.plotter.plot(x_vector=state["x_vector"], y_vector=state["y_vector"],
                              mu_objectives=state["mu_objectives"],
                              variance_objectives=state["variance_objectives"],
                              acquisition_values=state["acquisition_values"], x_next=state.get("x_next"))


class ProgressLogger:
    """Track the best value per objective and log one line per iteration (callbacks.py:44-145)."""

    def __init__(self, log_file: Optional[str] = None, verbose: bool = True):
        self.log_file = log_file
        self.verbose = verbose
        self.best_per_objective = None
        self.best_x_per_objective = []
        self.history = []
        if log_file:
            with open(log_file, "w", encoding="utf-8") as f:
                f.write("iteration,n_evaluations,time_total\n")

    def __call__(self, state):
        y = state["y_vector"]
        x = state["x_vector"]
        best = y.max(axis=0)
        arg = y.argmax(axis=0)
        self.best_per_objective = best
        self.best_x_per_objective = [x[i].copy() for i in arg]
        total = state["timings"]["total"]
        self.history.append((state["iteration"], state["n_evaluations"], total))
        if self.verbose:
            print(f"[iter {state['iteration']:4d}] evals={state['n_evaluations']:4d} "
                  f"best={np.array2string(best, precision=4)} t={total:.4f}s")
        if self.log_file:
            with open(self.log_file, "a", encoding="utf-8") as f:
                f.write(f"{state['iteration']},{state['n_evaluations']},{total:.6f}\n")


class OptimizationLogger:
    """Verbose per-iteration report including the stage timings (callbacks.py:148-200)."""

    def __init__(self, verbose: bool = True):
        self.verbose = verbose
        self.records = []

    def __call__(self, state):
        t = state["timings"]
        self.records.append(dict(iteration=state["iteration"], x_next=np.array(state["x_next"]),
                                 hyperparams=np.array(state["hyperparams"]), timings=dict(t)))
        if self.verbose:
            print(f"{datetime.now().strftime('%H:%M:%S')} iteration {state['iteration']}: "
                  f"hyperparams {t['hyperparams']:.4f}s | kernels {t['kernels']:.4f}s | "
                  f"acquisition {t['acquisition']:.4f}s | eval {t['eval']:.4f}s | total {t['total']:.4f}s")


class PerformanceMonitor:
    """Accumulate the five stage timings and summarise them (callbacks.py:203-245)."""

    KEYS = ("hyperparams", "kernels", "acquisition", "eval", "total")

    def __init__(self):
        self.timings = {k: [] for k in self.KEYS}

    def __call__(self, state):
        for k in self.KEYS:
            self.timings[k].append(state["timings"][k])

    def summary(self) -> dict:
        return {k: (float(np.mean(v)) if v else 0.0) for k, v in self.timings.items()}

    def print_summary(self) -> None:
        s = self.summary()
        total = s["total"] or 1.0
        print("Performance summary (average per iteration):")
        for k in self.KEYS:
            print(f"  {k:12s} {s[k]:.4f}s ({100.0 * s[k] / total:5.1f}%)")
