"""Parity gates shared by the GPU tests.

north_star's gate is normwise:  ||X_gpu - X_ref||_1 / ||X_ref||_1 <= 64 n eps  (tol(n)).  Every assert message carries the measured
err / tol(n) ratio, and FLOMAT_PARITY_REPORT=<file> appends one JSON line per gate so a run leaves a table of the ratios
(profiles/r02_parity_ratios.jsonl comes from that).

Solves: two backward-stable solutions of the same system may differ by cond(A) times their backward errors, whatever the
implementation.  `gate_solve` therefore gates at 1 x tol(n) and only if that fails falls back to the perturbation bound
2 cond_1(A) (eta_gpu + eta_ref) with the MEASURED cond and the MEASURED backward errors eta = ||A X - B||_1 / (||A||_1 ||X||_1) --
never a blanket slack factor.
"""
from __future__ import annotations

import json
import os

import numpy as np

from oracle.flomat_oracle import normwise_rel_err, tol


def _report(rec: dict) -> None:
    path = os.environ.get("FLOMAT_PARITY_REPORT")
    if path:
        with open(path, "a") as f:
            f.write(json.dumps(rec) + "\n")


def gate(name: str, got, want, n: int, factor: float = 1.0) -> float:
    """||got - want||_1 / ||want||_1 <= factor * 64 n eps; returns the ratio err / tol(n)."""
    err = normwise_rel_err(got, want)
    ratio = err / tol(n)
    _report({"gate": name, "n": n, "err": err, "tol": tol(n), "ratio": ratio, "factor": factor})
    assert err <= factor * tol(n), f"{name}: normwise error {err:.3e} = {ratio:.3f} x 64 n eps (n = {n}, allowed {factor} x)"
    return ratio


def backward_error(a, x, b) -> float:
    r = np.abs(a @ x - b).sum(axis=0).max()
    return float(r / (np.abs(a).sum(axis=0).max() * np.abs(x).sum(axis=0).max()))


def gate_solve(name: str, a, b, x_gpu, x_ref, n: int) -> float:
    """Forward parity of a solve at 1 x tol(n); if (and only if) that fails, the cond-scaled perturbation bound with measured numbers.
    The backward error of the GPU solution is always gated at tol(n)."""
    err = normwise_rel_err(x_gpu, x_ref)
    ratio = err / tol(n)
    eta_g, eta_r = backward_error(a, x_gpu, b), backward_error(a, x_ref, b)
    rec = {"gate": name, "n": n, "err": err, "tol": tol(n), "ratio": ratio, "eta_gpu": eta_g, "eta_ref": eta_r}
    assert eta_g <= tol(n), f"{name}: backward error {eta_g:.3e} exceeds 64 n eps = {tol(n):.3e}"
    if err <= tol(n):
        rec["bound"] = "64 n eps"
        _report(rec)
        return ratio
    cond = float(np.linalg.cond(a, 1))
    bound = 2.0 * cond * (eta_g + eta_r)
    rec.update({"bound": "2 cond1 (eta_gpu + eta_ref)", "cond1": cond, "cond_bound": bound})
    _report(rec)
    assert err <= bound, (f"{name}: forward error {err:.3e} = {ratio:.2f} x 64 n eps and above the perturbation bound {bound:.3e} "
                          f"(cond_1 = {cond:.3e}, backward errors {eta_g:.2e} / {eta_r:.2e})")
    return ratio


def gate_inverse(name: str, a, x_gpu, x_ref, n: int) -> float:
    """Same two-tier rule for `inv`: 1 x tol(n), else cond_1(A) times the measured residuals ||A X - I||_1 / (||A||_1 ||X||_1)."""
    err = normwise_rel_err(x_gpu, x_ref)
    ratio = err / tol(n)
    rec = {"gate": name, "n": n, "err": err, "tol": tol(n), "ratio": ratio}
    if err <= tol(n):
        rec["bound"] = "64 n eps"
        _report(rec)
        return ratio
    eye = np.eye(n)
    eta_g, eta_r = backward_error(a, x_gpu, eye), backward_error(a, x_ref, eye)
    cond = float(np.linalg.cond(a, 1))
    bound = 2.0 * cond * (eta_g + eta_r)
    rec.update({"bound": "2 cond1 (eta_gpu + eta_ref)", "cond1": cond, "cond_bound": bound, "eta_gpu": eta_g, "eta_ref": eta_r})
    _report(rec)
    assert err <= bound, (f"{name}: inverse differs by {err:.3e} = {ratio:.2f} x 64 n eps, above the perturbation bound {bound:.3e} "
                          f"(cond_1 = {cond:.3e}, residuals {eta_g:.2e} / {eta_r:.2e})")
    return ratio
