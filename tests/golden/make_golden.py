"""Writes tests/golden/reference_tests.json and tests/golden/oracle_fixtures.npz.

Part 1 (reference_tests.json): the known-answer vectors of the reference's OWN rackunit suite for the hot path,
transcribed by hand from /root/reference/flomat/flomat.rkt (the `module+ test` at :3317-3660) -- the reference is Racket
and cannot be executed in this image, so these literals (with their line citations) are what pins the oracle.
Matrices are written row-major as in the Racket literals (`list->flomat`, `flomat/dim`).

Part 2 (oracle_fixtures.npz): seeded inputs with outputs of the ORACLE (OpenBLAS provider) at sizes a CPU test can
afford; they pin the oracle's own behaviour over time (regression) and give the GPU tests committed expected outputs.
These are oracle-generated, not reference-generated, and say so in their key names.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, os.pardir, os.pardir))

REFERENCE_TESTS = {
    "source": "soegaard/sci flomat/flomat.rkt (module+ test), transcribed with line numbers",
    "gemm": [
        {"cite": "flomat.rkt:3335-3340, 3655-3657", "A": [[1, 2], [3, 4]], "B": [[5, 6], [7, 8]], "AB": [[19, 22], [43, 50]]},
        {"cite": "flomat.rkt:3342-3347, 3658-3660", "A": [[1, 2], [3, 4]], "B": [[5, 6, 7], [8, 9, 10]],
         "AB": [[21, 24, 27], [47, 54, 61]]},
        {"cite": "flomat.rkt:3349-3353", "A": [[0, 0, 1], [0, 0, 0]], "B": [[1, 2], [3, 4], [5, 6]], "AB": [[5, 6], [0, 0]]},
    ],
    "expt": {"cite": "flomat.rkt:3421-3430", "A": [[1, 2], [3, 4]],
             "powers": {"0": [[1, 0], [0, 1]], "1": [[1, 2], [3, 4]], "2": [[7, 10], [15, 22]], "3": [[37, 54], [81, 118]],
                        "8": [[165751, 241570], [362355, 528106]]}},
    "plus_minus": {"cite": "flomat.rkt:3405-3419", "A": [[1, 2], [3, 4]], "B": [[5, 6], [7, 8]], "A+B": [[6, 8], [10, 12]],
                   "A-B": [[-4, -4], [-4, -4]], "-A": [[-1, -2], [-3, -4]]},
    "scale": {"cite": "flomat.rkt:3566-3569", "s": 2, "A": [[1, 2], [3, 4]], "sA": [[2, 4], [6, 8]]},
    "solve": {"cite": "flomat.rkt:3547-3551", "M": [[1, 5], [2, 3]], "b": [[5], [5]], "check": "M * solve(M,b) == b (abs 1e-13)"},
    "inverse": {"cite": "flomat.rkt:3552-3557", "M": [[1, 2], [3, 4]], "check": "M*inv(M) == I and inv(M)*M == I (abs 1e-13)"},
    "det": [
        {"cite": "flomat.rkt:3560", "M": [[3]], "det": 3.0, "tol": 0.0},
        {"cite": "flomat.rkt:3561", "M": [[1, 2], [3, 4]], "det": -2.0, "tol": 0.0},
        {"cite": "flomat.rkt:3562", "M": [[1, 2, 3], [4, 5, 6], [7, 8, 9]], "det": 0.0, "tol": 1e-5},
        {"cite": "flomat.rkt:3563", "M": [[1, 2, 3], [4, -5, 6], [7, 8, 9]], "det": 120.0, "tol": 1e-5},
        {"cite": "flomat.rkt:3564-3565", "M": [[1, 2, 3, 4], [-5, 6, 7, 8], [9, 10, -11, 12], [13, 14, 15, 16]], "det": 5280.0,
         "tol": 1e-5},
    ],
    "plu": {"cite": "flomat.rkt:3609-3617", "M": [[1, 1, 0, 3], [2, 1, -1, 1], [3, -1, -1, 2], [-1, 2, 3, -1]],
            "check": "P*L*U == M (abs 1e-13)"},
    "identity": {"cite": "flomat.rkt:3361-3366", "n": [1, 2, 3]},
    "const": {"cite": "flomat.rkt:3369", "m": 2, "n": 3, "x": 0.0},
    "flomat_epsilon": {"cite": "flomat.rkt:160, 204-215", "abs_tol": 1e-13},
    "expm_doc_examples": {
        "cite": "flomat-doc/manual-flomat.scrbl:823-826, 1716-1719 (outputs are NOT stored in the reference)",
        "note": "closed forms used instead: exp(diag(1,2)) = diag(e, e^2)",
        "cases": [{"A": [[1, 0], [0, 2]]}, {"A": [[1, 2], [3, 4]]}],
    },
    "pade8_coefficients_hex": {
        "cite": "expm.rkt:136-146 (exact rationals (16-i)! 8! / (16! i! (8-i)!) converted to double)",
        "c": ["0x1.0000000000000p+0", "0x1.0000000000000p-1", "0x1.ddddddddddddep-4", "0x1.1111111111111p-6",
              "0x1.a41a41a41a41ap-10", "0x1.c01c01c01c01cp-14", "0x1.45e5d2ba42ea0p-18", "0x1.29f6b20961c00p-23",
              "0x1.08db48ebe51c7p-29"],
    },
}


def main():
    with open(os.path.join(HERE, "reference_tests.json"), "w") as f:
        json.dump(REFERENCE_TESTS, f, indent=1)
    from oracle.flomat_oracle import Flomat, OpenBLASProvider

    F = Flomat(OpenBLASProvider(threads=1))  # single thread: reproducible summation order
    out = {}
    rng = np.random.default_rng(20261019)
    for n in (8, 33, 64, 130):
        a = np.asfortranarray(rng.standard_normal((n, n)) / np.sqrt(n))
        b = np.asfortranarray(rng.standard_normal((n, 5)))
        out[f"expm_in_{n}"] = a
        out[f"oracle_expm_out_{n}"] = F.expm(a)
        out[f"oracle_expm_s_{n}"] = np.array(F.scaling_exponent(a))
        m = np.asfortranarray(rng.standard_normal((n, n)))
        out[f"lu_in_{n}"] = m
        out[f"rhs_in_{n}"] = b
        out[f"oracle_solve_out_{n}"] = F.mldivide(m, b)
        out[f"oracle_inv_out_{n}"] = F.inv(m)
        out[f"oracle_det_out_{n}"] = np.array(F.det(m))
        lu, ipiv, info = F.lu(m)
        out[f"oracle_ipiv_{n}"] = ipiv
    # small-norm quirk: s <= 0 scales up and never squares
    q = np.asfortranarray(np.diag([0.1, 0.05]))
    out["expm_in_quirk"] = q
    out["oracle_expm_out_quirk"] = F.expm(q)
    np.savez_compressed(os.path.join(HERE, "oracle_fixtures.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
