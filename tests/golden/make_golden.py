"""How tests/golden/reference_golden.json was made.

The reference (MatchCake v0.2.3) is pure Python on PennyLane + torch_pfaffian; neither is installed in
the build image and there is no network, so the reference cannot be *run* to generate vectors.  The
fixture therefore holds the known answers the reference itself PUBLISHES for this path (JOSS paper,
tutorial notebook outputs, and the known-answer tests of its own test-suite), transcribed by hand.

Running this script with /root/reference mounted re-checks every transcribed number against the file
it was taken from (it greps the printed digits), so a typo in the fixture cannot go unnoticed:

    python tests/golden/make_golden.py
"""
import json
import os
import re
import sys

REF = os.environ.get("MATCHCAKE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _read(rel):
    with open(os.path.join(REF, rel), "r", encoding="utf-8") as f:
        return f.read()


def _fmt(v, decimals):
    s = f"{v:.{decimals}f}"
    return s


def main():
    g = json.load(open(os.path.join(HERE, "reference_golden.json")))
    if not os.path.isdir(REF):
        print(f"{REF} not mounted: nothing to cross-check")
        return 0
    checks = []
    checks.append(("joss/paper.md", "Expectation value: " + g["joss_expval"]["printed"]))
    nb = _read("tutorials/probabilities.ipynb")
    for key in ("probs_basis_101", "probs_plus", "probs_mixed"):
        for v in g[key]["value"]:
            if v != 0.0:
                checks.append(("tutorials/probabilities.ipynb", _fmt(v, 6)))
    checks.append(("tutorials/probabilities.ipynb", "0.7614 0.2386"))
    checks.append(("tutorials/probabilities.ipynb", "0.755987"))
    for k in ("expval_basis_101", "expval_plus", "expval_mixed"):
        checks.append(("tutorials/expectation_values.ipynb", _fmt(g["hamiltonian"][k], 6)))
    checks.append(("tutorials/iris_classification.ipynb", "90.00%"))
    bad = 0
    cache = {}
    for rel, needle in checks:
        txt = cache.setdefault(rel, _read(rel))
        ok = needle in txt
        bad += not ok
        print(("ok   " if ok else "MISSING ") + f"{rel}: {needle}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
