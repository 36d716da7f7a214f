#!/usr/bin/env python
"""Fold gpurun_out/parity_r2.jsonl (written by tests/conftest.py:parity during `pytest -m gpu` on the B200 box) into
profiles/parity_r2.json: for every test and tensor the measured error beside its tolerance (1e-5 fp32 mode / 2e-2 bf16 mode),
and for the ill-conditioned fp32 reductions the kernel's and the fp32 reference's distance from the fp64 truth.
--merge: keep the tests already in the destination that the new ledger does not contain (the 2-GPU tests run on another box than
the 1-GPU suite)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    args = [a for a in sys.argv[1:] if a != "--merge"]
    src = args[0] if len(args) > 0 else os.path.join(ROOT, "gpurun_out", "parity_r2.jsonl")
    dst = args[1] if len(args) > 1 else os.path.join(ROOT, "profiles", "parity_r2.json")
    by_test = {}
    if "--merge" in sys.argv and os.path.exists(dst):
        by_test = json.load(open(dst))["tests"]
    for line in open(src):
        r = json.loads(line)
        by_test.setdefault(r.pop("test"), {})[r.pop("tensor")] = r   # a re-run overwrites the earlier line
    worst = {}
    for t, d in by_test.items():
        for k, r in d.items():
            tol = r["tol"]
            w = worst.setdefault(str(tol), {"tol": tol, "n": 0, "n_within_tol_of_ref": 0, "n_judged_by_a_stated_exception": 0, "worst_err_vs_ref": 0.0})
            w["n"] += 1
            if r["err_vs_ref"] <= tol:
                w["n_within_tol_of_ref"] += 1
                w["worst_err_vs_ref"] = max(w["worst_err_vs_ref"], r["err_vs_ref"])
            else:
                w["n_judged_by_a_stated_exception"] += 1     # the fp64-truth criterion / the bf16 emulation floor
    out = {"metric": "per tensor max|a-b| / max|b|", "summary_by_tolerance": worst, "failed": [f"{t}:{k}" for t, d in by_test.items() for k, r in d.items() if not r["pass"]],
           "tests": by_test}
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst, {k: (v["n"], v["n_within_tol_of_ref"], v["n_judged_by_a_stated_exception"]) for k, v in worst.items()})


if __name__ == "__main__":
    main()
