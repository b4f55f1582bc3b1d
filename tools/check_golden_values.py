#!/usr/bin/env python
"""Compare the scalars a run logged against a golden-values file (reference ``tools/check_golden_values.py`` /
``tests/functional_tests/python_test_utils``).

    python tools/check_golden_values.py --run-dir runs/tb --golden tests/functional_tests/test_cases/gpt/gpt_tiny_tp1_cpu/golden_values_cpu.json \\
        [--rtol 1e-4] [--exact lm\\ loss] [--approx iteration-time:0.05]

Exit code 0 when every golden series matches (``--exact`` tags bit for bit, everything else within ``--rtol``; ``--approx tag:tol`` overrides per tag)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "functional_tests", "python_test_utils"))


def main(argv=None) -> int:
    from run_case import read_scalars

    ap = argparse.ArgumentParser()
    ap.add_argument("--run-dir", required=True, help="directory the run wrote its scalars to (--tensorboard-dir)")
    ap.add_argument("--golden", required=True)
    ap.add_argument("--rtol", type=float, default=1e-4)
    ap.add_argument("--exact", nargs="*", default=[])
    ap.add_argument("--approx", nargs="*", default=[], help="tag:relative_tolerance")
    a = ap.parse_args(argv)
    got = read_scalars(a.run_dir)
    golden = json.load(open(a.golden))
    tol = {t.rsplit(":", 1)[0]: float(t.rsplit(":", 1)[1]) for t in a.approx}
    bad = []
    for tag, series in golden.items():
        if tag not in got:
            bad.append(f"{tag}: missing from the run")
            continue
        for step, ref in series.items():
            val = got[tag].get(int(step))
            if val is None:
                bad.append(f"{tag}@{step}: missing")
            elif tag in a.exact:
                if val != ref:
                    bad.append(f"{tag}@{step}: {val!r} != {ref!r} (exact)")
            elif abs(val - ref) > tol.get(tag, a.rtol) * max(abs(ref), 1e-12):
                bad.append(f"{tag}@{step}: {val} vs {ref} (rtol {tol.get(tag, a.rtol)})")
    for b in bad:
        print("MISMATCH", b)
    print(f"{'FAILED' if bad else 'PASSED'}: {sum(len(s) for s in golden.values())} golden points, {len(bad)} mismatches")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
