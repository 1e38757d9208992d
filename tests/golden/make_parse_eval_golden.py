"""Golden for evaluate.parse_eval_results: the reference's own parse_eval_results.py (plain Python 3-compatible, no TensorFlow) run in
the build container on an eval-detailed.txt written from seeded random rows.  Output: tests/golden/parse_eval_results.json
(the rows, the sample ids and the script's stdout)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from spatialaudiogen_b200 import evaluate as E  # noqa: E402

REF = '/root/reference/parse_eval_results.py'


def main():
    rng = np.random.RandomState(7)
    ids, rows = [], []
    for vid, n in (('vidB', 5), ('vidA', 7), ('vidC', 3)):
        for t in rng.permutation(n):                               # windows out of time order, videos out of name order
            ids.append('%s %.1f' % (vid, 0.5 + t))
            rows.append(np.abs(rng.randn(28)) * np.linspace(0.01, 2, 28))
    rows = np.asarray(rows)
    fn = os.path.join(tempfile.mkdtemp(), 'eval-detailed.txt')
    E.write_eval_detailed(fn, ids, rows)
    out = subprocess.run([sys.executable, REF, fn], capture_output=True, text=True, check=True)
    json.dump({'ids': ids, 'rows': rows.tolist(), 'reference_stdout': out.stdout,
               'generator': 'tests/golden/make_parse_eval_golden.py: %s run on the file written from these rows' % REF},
              open(os.path.join(ROOT, 'tests', 'golden', 'parse_eval_results.json'), 'w'))
    print(out.stdout)


if __name__ == '__main__':
    main()
