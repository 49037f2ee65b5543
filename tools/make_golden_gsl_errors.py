"""tests/golden/gsl_errors.json: what the compiled reference (oracle/_ref) PRINTS for pathological rows of
get_loglike_matrix — src/error.c:45-48 reports every failing GSL call (file, line, reason) and carries on.  The rows and
the captured text (file names stripped to their base name, as an R build of the package prints them) are the fixture the
device's per-cell error log is compared with (tests/test_gpu_parity.py::test_gsl_error_log_matches_the_reference)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, numpy as np
sys.path.insert(0, %r)
from oracle import ref
d = json.loads(sys.argv[1])
r = ref.get_loglike_matrix(np.array(d["phi"]), np.array(d["expected"]), np.array(d["total"], np.int32), np.array(d["observed"], np.int32), 1.0)
sys.stderr.flush()
print(json.dumps(np.where(np.isnan(r), None, r).tolist()))
''' % ROOT

CASES = {
    # per-bin vectors: healthy rows between failing ones; phi >= ~0.83 makes a1 negative for the low copy-number state, phi > 1 for all
    # (phi = 1 puts both shape parameters at exactly 0: the x == 0 site of src/beta.c:56 instead of the sign test of :44)
    "per_bin": dict(phi=[0.93, 0.001, 0.93, 0.5, 2.0, 0.93, 0.01, 0.99, 1.0, 1.0], expected=[0.2, 0.2, 0.2, 0.2, 0.2, 0.9, 0.3, 0.5, 0.2, 0.4],
                    total=[50, 80, 0, 50, 50, 30, 1000, 7, 5, 9], observed=[10, 20, 0, 10, 10, 3, 200, 7, 0, 4]),
    # one (phi, expected) pair for all bins, as new('ExomeDepth') hands it over (R/class_definition.R:119, 168)
    "constant": dict(phi=[0.93] * 5, expected=[0.2] * 5, total=[50, 0, 12, 7, 100], observed=[10, 0, 12, 0, 50]),
}

out = {}
for name, d in CASES.items():
    p = subprocess.run([sys.executable, "-c", CHILD, json.dumps(d)], capture_output=True, text=True, check=True)
    text = "\n".join(line.replace(os.path.dirname(l.split()[1]) + "/", "") if (l := line).startswith("ERROR /") else line
                     for line in p.stderr.splitlines())
    out[name] = dict(inputs=d, printed=text + ("\n" if text else ""), ll=json.loads(p.stdout.strip().splitlines()[-1]))
    print(name, len(text.splitlines()), "lines")
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "gsl_errors.json"), "w"), indent=1)
