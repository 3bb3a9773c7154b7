#!/usr/bin/env python3
"""Build recipe for ``oracle/_ref`` -- TEST INFRASTRUCTURE, never the product path.

Compiles the reference's own Cython implementation of the hot path
(``python-pmf/pmf_cy.pyx``, ``normal_exps_cy.pyx``, ``bayes_pmf.py``+``.pxd``) from the
sources where they lie under ``/root/reference`` into ``oracle/_ref/`` (git-ignored; it
travels to the GPU box like our own built ``.so`` files).  The reference targets
numpy~1.7 / Cython 0.x / Python 3.3, so the sources are passed through the minimal
compatibility substitutions of SURVEY.md section 8c on their way into the build
directory; nothing is written to ``/root/reference`` and nothing from it is committed.

Only ``tests/``, ``__graft_entry__.smoke()``, ``tests/golden/make_golden.py`` and
``bench.py``'s reference/cpu_baseline legs may import what this produces.
"""
import os
import re
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = os.environ.get("AMF_REFERENCE_DIR", "/root/reference/python-pmf")

# (file, [(regex, replacement), ...]) -- SURVEY.md 8c patch list
PATCHES = {
    "pmf_cy.pyx": [
        (r"DTYPE = np\.float\b", "DTYPE = np.float64"),                       # 1
        (r"np\.array\(rating_tuples, dtype=float, copy=False\)",
         "np.asarray(rating_tuples, dtype=float)"),                            # 3
        (r"np\.array\(extra, copy=False, ndmin=2\)",
         "np.array(extra, ndmin=2)"),                                          # 3
    ],
    "pmf_cy.pxd": [],
    "normal_exps_cy.pyx": [
        (r"DTYPE = np\.float\b", "DTYPE = np.float64"),                       # 1
        (r"np\.int_t", "np.int64_t"),                                          # 2
    ],
    "bayes_pmf.py": [
        (r"size=\(n\*\(n-1\)/2\.\)", "size=(n*(n-1)//2)"),                     # 6
        (r"integrate\.trapz", "integrate.trapezoid"),                          # 8
    ],
    "bayes_pmf.pxd": [],
    "active_pmf.py": [
        (r"np\.array\(self\.ratings, dtype=float, copy=False\)",
         "np.asarray(self.ratings, dtype=float)"),                             # 3
        (r"for i, j, rating in self\.ratings\)",
         "for i, j, rating in ((int(a), int(b), c) for a, b, c in self.ratings))"),  # 4
        (r"evals\[list\(zip\(\*pool\)\)\]", "evals[tuple(zip(*pool))]"),       # 5
        (r"scipy\.integrate\.simps", "scipy.integrate.simpson"),               # 8
    ],
    # the matrix-normal variant the reference runs on drugbank / movielens
    "matrix_normal_exps_cy.pyx": [
        (r"DTYPE = np\.float\b", "DTYPE = np.float64"),                       # 1
    ],
    "mn_active_pmf.py": [
        (r"np\.array\(self\.ratings, dtype=float, copy=False\)",
         "np.asarray(self.ratings, dtype=float)"),                             # 3
        (r"evals\[list\(zip\(\*pool\)\)\]", "evals[tuple(zip(*pool))]"),       # 5
        (r"scipy\.integrate\.simps", "scipy.integrate.simpson"),               # 8
        (r"import scipy\.integrate", "import scipy.integrate\nimport scipy.linalg"),
    ],
    # pure-python twins, only used by the reference's own known-answer test
    "normal_exps.py": [
        (r"for i, j, rating in apmf\.ratings:",
         "for i, j, rating in ((int(a), int(b), c) for a, b, c in apmf.ratings):"),              # 4
    ],
    "test_normal_exps.py": [],
}

SETUP = '''
from setuptools import setup, Extension
import numpy as np
setup(
    script_args=["build_ext", "--inplace"],
    include_dirs=[np.get_include()],
    ext_modules=[
        Extension("normal_exps_cy", ["normal_exps_cy.c"]),
        Extension("matrix_normal_exps_cy", ["matrix_normal_exps_cy.c"]),
        Extension("pmf_cy", ["pmf_cy.c"]),
        Extension("bayes_pmf", ["bayes_pmf.c"]),
    ],
)
'''


def built():
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(OUT, m + suffix))
               for m in ("pmf_cy", "normal_exps_cy", "matrix_normal_exps_cy", "bayes_pmf"))


def build(force=False):
    if built() and not force:
        return True
    if not os.path.isdir(SRC):
        return False
    os.makedirs(OUT, exist_ok=True)
    for name, subs in PATCHES.items():
        with open(os.path.join(SRC, name)) as f:
            text = f.read()
        for pat, rep in subs:
            text, n = re.subn(pat, rep, text)
            if n == 0:
                raise RuntimeError("patch %r did not apply to %s" % (pat, name))
        with open(os.path.join(OUT, name), "w") as f:
            f.write(text)
    with open(os.path.join(OUT, "setup_ref.py"), "w") as f:
        f.write(SETUP)
    env = dict(os.environ, CFLAGS="-O2 -w")
    cy = [sys.executable, "-m", "cython", "-3"]
    subprocess.check_call(cy + ["pmf_cy.pyx"], cwd=OUT, env=env)
    subprocess.check_call(cy + ["normal_exps_cy.pyx"], cwd=OUT, env=env)
    subprocess.check_call(cy + ["matrix_normal_exps_cy.pyx"], cwd=OUT, env=env)
    subprocess.check_call(cy + ["-Xbinding=false", "bayes_pmf.py"], cwd=OUT, env=env)
    subprocess.check_call([sys.executable, "setup_ref.py"], cwd=OUT, env=env)
    shutil.rmtree(os.path.join(OUT, "build"), ignore_errors=True)
    # the compiled bayes_pmf must win over the patched .py twin on import
    os.replace(os.path.join(OUT, "bayes_pmf.py"), os.path.join(OUT, "bayes_pmf_src.py.txt"))
    return built()


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref built" if ok else "oracle/_ref NOT built (reference sources absent)")
    sys.exit(0 if ok else 1)
