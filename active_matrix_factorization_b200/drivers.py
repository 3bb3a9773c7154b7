"""Runs the reference's OWN driver functions against the GPU-backed classes.

The reference keeps its experiment drivers (``full_test``, ``compare`` / ``compare_active``,
``make_fake_data``, ``get_ratings``, ``main`` and the argparse tables) in the same files as the
model classes (python-pmf/active_pmf.py:796-1257, bayes_pmf.py:675-938,
mn_active_pmf.py:785-1132).  SURVEY.md section 2 keeps those callers as they are, so this
package does not restate them: ``load()`` executes the reference's source file from a checkout
the user points at, with ``pmf_cy`` / ``normal_exps_cy`` / ``matrix_normal_exps_cy`` resolved to
the mirrors of this package, then rebinds every class, registry and helper the mirror module
defines.  The driver functions look those names up in their module globals when they run, so
they construct and drive the GPU-backed classes.

Two textual substitutions are made on the way in, both because a model that lives on a GPU
must not be handed to forked worker processes:

* ``import multiprocessing as mp`` / ``multiprocessing.Pool`` -> ``InProcessPool`` below (same
  ``apply / map / apply_async / map_async / close / join`` surface, runs the call in the
  calling thread; the candidate pool is the data-parallel axis and it is already one launch);
* the numpy-2 / scipy-1.1x spellings of SURVEY.md 8c that occur inside driver functions.

    from active_matrix_factorization_b200 import drivers
    apmf = drivers.load("active_pmf", "/path/to/active-matrix-factorization/python-pmf")
    apmf.main()          # the reference's command line, GPU-backed classes
"""
import importlib
import os
import re
import sys
import types

_MIRRORS = {
    "active_pmf": "active_matrix_factorization_b200.active_pmf",
    "mn_active_pmf": "active_matrix_factorization_b200.mn_active_pmf",
    "bayes_pmf": "active_matrix_factorization_b200.bayes_pmf",
}
_BACKENDS = {
    "pmf_cy": "active_matrix_factorization_b200.pmf_cy",
    "normal_exps_cy": "active_matrix_factorization_b200.normal_exps_cy",
    "matrix_normal_exps_cy": "active_matrix_factorization_b200.matrix_normal_exps_cy",
}
_SUBS = [
    (r"import multiprocessing as mp\b",
     "from active_matrix_factorization_b200 import drivers as mp"),
    (r"\bmultiprocessing\.Pool\(", "InProcessPool("),
    (r"from multiprocessing import Pool\b",
     "from active_matrix_factorization_b200.drivers import InProcessPool as Pool"),
    (r"np\.array\(([^()]*), dtype=float, copy=False\)", r"np.asarray(\1, dtype=float)"),
    (r"evals\[list\(zip\(\*pool\)\)\]", "evals[tuple(zip(*pool))]"),
    (r"scipy\.integrate\.simps\b", "scipy.integrate.simpson"),
    (r"integrate\.trapz\b", "integrate.trapezoid"),
    (r"size=\(n\*\(n-1\)/2\.\)", "size=(n*(n-1)//2)"),
]


class _Now(object):
    def __init__(self, value):
        self._value = value

    def get(self, timeout=None):
        return self._value

    def wait(self, timeout=None):
        pass

    def ready(self):
        return True

    def successful(self):
        return True


class InProcessPool(object):
    """``multiprocessing.Pool`` surface used by the reference drivers, without processes."""

    def __init__(self, processes=None, *args, **kwargs):
        self.processes = processes

    def apply(self, fn, args=(), kwds=None):
        return fn(*args, **(kwds or {}))

    def apply_async(self, fn, args=(), kwds=None, callback=None):
        res = fn(*args, **(kwds or {}))
        if callback is not None:
            callback(res)
        return _Now(res)

    def map(self, fn, iterable, chunksize=None):
        return [fn(x) for x in iterable]

    imap = map

    def map_async(self, fn, iterable, chunksize=None, callback=None):
        res = self.map(fn, iterable)
        if callback is not None:
            callback(res)
        return _Now(res)

    def close(self):
        pass

    def join(self):
        pass

    terminate = close


Pool = InProcessPool        # `import ... drivers as mp; mp.Pool(n)`


def _source_path(name, ref_dir):
    for cand in (name + ".py", name + "_src.py.txt"):   # oracle/_ref keeps bayes_pmf's twin as .txt
        p = os.path.join(ref_dir, cand)
        if os.path.exists(p):
            return p
    raise FileNotFoundError("no %s.py under %s" % (name, ref_dir))


def load(name, ref_dir=None):
    """Module object holding the reference's ``name`` (``active_pmf``, ``mn_active_pmf`` or
    ``bayes_pmf``) driver functions bound to this package's classes."""
    if name not in _MIRRORS:
        raise ValueError("unknown driver module %r" % (name,))
    ref_dir = ref_dir or os.environ.get("AMF_REFERENCE_DIR")
    if not ref_dir:
        raise ValueError("pass the reference's python-pmf directory (or set AMF_REFERENCE_DIR)")
    path = _source_path(name, ref_dir)
    with open(path) as f:
        text = f.read()
    for pat, rep in _SUBS:
        text = re.sub(pat, rep, text)
    mirror = importlib.import_module(_MIRRORS[name])
    mod = types.ModuleType("amf_b200_reference_drivers." + name)
    mod.__file__ = path
    mod.InProcessPool = InProcessPool
    saved = {}
    try:
        for short, full in list(_BACKENDS.items()) + [(k, v) for k, v in _MIRRORS.items() if k != name]:
            saved[short] = sys.modules.get(short)
            sys.modules[short] = importlib.import_module(full)
        exec(compile(text, path, "exec"), mod.__dict__)
    finally:
        for short, old in saved.items():
            if old is None:
                sys.modules.pop(short, None)
            else:
                sys.modules[short] = old
    # the mirror's classes, registries and helpers replace the reference's own definitions
    for key, val in vars(mirror).items():
        if not key.startswith("__"):
            setattr(mod, key, val)
    return mod
