"""Generate tests/golden/fad.npz by running the REFERENCE's own FAD arithmetic (fadtk/fad.py, fadtk/utils.py).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_fad_golden.py

fadtk/fad.py and fadtk/utils.py import hypy_utils and the embedding-model loaders (absent here), so the functions are
cut out of the reference files with `ast` (read-only, unmodified) and executed with the names they use bound to the
same libraries (numpy, scipy.linalg, numpy.lib.scimath.sqrt), exactly as tests/golden/make_istft_golden.py does for the
export chain:

  * calc_embd_statistics          fadtk/fad.py:41-47
  * calc_frechet_distance         fadtk/fad.py:50-119   (scipy 1.18 removed sqrtm's `disp` argument: the call is
                                   served by a shim that returns (sqrtm, 0.0) like the old signature -- it only feeds a
                                   log message; the returned distance comes from the eigenvalue method)
  * FrechetAudioDistanceTK.score_inf   fadtk/fad.py:303-350, with a stand-in `self` (load_stats -> the baseline
                                   statistics; `ml.name`) and the embeddings read from .npy files as fadtk caches them
  * _process_file / calculate_embd_statistics_online   fadtk/utils.py:13-46 (pmap -> map)

Inputs are seeded (tests/stubs.py::fad_embeddings), so the fixture stores the reference OUTPUTS only.
"""
import ast
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
from numpy.lib.scimath import sqrt as scisqrt
from scipy import linalg as _linalg

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
FAD_SRC = "/root/reference/fadtk/fad.py"
UTILS_SRC = "/root/reference/fadtk/utils.py"


class _Log:
    def info(self, *a, **k):
        pass

    warning = error = info


class _Linalg:
    """scipy.linalg with the pre-1.18 `sqrtm(..., disp=False) -> (sqrtm, errest)` call form the reference uses."""

    def __getattr__(self, name):
        return getattr(_linalg, name)

    @staticmethod
    def sqrtm(a, disp=True, **kw):
        r = _linalg.sqrtm(a, **kw)
        return r if disp else (r, 0.0)


def _functions(src, names, ns):
    """execute the named top-level functions / methods of `src` (unmodified source text) in namespace `ns`"""
    tree = ast.parse(open(src).read())
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            node.decorator_list = []
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
            node.returns = None
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(ast.fix_missing_locations(mod), src, "exec"), ns)
            found[node.name] = ns[node.name]
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f"{missing} not found in {src}")
    return found


def reference_fad():
    from typing import NamedTuple

    class FADInfResults(NamedTuple):  # fadtk/fad.py:29-33
        score: float
        slope: float
        r2: float
        points: list

    ns = {"np": np, "linalg": _Linalg(), "scisqrt": scisqrt, "log": _Log(), "tq": lambda it, **k: it,
          "FADInfResults": FADInfResults}
    f = _functions(FAD_SRC, {"calc_embd_statistics", "calc_frechet_distance", "score_inf"}, ns)
    ns2 = {"np": np, "pmap": lambda fn, items, **k: [fn(i) for i in items]}
    u = _functions(UTILS_SRC, {"_process_file", "calculate_embd_statistics_online"}, ns2)
    return f, u


from tests.stubs import FAD_CASES, fad_embeddings  # noqa: E402  (seeded inputs, rebuilt by the tests)

if __name__ == "__main__":
    f, u = reference_fad()
    out = {}
    tmp = Path(tempfile.mkdtemp(prefix="fad_golden_"))
    for name, (n1, n2, d, parts) in FAD_CASES.items():
        a, b = fad_embeddings(name)           # fp16 (n, d), as fadtk caches embeddings (model_loader.py:46-48)
        mu1, c1 = f["calc_embd_statistics"](a)
        mu2, c2 = f["calc_embd_statistics"](b)
        out[name + "_mu1"], out[name + "_cov1"] = mu1, c1       # mu is fp16 (np.mean keeps the dtype), cov float64
        out[name + "_mu2"], out[name + "_cov2"] = mu2, c2
        out[name + "_fd"] = np.float64(np.real(f["calc_frechet_distance"](mu1, c1, mu2, c2)))
        files = []
        for i, chunk in enumerate(np.array_split(a, parts)):
            p = tmp / f"{name}_{i}.npy"
            np.save(p, chunk)
            files.append(p)
        omu, ocov = u["calculate_embd_statistics_online"](files)
        out[name + "_online_mu"], out[name + "_online_cov"] = omu, ocov
        print(name, a.shape, b.shape, "fd", float(out[name + "_fd"]))
    # FAD-inf: numpy's global generator draws the bootstrap indices (fad.py:331): seeded here and in the tests
    name = "inf"
    a, b = fad_embeddings("d128")
    mu_b, cov_b = f["calc_embd_statistics"](b)
    files = []
    for i, chunk in enumerate(np.array_split(a, 3)):
        p = tmp / f"inf_{i}.npy"
        np.save(p, chunk)
        files.append(p)
    self = types.SimpleNamespace(load_stats=lambda baseline: (mu_b, cov_b), ml=types.SimpleNamespace(name="stub"))
    np.random.seed(1234)
    r = f["score_inf"](self, None, files, steps=6, min_n=200)
    out["inf_score"], out["inf_slope"], out["inf_r2"] = np.float64(r.score), np.float64(r.slope), np.float64(r.r2)
    out["inf_points"] = np.array(r.points, dtype=np.float64)
    print("inf", r.score, r.slope, r.r2)
    np.savez_compressed(os.path.join(HERE, "fad.npz"), **out)
    print("fad.npz:", len(out), "arrays")
