"""Golden fixture for the latent-translator Ridge stage: runs the LIVE reference function
`latent_translator/analysis.py::fit_translator_ridge` (scikit-learn Ridge + LeaveOneOut, the pinned third-party
algorithm: sklearn 1.9.0 in the build container) on seeded data and commits its outputs.

    python tests/golden/make_ridge_golden.py      # writes tests/golden/ridge_loocv.json
"""
import importlib.util
import json
import os

import numpy as np

REF = "/root/reference/latent_translator/analysis.py"
spec = importlib.util.spec_from_file_location("ref_lt_analysis", REF)
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

rng = np.random.default_rng(0)
N, D, Fm = 16, 512, 6
Z = rng.standard_normal((N, D)).astype(np.float32)
Wtrue = rng.standard_normal((D, Fm)) * 0.05
M = (Z @ Wtrue + 0.1 * rng.standard_normal((N, Fm)) + np.array([1.0, -2.0, 0.5, 0.0, 3.0, -1.0])).astype(np.float32)
names = [f"f{j}" for j in range(Fm)]
model, metrics, Mhat, W = mod.fit_translator_ridge(Z, M, feature_names=names, alpha=1.0)
import sklearn
out = {"sklearn": sklearn.__version__, "seed": 0, "N": N, "D": D, "F": Fm, "alpha": 1.0,
       "metrics": metrics.to_dict(orient="records"),
       "Mhat": np.asarray(Mhat, dtype=np.float64).tolist(),
       "W_absmax": float(np.abs(W).max()), "W_sample": np.asarray(W, dtype=np.float64)[:, ::37].tolist(),
       "intercept": np.asarray(model.intercept_, dtype=np.float64).tolist()}
with open(os.path.join(os.path.dirname(__file__), "ridge_loocv.json"), "w") as f:
    json.dump(out, f)
print("wrote ridge_loocv.json", metrics.head(3).to_dict(orient="records"))
