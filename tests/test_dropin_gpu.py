"""The drop-in claim end to end (last row): the UNMODIFIED reference training script's own functions
-- train_one_epoch / validate / loss_function of vessel_analysis/01_train/train.py, with torch's clip_grad_norm_ and
Adam -- driven through `python -m causal_vae_b200.run`, whose finder makes the script's `from models import ...`
resolve to the native classes, and cross-checked against the unmodified reference modules on the same GPU.

Needs baseline/_ref (python baseline/install_reference.py in the build container; it ships to the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.path.join(ROOT, "baseline", "_ref")

DRIVER = r'''
import json, os, sys, types
for n in ["matplotlib", "matplotlib.pyplot", "tifffile", "skimage", "skimage.measure", "skimage.morphology", "seaborn"]:
    m = types.ModuleType(n); m.__path__ = []; sys.modules[n] = m
ds = types.ModuleType("dataset"); ds.VesselDataset = object; sys.modules["dataset"] = ds      # file discovery: host code
import importlib.util
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from config import CONFIG                      # the reference's config.py (00_core)
CONFIG["IMG_HEIGHT"] = CONFIG["IMG_WIDTH"] = 64
CONFIG["DEVICE"] = torch.device("cuda")
import train as T                              # the reference's train.py, byte for byte
import models                                  # what its `from models import ...` got: the native classes
assert models.CausalViTVAE.__module__.startswith("causal_vae_b200"), models.CausalViTVAE.__module__
assert T.CausalVesselVAE.__module__.startswith("causal_vae_b200")
sys.path.insert(0, ROOT)
from oracle import cvae_oracle as O            # deterministic weights / inputs only


class Loader(list):
    dataset = range(8)
loader = Loader([O.vessel_inputs(4, 64, 64, seed=s)[:3] for s in (3, 4)])
sd = O.fill_state_dict(O.vessel_shapes(64, 64), seed=0)

def run(make_model):
    vae = make_model().to(CONFIG["DEVICE"])
    vae.load_state_dict(sd)
    torch.manual_seed(11)
    v0 = T.validate(vae, loader)               # eval mode: the only random draw is eps (same generator, same shape)
    for mod in vae.modules():                  # dropout off so the two implementations see the same training problem
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if hasattr(mod, "in_proj_weight"):
            mod.dropout = 0.0
    opt = torch.optim.Adam(vae.parameters(), lr=CONFIG["LEARNING_RATE"])
    losses = []
    for ep in range(3):
        torch.manual_seed(100 + ep)
        losses.append(T.train_one_epoch(ep, vae, loader, opt))
    torch.manual_seed(12)
    return {"val0": v0, "train": losses, "val1": T.validate(vae, loader)}

native = run(models.CausalViTVAE)
core = os.path.join(REF, "vessel_analysis", "00_core")
spec = importlib.util.spec_from_file_location("ref_models_file", os.path.join(core, "models.py"))
ref_models = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_models)
assert not ref_models.CausalViTVAE.__module__.startswith("causal_vae_b200")
reference = run(ref_models.CausalViTVAE)
print("RESULT " + json.dumps({"native": native, "reference": reference}))
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="baseline/_ref not installed")
def test_reference_training_loop_on_native_models(tmp_path):
    # mirror of the reference layout: .../vessel_analysis/{00_core -> reference, 01_train/{train.py -> reference, driver.py}}
    va = tmp_path / "vessel_analysis"
    (va / "01_train").mkdir(parents=True)
    os.symlink(os.path.join(REF, "vessel_analysis", "00_core"), va / "00_core")
    os.symlink(os.path.join(REF, "vessel_analysis", "01_train", "train.py"), va / "01_train" / "train.py")
    driver = va / "01_train" / "driver.py"
    driver.write_text(f"ROOT = {ROOT!r}\nREF = {REF!r}\n" + DRIVER)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    out = subprocess.run([sys.executable, "-m", "causal_vae_b200.run", str(driver)], capture_output=True, text=True,
                         env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    r = json.loads(line[len("RESULT "):])
    nat, ref = r["native"], r["reference"]
    # the reference's own validate() on native modules vs on its own modules: same weights, same eps draw
    assert abs(nat["val0"] - ref["val0"]) <= 2e-5 * abs(ref["val0"]), (nat["val0"], ref["val0"])
    # its own train_one_epoch (loss_function, backward, clip_grad_norm_(5), Adam) drives both down the same path
    # Two fp32 implementations follow the same trajectory only until rounding differences have been amplified by Adam's
    # m / sqrt(v) normalisation (a gradient component near zero flips the direction of its first updates): observed over
    # repeated runs 2e-5 / 2e-4 / 3e-3 relative difference of the epoch losses (ours varies run to run with the order of fp32
    # atomics).  The bound grows with the epoch accordingly; the first epoch -- before any divergence -- stays tight.
    for (a, b), tol in zip(zip(nat["train"], ref["train"]), (2e-4, 2e-3, 2e-2)):
        assert abs(a - b) <= tol * abs(b), (nat["train"], ref["train"])
    assert nat["train"][-1] < nat["train"][0]
    assert abs(nat["val1"] - ref["val1"]) <= 5e-2 * abs(ref["val1"]), (nat["val1"], ref["val1"])
