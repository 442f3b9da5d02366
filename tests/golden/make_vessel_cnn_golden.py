"""Golden fixture for the CNN vessel model (vessel_analysis/00_core/models.py:9-166) from the LIVE reference.

    python tests/golden/make_vessel_cnn_golden.py        # writes tests/golden/vessel_cnn_768x1280_b4.json

The unmodified `CausalVesselVAE` (fixed 768x1280 input) and `loss_function` (01_train/train.py:18-60) run on the
deterministic weights of `oracle.cvae_oracle.fill_state_dict` and the synthetic vessel batch; outputs, losses and
per-parameter gradient summaries are committed, plus the reference's own fp32-vs-fp64 gradient discrepancy.
"""
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(__file__))
import make_golden as MG  # noqa: E402
from make_golden import O, REF, EpsInjector, load_ref, noise_floor, shapes_of, summarize  # noqa: E402

B, H, W = 4, 768, 1280
NAMES = ["recon_x", "m_hat", "mu", "logvar", "m_mu", "m_logvar"]


def main():
    MG._stub(["matplotlib", "matplotlib.pyplot", "tifffile", "skimage", "skimage.measure", "skimage.morphology",
              "seaborn", "tqdm"])
    sys.modules["tqdm"].tqdm = lambda x, **k: x
    core = os.path.join(REF, "vessel_analysis/00_core")
    sys.modules.pop("config", None)
    sys.path.insert(0, core)
    import config as vcfg  # noqa
    vcfg.CONFIG["DEVICE"] = torch.device("cpu")
    models = load_ref(os.path.join(core, "models.py"), "ref_vessel_models", purge=("models", "vit_backbone"))
    sys.modules.setdefault("dataset", types.ModuleType("dataset")).VesselDataset = object
    sys.modules["models"] = models
    train = load_ref(os.path.join(REF, "vessel_analysis/01_train/train.py"), "ref_vessel_train", purge=())
    sys.path.remove(core)
    zd, md, td = vcfg.CONFIG["Z_DIM"], vcfg.CONFIG["M_DIM"], vcfg.CONFIG["T_DIM"]

    model = models.CausalVesselVAE()
    shp = shapes_of(model)
    want = O.vessel_cnn_shapes(zd, md, td)
    assert {k: tuple(v) for k, v in shp.items()} == want and list(shp) == list(want), "oracle shape table drifted"
    P = O.fill_state_dict(want, seed=0)
    model.load_state_dict(P, strict=True)
    x, m, t, eps = O.vessel_inputs(B, H, W, md, td, zd, seed=0)
    rec = {"config": {"H": H, "W": W, "B": B, "seed": 0, "beta": 0.5, "z_dim": zd, "m_dim": md, "t_dim": td},
           "state_dict_shapes": shp}

    model.eval()
    with torch.no_grad(), EpsInjector(eps):
        outs = model(x, m, t)
    rec["eval"] = {n: summarize(o) for n, o in zip(NAMES, outs)}

    model.train()
    with EpsInjector(eps):
        outs = model(x, m, t)
    recon, kld, morph, sp = train.loss_function(outs[0], x, outs[1], m, outs[2], outs[3], outs[4], outs[5])
    loss = recon + 0.5 * kld + morph + 0.3 * sp
    loss.backward()
    g32 = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    rec["train"] = {"loss": loss.item(), "recon": recon.item(), "kld": kld.item(), "morph": morph.item(),
                    "sparsity": sp.item(), "outputs": {n: summarize(o) for n, o in zip(NAMES, outs)},
                    "grads": {k: summarize(v) for k, v in g32.items()},
                    "running": {k: summarize(v) for k, v in model.state_dict().items()
                                if k.endswith(("running_mean", "running_var"))}}
    m64 = models.CausalVesselVAE().double()
    m64.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in P.items()})
    m64.train()
    with EpsInjector(eps.double()):
        o64 = m64(x.double(), m.double(), t.double())
    r64 = train.loss_function(o64[0], x.double(), o64[1], m.double(), o64[2], o64[3], o64[4], o64[5])
    l64 = r64[0] + 0.5 * r64[1] + r64[2] + 0.3 * r64[3]
    l64.backward()
    rec["train"]["loss_fp64"] = l64.item()
    rec["train"]["grad_noise_fp32_vs_fp64"] = noise_floor(
        g32, {k: p.grad for k, p in m64.named_parameters() if p.grad is not None})
    with open(os.path.join(os.path.dirname(__file__), "vessel_cnn_768x1280_b4.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print("loss", rec["train"]["loss"], "fp64", rec["train"]["loss_fp64"],
          "max grad noise", max(rec["train"]["grad_noise_fp32_vs_fp64"].values()))


if __name__ == "__main__":
    main()
