"""CPU-only checks: the C-ABI library loads and exports every symbol include/cvae_b200.h declares,
the drop-in modules expose the reference's state_dict keys (golden tables recorded from the live
reference), there is no CPU fallback, and the product never imports the oracle."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def built():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()


def test_header_symbols_exported():
    from causal_vae_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cvae_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t)\s+(cvae_\w+)\s*\(", hdr, flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(_lib.lib, name), name
    assert _lib.lib.cvae_version() >= 100 and _lib.lib.cvae_built_arch() == 100
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cvae_\w+)", out))
    assert declared <= exported


def test_sass_is_sm100a():
    from causal_vae_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_state_dict_keys_match_reference():
    from causal_vae_b200.vessel import models
    for tag in ("vessel_64x64_b4", "vessel_256x256_b8"):
        g = json.load(open(os.path.join(G, tag + ".json")))
        models.CONFIG["IMG_HEIGHT"], models.CONFIG["IMG_WIDTH"] = g["config"]["H"], g["config"]["W"]
        sd = models.CausalViTVAE().state_dict()
        assert {k: list(v.shape) for k, v in sd.items()} == g["state_dict_shapes"]
        assert list(sd.keys()) == list(g["state_dict_shapes"].keys()), "key order"
        assert sd["backbone.stem.1.num_batches_tracked"].dtype == torch.int64


def test_cnn_variant_state_dict_keys_match_reference():
    """CausalVesselVAE (vessel_analysis/00_core/models.py:9-166): key names, order and shapes of the live reference."""
    from causal_vae_b200.vessel import models
    g = json.load(open(os.path.join(G, "vessel_cnn_768x1280_b4.json")))
    models.CONFIG.update(Z_DIM=g["config"]["z_dim"], M_DIM=g["config"]["m_dim"], T_DIM=g["config"]["t_dim"])
    model = models.CausalVesselVAE()
    sd = model.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == g["state_dict_shapes"]
    assert list(sd.keys()) == list(g["state_dict_shapes"].keys()), "key order"
    assert not hasattr(model, "dec_adapter") and hasattr(model, "dec_fc") and len(model.dec_conv) == 27
    assert isinstance(model.dec_conv[0], torch.nn.Upsample) and isinstance(model.enc_conv[21], torch.nn.Flatten)


def test_no_cpu_fallback():
    from causal_vae_b200 import nn
    with pytest.raises(RuntimeError, match="CUDA"):
        nn.Conv2d(1, 8, 3, 2, 1)(torch.zeros(1, 1, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        nn.Sequential(nn.Linear(8, 8), nn.LeakyReLU(0.2))(torch.zeros(2, 8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "causal_vae_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle/", "").lower() or f == "__init__.py" and False, \
                    f"{f} mentions the oracle"


def test_sequential_fusion_plan():
    from causal_vae_b200 import nn
    from causal_vae_b200.chain import ResUnit, Unit
    seq = nn.Sequential(nn.ConvTranspose2d(8, 8, 3, 2, 1, 1), nn.BatchNorm2d(8), nn.LeakyReLU(), nn.ResBlock(8),
                        nn.Conv2d(8, 1, 3, padding=1))
    plan = seq._plan()
    assert len(plan) == 1 and plan[0][0] == "chain"
    u = plan[0][1]
    assert isinstance(u[0], Unit) and u[0].bn is seq[1] and u[0].act == pytest.approx(0.01)
    assert isinstance(u[1], ResUnit) and u[1].u1.act == pytest.approx(0.2) and u[1].u2.act is None
    assert isinstance(u[2], Unit) and u[2].bn is None
    mlp = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Dropout(0.1), nn.Linear(16, 8), nn.Dropout(0.1))
    kinds = [k for k, _ in mlp._plan()]
    assert kinds == ["chain", "mod", "mod", "chain", "mod"]


def test_runner_redirects_models_for_an_unmodified_script(tmp_path):
    """SURVEY 8b last row: `python -m causal_vae_b200.run <reference script>` makes bare `import models` /
    `from vit_backbone import ViTVAE` resolve to the native classes while `config` stays the script tree's own."""
    core = tmp_path / "vessel_analysis" / "00_core"
    work = tmp_path / "vessel_analysis" / "01_train"
    core.mkdir(parents=True); work.mkdir(parents=True)
    (core / "config.py").write_text(
        "import torch\nCONFIG = dict(DEVICE=torch.device('cpu'), IMG_HEIGHT=64, IMG_WIDTH=64, T_DIM=19, M_DIM=12, "
        "Z_DIM=128, BETA=0.5, LEARNING_RATE=1e-4, BATCH_SIZE=8, EPOCHS=1, LAMBDA_MORPH=10000)\n")
    (core / "models.py").write_text("raise RuntimeError('the reference models.py must not be imported')\n")
    (core / "vit_backbone.py").write_text("raise RuntimeError('the reference vit_backbone.py must not be imported')\n")
    (work / "script.py").write_text(
        "import sys, os, json\n"
        "sys.path.append(os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '00_core')))\n"
        "from models import CausalVesselVAE, CausalViTVAE\n"
        "from vit_backbone import ViTVAE\n"
        "from config import CONFIG\n"
        "import models\n"
        "CONFIG['IMG_HEIGHT'] = 96\n"
        "m = CausalViTVAE()\n"
        "json.dump({'cls': CausalViTVAE.__module__, 'vit': ViTVAE.__module__, 'same_config': models.CONFIG is CONFIG,\n"
        "           'pos': list(m.backbone.pos_embedding.shape), 'argv': sys.argv[1:], 'name': __name__}, open(sys.argv[1], 'w'))\n")
    out = tmp_path / "out.json"
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "causal_vae_b200.run", str(work / "script.py"), str(out)],
                       capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.load(open(out))
    assert got["cls"] == "causal_vae_b200.vessel.models" and got["vit"] == "causal_vae_b200.vessel.vit_backbone"
    assert got["same_config"] and got["name"] == "__main__" and got["argv"] == [str(out)]
    assert got["pos"] == [1, (96 // 32) * (64 // 32) + 1, 256]          # the script's CONFIG edit reached the native model

    from causal_vae_b200 import run
    assert run.family_of("/x/mnist_test/06_model_experiment/main.py") == "mnist06"
    assert run.family_of("/x/mnist_test/01_baseline_causal_vae/main.py") == "mnist01"
    assert run.family_of("/x/causal_cascade/main.py") == "cascade"
    assert run.family_of("/x/latent_translator/main.py") == "latent_translator"


def test_step_roofline_bytes_follow_the_counting_rule():
    """bench.py's BYTES_PER_SAMPLE (77 MB activations + 8.8 MB parameter traffic per sample at B = 64)
    against the layer-table derivation of scripts/algorithmic_bytes.py."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("algbytes", os.path.join(ROOT, "scripts", "algorithmic_bytes.py"))
    ab = importlib.util.module_from_spec(spec); spec.loader.exec_module(ab)
    r = ab.vessel_step_bytes(256, 256, 64)
    assert r["stem_KB"] == [2048, 1024, 512, 256, 64] and sum(r["decoder_KB"]) == 8128 and r["skip_KB"] == 896
    assert r["params"] == 14065897 and abs(r["param_MB_per_sample"] - 8.8) < 0.05
    src = open(os.path.join(ROOT, "bench.py")).read()
    m = re.search(r"BYTES_PER_SAMPLE\s*=\s*([0-9.e+]+)\s*\+\s*([0-9.e+]+)", src)
    bench_bytes = float(m.group(1)) + float(m.group(2))
    # the rule gives 80.7 + 8.8 MB in decimal bytes; bench.py uses SURVEY's MiB-rounded 77 + 8.8 (stricter): within 5 %
    assert bench_bytes <= r["total_MB_per_sample"] * 1e6 <= 1.05 * bench_bytes
