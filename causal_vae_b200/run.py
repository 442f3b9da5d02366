"""Run an unmodified reference script against the native models:

    python -m causal_vae_b200.run /path/to/causal-vae/vessel_analysis/01_train/main.py [script args...]
    python -m causal_vae_b200.run --family mnist06 some_script.py

The reference's scripts import bare `models` / `vit_backbone` from their own directory (`sys.path[0]`, which outranks
PYTHONPATH) or from `../00_core` (last row).  The runner puts the script's directories on `sys.path` the
way `python script.py` would, seeds `sys.modules["models"]` (and `"vit_backbone"`) with modules that re-export the
native classes of the script's experiment family, and then executes the script as `__main__`.  Everything else the
script imports (`config`, `dataset`, `train`, `utils` ...) stays the reference's own; the native models read the
reference's `config.CONFIG` dict (same object), so `CONFIG[...] = ...` edits made by the script are honoured.
"""
import importlib
import os
import runpy
import sys
import types

FAMILIES = {
    # family: {module name the scripts import: (native module, {alias: native name})}
    "vessel": {"models": ("causal_vae_b200.vessel.models", {}),
               "vit_backbone": ("causal_vae_b200.vessel.vit_backbone", {})},
    "mnist01": {"models": ("causal_vae_b200.mnist.models", {})},
    "mnist06": {"models": ("causal_vae_b200.mnist.models", {"CausalMorphVAE12": "CausalMorphVAE12Prob"})},
    "cascade": {"models": ("causal_vae_b200.cascade.models", {})},
    "latent_translator": {"models": ("causal_vae_b200.latent_translator.models", {})},
}


def family_of(script):
    """Experiment family from the script's location in the reference tree."""
    parts = os.path.abspath(script).replace("\\", "/").split("/")
    if "vessel_analysis" in parts:
        return "vessel"
    if "causal_cascade" in parts:
        return "cascade"
    if "latent_translator" in parts:
        return "latent_translator"
    if "mnist_test" in parts:
        sub = parts[parts.index("mnist_test") + 1] if parts.index("mnist_test") + 1 < len(parts) else ""
        return "mnist06" if sub.startswith("06") else "mnist01"
    raise SystemExit(f"cannot tell the experiment family of {script}; pass --family {{{', '.join(FAMILIES)}}}")


def install(family, script_dir):
    """Seed sys.modules with the native re-exports of `family`.  Returns the installed module names."""
    if family not in FAMILIES:
        raise SystemExit(f"unknown family {family!r}; one of {', '.join(FAMILIES)}")
    if script_dir not in sys.path:
        sys.path.insert(0, script_dir)
    core = os.path.abspath(os.path.join(script_dir, "..", "00_core"))     # vessel scripts append it themselves
    if family == "vessel" and os.path.isdir(core) and core not in sys.path:
        sys.path.append(core)
    visible = os.environ.get("CUDA_VISIBLE_DEVICES")
    done = []
    for name, (target, aliases) in FAMILIES[family].items():
        native = importlib.import_module(target)          # imports the reference's `config` if it is on the path
        mod = types.ModuleType(name)
        mod.__dict__.update({k: v for k, v in vars(native).items() if not k.startswith("__")})
        for alias, real in aliases.items():
            setattr(mod, alias, getattr(native, real))
        mod.__file__ = native.__file__
        mod.__native__ = target
        sys.modules[name] = mod
        done.append(name)
    # mnist_test/*/config.py:4 pins CUDA_VISIBLE_DEVICES="0" at import: keep the launcher's choice (torchrun ranks)
    if visible is not None:
        os.environ["CUDA_VISIBLE_DEVICES"] = visible
    return done


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    family = None
    if argv and argv[0] == "--family":
        family, argv = argv[1], argv[2:]
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    install(family or family_of(script), os.path.dirname(script))
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
