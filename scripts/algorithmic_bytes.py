"""The frozen counting rule behind `step_roofline`: algorithmic HBM bytes per sample of one vessel
training step, derived from the model's own layer table so that builder and judge compute the same figure.

    python scripts/algorithmic_bytes.py [H W B]        # default 256 256 64

Rule.  fp32 storage.  Every tensor that MUST cross a kernel boundary is counted once written + once read.  Training-mode
BatchNorm makes every conv / linear output that feeds a BatchNorm such a tensor (its statistics need the whole batch
before the activation can be applied), so forward = 2 x (conv outputs of the stem and the decoder) + the ResBlock skip
re-reads + the input image read twice (stem and loss) + the transformer trunk (fused per block).  Backward reads each
saved tensor once and moves one gradient in and out per tensor = 2 x forward.  Per STEP (not per sample): weights read in
forward and backward, weight gradients written, fused clip + Adam at 28 B / parameter.

The layer table comes from the state_dict shapes (`oracle.cvae_oracle.vessel_shapes`, pinned to the reference's own
key / shape table by tests/test_oracle_golden.py): stem = five stride-2 convolutions from H x W, decoder = five stride-2
transposed convolutions from (H/32) x (W/32) with a ResBlock (two convolutions) after stages 1-3 and a final 16 -> 1
convolution (vit_backbone.py:74-90,119-156).
"""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)


def vessel_step_bytes(H=256, W=256, B=64, tokens_dim=256, depth=6, mlp_dim=512):
    from oracle import cvae_oracle as O
    shapes = O.vessel_shapes(H, W)
    f = 4  # fp32
    # ---- stem: conv k -> output [Cout, H/2^k, W/2^k]
    stem, h, w = [], H, W
    for i in (0, 3, 6, 9, 12):
        cout = shapes[f"backbone.stem.{i}.weight"][0]
        h, w = h // 2, w // 2
        stem.append(cout * h * w * f)
    gh, gw = h, w
    # ---- decoder: decoder_input -> [256, gh, gw]; ConvT stages double the map; ResBlocks keep it
    dec = [shapes["backbone.decoder_input.weight"][0] * f]
    skip = 0
    h, w = gh, gw
    for i in (0, 4, 8, 12, 15):
        cout = shapes[f"backbone.decoder.{i}.weight"][1]          # ConvTranspose2d weight is (Cin, Cout, k, k)
        h, w = 2 * h, 2 * w
        dec.append(cout * h * w * f)
        rb = f"backbone.decoder.{i + 3}.conv.0.weight"
        if rb in shapes:                                          # ResBlock: two conv outputs + the skip re-read
            dec += [cout * h * w * f] * 2
            skip += cout * h * w * f
    dec.append(1 * h * w * f)                                     # final Conv 16 -> 1
    image = 1 * H * W * f
    # ---- transformer trunk, fused per block (SURVEY's rule): a block reads its token matrix and writes it back,
    # qkv / scores / the MLP hidden stay on chip
    n_tok = gh * gw + 1
    trunk = depth * 2 * n_tok * tokens_dim * f
    fwd = 2 * (sum(stem) + sum(dec)) + skip + 2 * image + trunk
    bwd = 2 * fwd
    n_param = sum(int(__import__("math").prod(v)) for k, v in shapes.items()
                  if not k.endswith(("running_mean", "running_var", "num_batches_tracked")))
    per_step = n_param * (4 + 4 + 4 + 28)                         # weights fwd + bwd, wgrad write, clip + Adam
    return {"H": H, "W": W, "B": B, "stem_KB": [b // 1024 for b in stem], "decoder_KB": [b // 1024 for b in dec],
            "skip_KB": skip // 1024, "trunk_KB": trunk // 1024, "fwd_MB_per_sample": fwd / 1e6,
            "fwd_bwd_MB_per_sample": (fwd + bwd) / 1e6, "params": n_param, "per_step_MB": per_step / 1e6,
            "param_MB_per_sample": per_step / B / 1e6,
            "total_MB_per_sample": (fwd + bwd) / 1e6 + per_step / B / 1e6,
            "note": "SURVEY 8(d) quotes the same table in MiB-rounded form (fwd 26, fwd+bwd 77, +8.8 per sample at B=64); "
                    "bench.py keeps SURVEY's 77e6 + 8.8e6 B (the smaller, i.e. stricter, roofline denominator)"}


def main():
    a = [int(v) for v in sys.argv[1:4]]
    r = vessel_step_bytes(*a) if a else vessel_step_bytes()
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        hbm = json.load(open(pk))["hbm_gbs"]
        r["hbm_gbs"] = hbm
        r["roofline_samples_per_s_per_gpu"] = hbm * 1e9 / (r["total_MB_per_sample"] * 1e6)
    print(json.dumps(r, indent=1))


if __name__ == "__main__":
    main()
