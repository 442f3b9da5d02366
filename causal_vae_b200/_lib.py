"""ctypes binding of libcvae_b200.so (the C ABI declared in include/cvae_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every
entry point raises RuntimeError on a non-zero status.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CVAE_LIB: a differently-built copy of the SAME library (scripts/: the -DCVAE_TIMING role-timer build); never a fallback
LIB_PATH = os.environ.get("CVAE_LIB") or os.path.join(_HERE, "libcvae_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). causal_vae_b200 has no CPU / ATen fallback.")

lib = C.CDLL(LIB_PATH)

vp, f32, i32, i64, u64, f64 = C.c_void_p, C.c_float, C.c_int, C.c_int64, C.c_uint64, C.c_double


class Xform(C.Structure):
    _fields_ = [("scale", vp), ("shift", vp), ("center", vp), ("slope", f32)]


class ConvParams(C.Structure):
    _fields_ = [("src", vp), ("wt", vp), ("bias", vp), ("dst", vp), ("in_", Xform), ("epi", i32),
                ("epi_ref", vp), ("epi_add", vp), ("epi_x", Xform), ("stats", vp),
                ("N", i32), ("Hs", i32), ("Ws", i32), ("Cs", i32), ("Hd", i32), ("Wd", i32), ("Cd", i32),
                ("kh", i32), ("kw", i32), ("stride", i32), ("pad", i32), ("mode", i32)]


class WgradParams(C.Structure):
    _fields_ = [("ga", vp), ("db", vp), ("xa", Xform), ("xb", Xform), ("partial", vp), ("splits", i32),
                ("N", i32), ("Ha", i32), ("Wa", i32), ("Ca", i32), ("Hq", i32), ("Wq", i32), ("Cb", i32),
                ("kh", i32), ("kw", i32), ("stride", i32), ("pad", i32)]


class PreprocParams(C.Structure):
    _fields_ = [("raw", vp), ("resized", vp), ("stats", vp), ("mask", vp), ("thr", vp), ("aug_mode", vp),
                ("xmin", vp), ("xsize", vp), ("xw", vp), ("ymin", vp), ("ysize", vp), ("yw", vp),
                ("B", i32), ("Hin", i32), ("Win", i32), ("H", i32), ("W", i32)]


EPI_PLAIN, EPI_STATS, EPI_DACT = 0, 1, 2
MODE_GATHER, MODE_SCATTER = 0, 1
ACT_LRELU, ACT_GELU, ACT_SIGMOID = 0, 1, 2

_SIGS = {
    "cvae_version": [],
    "cvae_built_arch": [],
    "cvae_conv_gather": [C.POINTER(ConvParams), vp],
    "cvae_wgrad_splits": [i32, i32, i32],
    "cvae_conv_wgrad": [C.POINTER(WgradParams), vp],
    "cvae_tc_pack_rows_floats": [i64, i32],
    "cvae_tc_pack_rows": [vp, vp, i64, i32, vp],
    "cvae_linear_tc_packed": [C.POINTER(ConvParams), vp, vp],
    "cvae_wgrad_reduce": [vp, i32, i32, i32, i32, i32, vp, i32, vp],
    "cvae_pack_weight": [vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "cvae_conv_few_eligible": [i32] * 12,
    "cvae_head_bwd_eligible": [i32, i32, i32, i32],
    "cvae_head_bwd": [vp, vp, Xform, vp, vp, vp, vp, i32, i32, i32, i32, vp],
    "cvae_pack_batch_blocks": [i32, i32, i32, i32],
    "cvae_pack_batch": [vp, i32, i32, vp],
    "cvae_tc_eligible": [i32, i32, i64],
    "cvae_tc_pack_floats": [i32, i32, i32],
    "cvae_tc_pack_weight": [vp, vp, i32, i32, i32, i32, i32, i32, vp],
    "cvae_conv_gather_tc": [C.POINTER(ConvParams), vp],
    "cvae_wgrad_tc_eligible": [i32, i32, i32],
    "cvae_wgrad_tc_splits": [i32, i32, i32],
    "cvae_conv_wgrad_tc": [C.POINTER(WgradParams), vp],
    "cvae_conv_wgrad_tc_direct": [C.POINTER(WgradParams), vp, i32, vp],
    "cvae_wgrad_tile_splits": [i32, i32, i32, i32, i32, i32],
    "cvae_conv_wgrad_tile": [C.POINTER(WgradParams), vp],
    "cvae_bn_finalize": [vp, i32, f64, vp, vp, f32, f32, vp, vp, vp, vp, vp, vp, vp, vp],
    "cvae_bn_eval_coeffs": [vp, vp, vp, vp, f32, i32, vp, vp, vp],
    "cvae_col_stats": [vp, i64, i32, vp, vp],
    "cvae_bn_bwd_finalize": [vp, i32, f64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "cvae_affine_act": [vp, Xform, vp, Xform, vp, i64, i32, vp],
    "cvae_bn_bwd_apply": [vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
    "cvae_bn_bwd_fused_ok": [i32],
    "cvae_bn_bwd": [vp, vp, vp, C.c_double, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
    "cvae_dact_stats": [vp, vp, Xform, vp, vp, i64, i32, vp],
    "cvae_layernorm_fwd": [vp, vp, vp, vp, vp, vp, i64, i32, i64, f32, vp],
    "cvae_layernorm_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, i64, i64, i32, vp],
    "cvae_add_layernorm_fwd": [vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, f32, vp],
    "cvae_layernorm_bwd_add": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i32, vp],
    "cvae_attention_fwd": [vp, vp, vp, i32, i32, i32, i32, f32, u64, u64, vp, vp],
    "cvae_attention_bwd": [vp, vp, vp, vp, i32, i32, i32, i32, f32, u64, u64, vp, vp],
    "cvae_attention_ws_bytes": [i32, i32, i32, i32],
    "cvae_attention_bwd_ws": [vp, vp, vp, vp, vp, i64, i32, i32, i32, i32, f32, vp],
    "cvae_act_fwd": [vp, vp, i64, i32, f32, vp],
    "cvae_act_bwd": [vp, vp, vp, i64, i32, f32, vp],
    "cvae_add": [vp, vp, vp, i64, vp],
    "cvae_dropout": [vp, vp, i64, f32, u64, u64, vp, vp],
    "cvae_dropout_fused": [vp, vp, vp, i64, i32, i32, f32, f32, u64, u64, vp, vp],
    "cvae_clamp_fwd": [vp, vp, i64, f32, f32, vp],
    "cvae_clamp_bwd": [vp, vp, vp, i64, f32, f32, vp],
    "cvae_kld_fwd": [vp, vp, i64, vp, vp],
    "cvae_tokens_fwd": [vp, vp, vp, vp, i32, i32, i32, vp],
    "cvae_tokens_bwd": [vp, vp, vp, vp, i32, i32, i32, vp],
    "cvae_transpose_bc": [vp, vp, i32, i32, i32, vp],
    "cvae_copy_cols": [vp, i64, i32, vp, i64, i32, i64, i32, i32, vp],
    "cvae_fill": [vp, i64, f32, vp],
    "cvae_col_sum": [vp, i64, i32, vp, i32, vp],
    "cvae_latent_fwd": [vp, vp, vp, vp, vp, vp, i32, i32, f32, f32, vp],
    "cvae_latent_bwd": [vp, vp, vp, vp, vp, vp, i32, i32, f32, f32, vp],
    "cvae_gauss_nll_fwd": [vp, vp, vp, vp, vp, i64, f32, vp],
    "cvae_gauss_nll_bwd": [vp, vp, vp, vp, f32, vp, vp, i64, f32, vp],
    "cvae_kld_bwd": [vp, vp, vp, f32, vp, vp, i64, i32, vp],
    "cvae_vessel_xsum": [vp, i64, vp, vp],
    "cvae_vessel_recon_fwd": [vp, vp, i64, vp, vp],
    "cvae_vessel_recon_bwd": [vp, vp, i64, vp, vp, vp, vp, vp],
    "cvae_mse_fwd": [vp, vp, i64, vp, vp],
    "cvae_mse_bwd": [vp, vp, i64, vp, f32, vp, vp],
    "cvae_bce_fwd": [vp, vp, i64, vp, vp],
    "cvae_bce_bwd": [vp, vp, i64, vp, f32, vp, vp],
    "cvae_finish_scalar": [vp, f32, vp, vp],
    "cvae_scalar_combine": [vp, vp, vp, vp, f32, f32, f32, f32, vp, vp],
    "cvae_scalar_scale4": [vp, f32, f32, f32, f32, vp, vp],
    "cvae_argmax_rows": [vp, i64, i32, vp, vp],
    "cvae_one_hot": [vp, i64, i32, vp, vp],
    "cvae_softmax_ce_fwd": [vp, vp, i64, i32, vp, vp],
    "cvae_softmax_ce_bwd": [vp, vp, i64, i32, vp, f32, vp, vp],
    "cvae_uniform_kl_fwd": [vp, i64, i32, vp, vp],
    "cvae_uniform_kl_bwd": [vp, i64, i32, vp, f32, vp, vp],
    "cvae_debug_read": [vp, i32],
    "cvae_do_expand": [vp, vp, vp, i32, i32, i32, i32, f32, vp],
    "cvae_rowdiff_l2": [vp, vp, vp, i64, i64, i32, vp],
    "cvae_pair_expand": [vp, vp, vp, i32, i32, i32, f32, vp],
    "cvae_ensemble_mean_std": [vp, i32, vp, vp, i64, vp],
    "cvae_sumsq": [vp, i64, vp, vp],
    "cvae_clip_adam": [vp, vp, vp, vp, i64, vp, f32, f32, f32, f32, f32, f32, vp, vp],
    "cvae_upsample2x_fwd": [vp, vp, i32, i32, i32, i32, vp],
    "cvae_upsample2x_bwd": [vp, vp, i32, i32, i32, i32, vp],
    "cvae_aa_max_interp": [i32, i32],
    "cvae_aa_weights": [i32, i32, vp, vp, vp, vp],
    "cvae_vessel_preprocess": [C.POINTER(PreprocParams), vp],
    "cvae_scaler_transform": [vp, vp, vp, vp, i64, i32, vp],
    "cvae_graph_set_priorities": [vp, i32, i32, vp],
    "cvae_graph_instantiate_prio": [vp, vp],
    "cvae_graph_launch": [vp, vp],
    "cvae_graph_exec_destroy": [vp],
}
EXPORTS = tuple(_SIGS)

for _name, _args in _SIGS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header / library drift: fail loudly
    _fn.argtypes = _args
    _fn.restype = C.c_int64 if _name in ("cvae_tc_pack_floats", "cvae_attention_ws_bytes", "cvae_tc_pack_rows_floats") else C.c_int

_ERR = {-1: "bad argument", -2: "unsupported shape", -3: "alignment", -4: "CUDA launch error"}

# number of native kernel-launching calls issued through this binding (bench.py's `gpu_launches`)
launch_count = 0


def check(rc, what):
    global launch_count
    launch_count += 1
    if rc != 0:
        raise RuntimeError(f"libcvae_b200: {what} failed: {_ERR.get(rc, rc)}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  The tensor must be a CUDA fp32/fp64/int64 tensor."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libcvae_b200 runs on CUDA tensors only (no CPU fallback)")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def xform(scale=None, shift=None, slope=1.0, center=None):
    return Xform(ptr(scale), ptr(shift), ptr(center), float(slope))


class PriorityGraphExec:
    """Executable form of a captured torch.cuda.CUDAGraph (keep_graph=True) whose kernel nodes carry launch priorities:
    the weight-gradient family yields SMs to the main chain (csrc/graph_prio.cu).  launch() replays on the current stream."""

    def __init__(self, graph, prio_main=-1, prio_side=0):
        self.graph = graph                        # keeps the cudaGraph_t and its memory pool alive
        counts = (C.c_int * 3)()
        raw = graph.raw_cuda_graph()
        rc = lib.cvae_graph_set_priorities(raw, prio_main, prio_side, C.cast(counts, C.c_void_p))
        if rc != 0:
            raise RuntimeError(f"libcvae_b200: graph_set_priorities failed: {_ERR.get(rc, rc)}")
        self.kernel_nodes, self.side_nodes, self.unnamed = counts[0], counts[1], counts[2]
        ex = C.c_void_p()
        rc = lib.cvae_graph_instantiate_prio(raw, C.byref(ex))
        if rc != 0:
            raise RuntimeError(f"libcvae_b200: graph_instantiate_prio failed: {_ERR.get(rc, rc)}")
        self.exec = ex

    def launch(self):
        rc = lib.cvae_graph_launch(self.exec, stream())
        if rc != 0:
            raise RuntimeError(f"libcvae_b200: graph_launch failed: {_ERR.get(rc, rc)}")

    def __del__(self):
        ex, self.exec = getattr(self, "exec", None), None
        if ex:
            try:
                lib.cvae_graph_exec_destroy(ex)
            except Exception:
                pass
