"""Device-side input pipeline of the vessel trainer — mirrors what `VesselDataset.__getitem__` computes per
sample on DataLoader worker CPUs (`vessel_analysis/00_core/dataset.py:193-249`), for a whole batch on the GPU:

    Resize((H, W), antialias=True) -> flip by idx % 4 -> per-image min-max -> (img > img.mean()).float()
    one-hot treatment, StandardScaler-normalised morphology features

File discovery, CSV matching and TIFF decoding stay on the host (out of scope): the caller hands over
the decoded raw images as one [B, Hin, Win] fp32 CUDA tensor (pinned-host -> device copy on its own stream).
All arithmetic runs in libcvae_b200 (`cvae_vessel_preprocess`, `cvae_scaler_transform`, `cvae_one_hot`); there
is no CPU fallback.
"""
import ctypes as C

import torch

from .. import _lib as L
from .. import functional as F
from .config import CONFIG


class _Axis:
    """Index ranges + weights of one resize axis (computed on the device once per (in, out))."""
    def __init__(self, in_size, out_size, device):
        self.taps = L.lib.cvae_aa_max_interp(int(in_size), int(out_size))
        if self.taps <= 0:
            raise RuntimeError(f"libcvae_b200: bad resize axis {in_size} -> {out_size}")
        self.xmin = torch.empty(out_size, dtype=torch.int32, device=device)
        self.xsize = torch.empty(out_size, dtype=torch.int32, device=device)
        self.w = torch.empty(out_size, self.taps, dtype=torch.float32, device=device)
        L.check(L.lib.cvae_aa_weights(int(in_size), int(out_size), L.ptr(self.xmin), L.ptr(self.xsize), L.ptr(self.w),
                                      L.stream()), "aa_weights")


class VesselBatchTransform:
    """`transform(raw, aug_mode) -> x[B,1,H,W]` with the reference's per-sample arithmetic.

    img_h / img_w default to CONFIG["IMG_HEIGHT"] / CONFIG["IMG_WIDTH"] (dataset.py:182-183).  Workspaces are
    cached per batch size so a steady-state call allocates nothing (graph-capturable)."""

    def __init__(self, img_h=None, img_w=None, t_dim=None):
        self.img_h = int(CONFIG["IMG_HEIGHT"] if img_h is None else img_h)
        self.img_w = int(CONFIG["IMG_WIDTH"] if img_w is None else img_w)
        self.t_dim = int(CONFIG["T_DIM"] if t_dim is None else t_dim)
        self._axes = {}
        self._ws = {}
        self.mean_ = self.scale_ = None

    # -- images ------------------------------------------------------------------------------------------
    def _axis(self, n_in, n_out, device):
        key = (n_in, n_out, device)
        if key not in self._axes:
            self._axes[key] = _Axis(n_in, n_out, device)
        return self._axes[key]

    def transform(self, raw, aug_mode=None, out=None, return_threshold=False):
        """raw [B, Hin, Win] fp32 CUDA; aug_mode [B] int (idx % 4: 0 none, 1 hflip, 2 vflip, 3 both) or None
        (validation / test: no flips, dataset.py:197-198).  Returns the {0,1} image batch [B, 1, H, W]."""
        if raw.dim() == 4 and raw.shape[1] == 1:
            raw = raw[:, 0]
        if raw.dim() != 3 or raw.dtype != torch.float32 or not raw.is_cuda:
            raise RuntimeError("VesselBatchTransform expects a [B, Hin, Win] fp32 CUDA tensor (no CPU fallback)")
        raw = raw.contiguous()
        B, Hin, Win = raw.shape
        H, W, dev = self.img_h, self.img_w, raw.device
        ax, ay = self._axis(Win, W, dev), self._axis(Hin, H, dev)
        key = (B, dev)
        if key not in self._ws:
            self._ws[key] = (torch.empty(B, H, W, dtype=torch.float32, device=dev),
                             torch.empty(B, 4, dtype=torch.int32, device=dev),
                             torch.empty(B, dtype=torch.float32, device=dev))
        resized, stats, thr = self._ws[key]
        if out is None:
            out = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
        elif tuple(out.shape) != (B, 1, H, W) or not out.is_contiguous() or out.dtype != torch.float32:
            raise RuntimeError("out must be a contiguous fp32 [B, 1, H, W] tensor")
        if aug_mode is not None:
            aug_mode = aug_mode.to(device=dev, dtype=torch.int32).contiguous()
            if aug_mode.numel() != B:
                raise RuntimeError("aug_mode must hold one entry per image")
        p = L.PreprocParams(L.ptr(raw), L.ptr(resized), L.ptr(stats), L.ptr(out), L.ptr(thr), L.ptr(aug_mode),
                            L.ptr(ax.xmin), L.ptr(ax.xsize), L.ptr(ax.w), L.ptr(ay.xmin), L.ptr(ay.xsize), L.ptr(ay.w),
                            B, Hin, Win, H, W)
        L.check(L.lib.cvae_vessel_preprocess(C.byref(p), L.stream()), "vessel_preprocess")
        L.launch_count += 2          # three kernels per call
        return (out, thr.clone()) if return_threshold else out

    __call__ = transform

    # -- treatment / features ----------------------------------------------------------------------------
    def one_hot(self, t_idx):
        """dataset.py:243-245."""
        return F.one_hot(t_idx.to(torch.int64), self.t_dim)

    def fit_scaler(self, m_all):
        """StandardScaler().fit over ALL rows (dataset.py:113-115): population variance, zero scale -> 1.
        One-off at dataset construction; fp64 on the device."""
        m = torch.as_tensor(m_all, dtype=torch.float64)
        if not m.is_cuda:
            m = m.cuda()
        self.mean_ = m.mean(0)
        var = ((m - self.mean_) ** 2).mean(0)
        scale = var.sqrt()
        self.scale_ = torch.where(scale < 10 * torch.finfo(torch.float64).eps, torch.ones_like(scale), scale)
        return self

    def set_scaler(self, mean, scale, device="cuda"):
        self.mean_ = torch.as_tensor(mean, dtype=torch.float64).to(device).contiguous()
        self.scale_ = torch.as_tensor(scale, dtype=torch.float64).to(device).contiguous()
        return self

    def transform_m(self, m):
        """(m - mean_) / scale_ in fp64, stored fp32 (dataset.py:116,240)."""
        if self.mean_ is None:
            raise RuntimeError("fit_scaler / set_scaler first")
        m = m.to(device=self.mean_.device, dtype=torch.float64).contiguous()
        out = torch.empty(m.shape, dtype=torch.float32, device=m.device)
        L.check(L.lib.cvae_scaler_transform(L.ptr(m), L.ptr(self.mean_), L.ptr(self.scale_), L.ptr(out),
                                            m.shape[0], m.shape[1], L.stream()), "scaler_transform")
        return out
