"""Shared point-wise MLP (+ max-pool) host side -- SURVEY.md section 8(a) row a6.

Inference path: ONE launch of the hand-written tcgen05 kernel (csrc/mlp.cu) per SA / FP /
voting stage: neighbourhood gather -> 2-3 layer MLP on the tensor cores -> max-pool, with
the grouped tensor never materialised.  Features travel between stages channel-last in
bf16 ("cl"); the public, lineage-shaped (B,C,N) fp32 tensors ("cf") are still produced for
the caller, and carry their cl twin in the `_sad_cl` attribute so the next stage can skip
the layout bridge.  Storage precision: bf16 operands, fp32 accumulate / bias / ReLU / max
(2e-2 vs the fp32 oracle, BASELINE north_star).

Shapes the kernels do not cover (hidden widths not a multiple of 64 or > 256, more than 3
layers, nsample not a power of two) are rejected with SAD_EUNSUPPORTED (`UnsupportedShape`):
there is no library-matmul fallback in this package.  Training composes the same stage from
the autograd operators of `ops.py` and `torch.nn` layers (modules.SharedMLP).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops

_VP = ctypes.c_void_p


class UnsupportedShape(_lib.SadLibraryError):
    """SAD_EUNSUPPORTED (-3) raised on the host side: the fused kernels are not built for this MLP shape."""
    code = -3


def _reject(what: str, mlp: "PreparedMLP", S: int):
    raise UnsupportedShape(
        f"{what}: SAD_EUNSUPPORTED -- fused MLP kernels take 2-3 layers, hidden widths % 64 (tf32: 32) == 0 and <= 256, last width "
        f"<= 512, nsample a power of two <= 128; got widths {list(mlp.c_out)}, nsample {S}")

# Hand 128-row tiles to the persistent CTAs through an atomic counter (robust when other streams hold SMs).
DYNAMIC_TILES = True

# Shape-specialised SA kernel (csrc/mlp_sa.cu): True = use it whenever an instance matches the stage, "single" / "pair"
# = prefer the instances that run on single CTAs / CTA pairs, False = always the general kernel (tests compare them).
# Duplicate-free SA stages (sad_sa_mlp_dedup_fwd): a ball query pads short neighbourhoods with copies of the first hit and
# the max-pool ignores copies, so only each point's leading samples go through the MLP.  Bit-identical results; applies
# to the single-CTA instances with nsample 32 / 64 (SA1, SA2).
DEDUP_SA = [True]
# nsample values it is applied to: at nsample 64 (SA1: 30 of 64 samples distinct on average, 63 % of the rows remain) the
# 16-sample instance's higher cost per tile eats the saving; at nsample 32 (SA2: 7 of 32 distinct, half the rows remain)
# it pays.  The C entry point accepts both.
DEDUP_NSAMPLE = (32,)
DEDUP_SLOT = [0]      # samples per slot: 0 = the smallest the library has for the stage (tools / tests: 8 or 16)
FAST_SA = [True]
# Scheduling hint for the fused-MLP launches issued from Python (never changes results): minimum 128-row tiles per
# CTA.  engine.PipelinedHotPath raises it while it captures its graphs (narrower grids for the small stages).
TILES_PER_CTA = [1]
# general kernel only: 0 = automatic, 1 = never, 2 = always (where the shape allows) two tiles per context and phase
SUPER_TILES = [0]
# tools: skip the (B,C,P) f32 output of the specialised SA stages (only the channel-last bf16 twin is written)
_WANT_CF = [True]


class _SchedWords:
    """Device words the persistent kernels' tile schedulers count in.  A kernel leaves its words zero when it
    finishes, so one zero-initialised pool serves every launch: eager launches share one slot per stream (they are
    stream-ordered), every launch recorded into a CUDA graph gets a slot of its own (graphs of different pipeline
    slots replay concurrently).  No per-launch memset."""

    STRIDE = 64          # int32 words between slots (256 B: separate L2 atomic units)

    def __init__(self):
        self.pools = {}

    def take(self, dev: torch.device) -> torch.Tensor:
        key = (dev.type, dev.index)
        st = self.pools.get(key)
        capturing = torch.cuda.is_current_stream_capturing()
        if st is None or st["next"] + self.STRIDE > st["buf"].numel():
            st = self.pools[key] = {"buf": torch.zeros(self.STRIDE * 4096, dtype=torch.int32, device=dev), "next": 0,
                                    "by_stream": {}}
        if capturing:
            off = st["next"]
            st["next"] += self.STRIDE
        else:
            sid = torch.cuda.current_stream(dev).cuda_stream
            off = st["by_stream"].get(sid)
            if off is None:
                off = st["by_stream"][sid] = st["next"]
                st["next"] += self.STRIDE
        return st["buf"][off: off + 2]


_SCHED = _SchedWords()


def _ptr(t: Optional[torch.Tensor]):
    return _VP(t.data_ptr()) if t is not None else _VP(0)


def _stream(t: torch.Tensor):
    return _VP(torch.cuda.current_stream(t.device).cuda_stream)


def _round64(c: int) -> int:
    return (c + 63) // 64 * 64


class Layout:
    """Where the columns of the first layer's weight matrix come from (K order of the kernel):
    [feat_cl (c0) | feat2_cl (c1) | special chunk: xyz(3), extras (e)]."""

    def __init__(self, c0=0, c0_cols=(), c1=0, c1_cols=(), xyz_cols=None, extra_cols=()):
        self.c0, self.c1 = c0, c1                       # padded widths (multiples of 64)
        self.c0_cols, self.c1_cols = list(c0_cols), list(c1_cols)
        self.xyz_cols = list(xyz_cols) if xyz_cols is not None else None
        self.extra_cols = list(extra_cols)

    @property
    def has_special(self):
        return self.xyz_cols is not None or len(self.extra_cols) > 0

    def perm(self):
        p = np.full(self.c0 + self.c1 + (64 if self.has_special else 0), -1, dtype=np.int32)
        p[: len(self.c0_cols)] = self.c0_cols
        p[self.c0: self.c0 + len(self.c1_cols)] = self.c1_cols
        if self.has_special:
            s = self.c0 + self.c1
            if self.xyz_cols is not None:
                p[s: s + 3] = self.xyz_cols
            p[s + 3: s + 3 + len(self.extra_cols)] = self.extra_cols
        return p

    def key(self):
        return (self.c0, tuple(self.c0_cols), self.c1, tuple(self.c1_cols),
                tuple(self.xyz_cols) if self.xyz_cols is not None else None, tuple(self.extra_cols))


def sa_layout(c_feat: int, use_xyz: bool) -> Layout:
    """SA stage: original columns are [xyz(3) if use_xyz, features(c_feat)]."""
    off = 3 if use_xyz else 0
    xyz = [0, 1, 2] if use_xyz else None
    if 0 < c_feat <= 13:
        return Layout(xyz_cols=xyz, extra_cols=range(off, off + c_feat))
    return Layout(c0=_round64(c_feat), c0_cols=range(off, off + c_feat), xyz_cols=xyz)


def fp_layout(c_interp: int, c_skip: int) -> Layout:
    return Layout(c0=_round64(c_interp), c0_cols=range(c_interp),
                  c1=_round64(c_skip), c1_cols=range(c_interp, c_interp + c_skip))


class PreparedMLP:
    """BN-folded layers [(W (Cout,Cin) f32, b (Cout,) f32), ...] plus, lazily per layout, the
    packed bf16 weight images the tcgen05 kernel streams."""

    def __init__(self, layers: Sequence[Tuple[torch.Tensor, torch.Tensor]], dtype: str = "bf16"):
        if dtype not in ("bf16", "tf32"):
            raise ValueError(f"PreparedMLP dtype must be 'bf16' or 'tf32', got {dtype!r}")
        self.dtype = dtype            # operand precision of the fused kernels: bf16 (2e-2 bar) or tf32 (fp32 activations)
        self.layers = [(W.contiguous(), b.contiguous()) for (W, b) in layers]
        self.c_out = [int(W.shape[0]) for (W, _) in self.layers]
        self.c_in = int(self.layers[0][0].shape[1])
        self.W_bf16 = [W.to(torch.bfloat16) for (W, _) in self.layers]
        self._packed = {}

    def __len__(self):
        return len(self.layers)

    def __iter__(self):
        return iter(self.layers)

    def fusable(self, S: int) -> bool:
        n = len(self.layers)
        step = 32 if self.dtype == "tf32" else 64
        hidden_ok = all(c % step == 0 and c <= 256 for c in self.c_out[:-1])
        return 2 <= n <= 3 and hidden_ok and self.c_out[-1] <= 512 and S in (1, 2, 4, 8, 16, 32, 64, 128)

    def packed(self, layout: Layout, S: int = 1):
        """Weight images for this K layout.  The last layer's image depends on how the kernel evaluates
        it: transposed (mode 1) for pooled stages (S > 1), plain (mode 2) for S == 1."""
        last_mode = 1 if S > 1 else 2
        key = (layout.key(), last_mode)
        if key not in self._packed:
            lib = _lib.load()
            dev = self.layers[0][0].device
            imgs, biases = [], []
            n = len(self.layers)
            for li, (W, b) in enumerate(self.layers):
                Wn = np.ascontiguousarray(W.detach().cpu().numpy(), dtype=np.float32)
                cout, cin = Wn.shape
                perm = layout.perm() if li == 0 else np.arange(_round64(cin), dtype=np.int32)
                if li > 0:
                    perm[cin:] = -1
                kpad = int(perm.shape[0])
                is_last = last_mode if li == n - 1 else 0
                nbytes = lib.sad_mlp_weight_image_bytes(cout, kpad, is_last)
                if nbytes <= 0:
                    raise RuntimeError("sad_mlp_weight_image_bytes rejected the layer shape")
                img = np.zeros(nbytes, dtype=np.uint8)
                _lib.check(lib.sad_mlp_pack_weights(_VP(Wn.ctypes.data), cout, cin, _VP(perm.ctypes.data), kpad,
                                                    is_last, _VP(img.ctypes.data)), "mlp_pack_weights")
                imgs.append(torch.from_numpy(img).to(dev))
                biases.append(b.detach().float().contiguous().to(dev))
            self._packed[key] = (imgs, biases,
                                 (_VP * n)(*[t.data_ptr() for t in imgs]),
                                 (_VP * n)(*[t.data_ptr() for t in biases]),
                                 (ctypes.c_int * n)(*self.c_out))
        return self._packed[key]


_DEDUP_OK = {}


def _dedup_ok(inst: int) -> bool:
    """The instance is a single-CTA one with nsample 32 / 64 (the duplicate-free launch has a 16-sample sibling for it)."""
    if inst not in _DEDUP_OK:
        info = (ctypes.c_int * 5)()
        _lib.check(_lib.load().sad_sa_mlp_instance_info(int(inst), info), "sa_mlp_instance_info")
        cg, _, _, _, s_ = list(info)
        _DEDUP_OK[inst] = cg == 1 and s_ in DEDUP_NSAMPLE
    return _DEDUP_OK[inst]


def _fast_instance(mlp: "PreparedMLP", layout: Layout, S: int, P: int) -> int:
    """Instance id of the shape-specialised kernel for this stage, or -1 (general kernel)."""
    if not FAST_SA[0] or len(mlp) != 3 or layout.c1 or layout.xyz_cols is None or P & (P - 1) or P < 1:
        return -1
    h1, h2, c3 = mlp.c_out
    return int(_lib.load().sad_sa_mlp_query(layout.c0, h1, h2, c3, S, len(layout.extra_cols), 1,
                                            {"single": 1, "pair": 2}.get(FAST_SA[0], 0)))


def _packed_fast(mlp: "PreparedMLP", inst: int, layout: Layout):
    key = ("fast", inst, layout.key())
    if key not in mlp._packed:
        lib = _lib.load()
        dev = mlp.layers[0][0].device
        (W1, b1), (W2, b2), (W3, b3) = [(np.ascontiguousarray(W.detach().cpu().numpy(), dtype=np.float32),
                                         np.ascontiguousarray(b.detach().cpu().numpy(), dtype=np.float32))
                                        for (W, b) in mlp.layers]
        c3 = W3.shape[0]
        perm_feat = np.ascontiguousarray(layout.perm()[: layout.c0], dtype=np.int32) if layout.c0 else np.zeros(1, np.int32)
        perm_sp = np.full(7, -1, dtype=np.int32)
        perm_sp[:3] = layout.xyz_cols
        perm_sp[3: 3 + len(layout.extra_cols)] = layout.extra_cols
        nbytes = int(lib.sad_sa_mlp_image_bytes(inst))
        img = np.zeros(nbytes, dtype=np.uint8)
        _lib.check(lib.sad_sa_mlp_pack(inst, _VP(W1.ctypes.data), int(W1.shape[1]), _VP(perm_feat.ctypes.data),
                                       _VP(perm_sp.ctypes.data), _VP(b1.ctypes.data), _VP(W2.ctypes.data),
                                       _VP(b2.ctypes.data), _VP(W3.ctypes.data), int(c3), _VP(img.ctypes.data)),
                   "sa_mlp_pack")
        b3p = np.zeros(256, dtype=np.float32)
        b3p[:c3] = b3
        mlp._packed[key] = (torch.from_numpy(img).to(dev), torch.from_numpy(b3p).to(dev))
    return mlp._packed[key]


def fused_sa_fast(mlp: "PreparedMLP", inst: int, layout: Layout, B, N, P, feat_cl, xyz, new_xyz, idx, radius, radius_t,
                  normalize_xyz, extra, want_cf=True, want_cl=True):
    """Launcher of sad_sa_mlp_fwd -> (out_cf (B,C3,P) f32 | None, out_cl (B,P,C3) bf16 | None)."""
    img, b3p = _packed_fast(mlp, inst, layout)
    dev = img.device
    c3 = mlp.c_out[-1]
    out_cf = torch.empty((B, c3, P), dtype=torch.float32, device=dev) if want_cf else None
    out_cl = torch.empty((B, P, c3), dtype=torch.bfloat16, device=dev) if want_cl else None
    sched = _SCHED.take(dev)
    E = len(layout.extra_cols)
    xyzw = None
    with torch.cuda.device(dev):
        if E == 0:
            xyzw = getattr(xyz, "_sad_xyzw", None)      # written next to new_xyz by ops.gather_points
            if xyzw is not None and tuple(xyzw.shape) != (B, N, 4):
                xyzw = None
        if E == 1:
            twin = getattr(xyz, "_sad_xyzw1", None)     # prepack_xyzw ran earlier for this (xyz, feature) pair
            if twin is not None and twin[0] == extra.data_ptr() and twin[1] == extra._version and tuple(twin[2].shape) == (B, N, 4):
                xyzw = twin[2]
        if xyzw is None and E <= 1:      # gathered source of the special K step as one 16-byte row per point
            xyzw = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
            _lib.check(_lib.load().sad_pack_xyzw(B, N, _ptr(xyz), _ptr(extra if E else None), _ptr(xyzw), _stream(xyzw)),
                       "pack_xyzw")
        lib = _lib.load()
        common = (inst, B, N, P, _ptr(feat_cl), _ptr(xyz), _ptr(xyzw), _ptr(new_xyz), _ptr(idx), float(radius), _ptr(radius_t),
                  int(bool(normalize_xyz)), _ptr(extra), len(layout.extra_cols), _ptr(img), _ptr(b3p), c3,
                  _ptr(out_cl), _ptr(out_cf), _ptr(sched))
        stream = _VP(torch.cuda.current_stream(dev).cuda_stream)
        if DEDUP_SA[0] and _dedup_ok(inst):
            ws = torch.empty(int(lib.sad_sa_mlp_dedup_workspace_bytes(B, P)), dtype=torch.uint8, device=dev)
            rc = lib.sad_sa_mlp_dedup_fwd(*common, _ptr(ws), int(DEDUP_SLOT[0]), int(TILES_PER_CTA[0]), stream)
        else:
            rc = lib.sad_sa_mlp_fwd(*common, int(TILES_PER_CTA[0]), stream)
    _lib.check(rc, "sa_mlp")
    return out_cf, out_cl


def prepack_xyzw(xyz: torch.Tensor, extra: torch.Tensor) -> torch.Tensor:
    """(B,N,3) coordinates + ONE scalar feature per point ((B,1,N) or (B,N,1)) -> (B,N,4) {x,y,z,f}: the gathered
    source of a fused SA stage whose only feature is a scalar (SA1: height).  The result rides on `xyz` so that the
    stage finds it; a caller with a stream to spare runs this early, off the critical path (Pointnet2Backbone does,
    under the sampling chain)."""
    B, N, _ = xyz.shape
    out = torch.empty((B, N, 4), dtype=torch.float32, device=xyz.device)
    e = extra.contiguous()
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().sad_pack_xyzw(B, N, _ptr(xyz), _ptr(e), _ptr(out), _stream(out)), "pack_xyzw")
    xyz._sad_xyzw1 = (extra.data_ptr(), extra._version, out)
    return out


def prepare_layers(layers, dtype: str = "bf16") -> PreparedMLP:
    return PreparedMLP(layers, dtype)


# ----------------------------------------------------------------------------- tf32 mode (csrc/mlp_tf32.cu)
def to_cl_f32(features_cf: torch.Tensor) -> torch.Tensor:
    """(B,C,N) f32 -> (B,N,C4) f32 channel-last, C padded to a multiple of 4; the twin a tf32 stage left on its
    output is used when present."""
    twin = getattr(features_cf, "_sad_cl32", None)
    if twin is not None and getattr(features_cf, "_sad_cl_v", None) == features_cf._version:
        return twin
    B, C, N = features_cf.shape
    c4 = (C + 3) // 4 * 4
    if c4 == C:
        return features_cf.transpose(1, 2).contiguous()
    out = torch.zeros((B, N, c4), dtype=torch.float32, device=features_cf.device)
    out[:, :, :C] = features_cf.transpose(1, 2)
    return out


def _packed_tf32(mlp: PreparedMLP, ci: int, ci_real: int, cf: int, cf_real: int, special: bool, E: int, order: str):
    """Per-layer weight images + biases for sad_mlp_tf32_fwd.  Layer-1 operand order is [interp | feat | special];
    `order` says where those parts sit in the lineage weight matrix: "sa" = [xyz(3) | features] (features = the extras
    when E > 0), "fp" = [interp | skip], "pw" = [features]."""
    key = ("tf32", ci, ci_real, cf, cf_real, special, E, order)
    if key not in mlp._packed:
        lib = _lib.load()
        dev = mlp.layers[0][0].device
        n = len(mlp.layers)
        kc_i, kc_f = (ci + 31) // 32, (cf + 31) // 32
        kc1 = kc_i + kc_f + (1 if special else 0)
        imgs, biases = [], []
        for li, (W, b) in enumerate(mlp.layers):
            Wn = np.ascontiguousarray(W.detach().cpu().numpy(), dtype=np.float32)
            cout, cin = Wn.shape
            last = int(li == n - 1)
            if li == 0:
                kc = kc1
                Wp = np.zeros((cout, kc * 32), dtype=np.float32)
                if order == "sa":
                    x0 = 3 if special else 0
                    if E:
                        Wp[:, (kc_i + kc_f) * 32 + 3:(kc_i + kc_f) * 32 + 3 + E] = Wn[:, x0:x0 + E]
                    else:
                        Wp[:, kc_i * 32:kc_i * 32 + cf_real] = Wn[:, x0:x0 + cf_real]
                    if special:
                        Wp[:, (kc_i + kc_f) * 32:(kc_i + kc_f) * 32 + 3] = Wn[:, :3]
                else:
                    Wp[:, :ci_real] = Wn[:, :ci_real]
                    Wp[:, kc_i * 32:kc_i * 32 + cf_real] = Wn[:, ci_real:ci_real + cf_real]
            else:
                kc = cin // 32
                Wp = Wn
            nbytes = lib.sad_mlp_tf32_image_bytes(cout, kc, last)
            img = np.zeros(nbytes, dtype=np.uint8)
            _lib.check(lib.sad_mlp_tf32_pack(_VP(np.ascontiguousarray(Wp).ctypes.data), cout, kc, last, _VP(img.ctypes.data)),
                       "mlp_tf32_pack")
            imgs.append(torch.from_numpy(img).to(dev))
            bb = b.detach().float().cpu().numpy()
            if last:
                pad = np.zeros((cout + 127) // 128 * 128, dtype=np.float32)
                pad[:cout] = bb
                bb = pad
            biases.append(torch.from_numpy(np.ascontiguousarray(bb)).to(dev))
        mlp._packed[key] = (imgs, biases, (_VP * n)(*[t.data_ptr() for t in imgs]),
                            (_VP * n)(*[t.data_ptr() for t in biases]), (ctypes.c_int * n)(*mlp.c_out))
    return mlp._packed[key]


def _attach32(cf: torch.Tensor, cl: torch.Tensor):
    cf._sad_cl32 = cl
    cf._sad_cl_v = cf._version
    return cf


def _tf32_launch(mlp, pk, B, N, P, S, known_cl, m, CI, nn_idx, nn_w, feat_cl, CF, idx, xyz, new_xyz, radius, radius_t,
                 normalize, extra, E, last_relu):
    _, _, wp, bp, cp = pk
    dev = (feat_cl if feat_cl is not None else known_cl if known_cl is not None else xyz).device
    c_last = mlp.c_out[-1]
    out_cf = torch.empty((B, c_last, P), dtype=torch.float32, device=dev)
    out_cl = torch.empty((B, P, c_last), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().sad_mlp_tf32_fwd(B, N, P, S, _ptr(known_cl), int(m), int(CI), _ptr(nn_idx), _ptr(nn_w), _ptr(feat_cl),
                                          int(CF), _ptr(idx), _ptr(xyz), _ptr(new_xyz), float(radius), _ptr(radius_t),
                                          int(bool(normalize)), _ptr(extra), int(E), len(mlp), wp, bp, cp, int(bool(last_relu)),
                                          _ptr(out_cf), _ptr(out_cl), _VP(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "mlp_tf32")
    return _attach32(out_cf, out_cl)


def sa_group_mlp_tf32(xyz, new_xyz, features, idx, radius, mlp: PreparedMLP, use_xyz=True, normalize_xyz=True):
    B, P, S = idx.shape
    N = xyz.shape[1]
    c_feat = 0 if features is None else features.shape[1]
    special = bool(use_xyz or c_feat == 0)
    E, feat_cl, extra, CF = 0, None, None, 0
    if special and 0 < c_feat <= 4:
        E = c_feat                                  # few scalar features ride in the special K step (SA1: height)
        extra = features.contiguous() if c_feat == 1 else features.transpose(1, 2).contiguous()
    elif c_feat:
        feat_cl = to_cl_f32(features)
        CF = feat_cl.shape[2]
    radius_t = radius if torch.is_tensor(radius) else None
    pk = _packed_tf32(mlp, 0, 0, CF, 0 if E else c_feat, special, E, "sa")
    return _tf32_launch(mlp, pk, B, N, P, S, None, 0, 0, None, None, feat_cl, CF, idx, xyz if special else None, new_xyz,
                        0.0 if radius_t is not None else float(radius), radius_t, normalize_xyz, extra, E, True)


def fp_interp_mlp_tf32(known_feats, unknow_feats, idx, weight, mlp: PreparedMLP):
    B, C2, m = known_feats.shape
    n = idx.shape[1]
    known_cl = to_cl_f32(known_feats)
    skip_cl = to_cl_f32(unknow_feats) if unknow_feats is not None else None
    c1 = 0 if unknow_feats is None else unknow_feats.shape[1]
    CI, CF = known_cl.shape[2], 0 if skip_cl is None else skip_cl.shape[2]
    pk = _packed_tf32(mlp, CI, C2, CF, c1, False, 0, "fp")
    return _tf32_launch(mlp, pk, B, n, n, 1, known_cl, m, CI, idx, weight, skip_cl, CF, None, None, None, 0.0, None, False,
                        None, 0, True)


def pointwise_mlp_tf32(x, mlp: PreparedMLP, last_relu=True):
    B, C, n = x.shape
    x_cl = to_cl_f32(x)
    CF = x_cl.shape[2]
    pk = _packed_tf32(mlp, 0, 0, CF, C, False, 0, "pw")
    return _tf32_launch(mlp, pk, B, n, n, 1, None, 0, 0, None, None, x_cl, CF, None, None, None, 0.0, None, False, None, 0,
                        last_relu)


# ----------------------------------------------------------------------------- layout bridge
def to_cl_bf16(features_cf: torch.Tensor, pad_to: Optional[int] = None) -> torch.Tensor:
    """(B,C,N) f32 -> (B,N,Cpad) bf16 channel-last; reuses the `_sad_cl` twin when present."""
    B, C, N = features_cf.shape
    cpad = pad_to or C
    twin = getattr(features_cf, "_sad_cl", None)      # valid while the f32 tensor has not been written in place since
    if twin is not None and tuple(twin.shape) == (B, N, cpad) and getattr(features_cf, "_sad_cl_v", None) == features_cf._version:
        return twin
    x = features_cf.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    if cpad == C:
        out = torch.empty((B, N, C), dtype=torch.bfloat16, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().sad_cf_to_cl_bf16(B, C, N, _ptr(x), _ptr(out), _stream(x)), "cf_to_cl_bf16")
        return out
    out = torch.zeros((B, N, cpad), dtype=torch.bfloat16, device=x.device)
    out[:, :, :C] = to_cl_bf16(x)
    return out


def _attach(cf: Optional[torch.Tensor], cl: Optional[torch.Tensor]):
    if cf is not None and cl is not None:
        cf._sad_cl = cl
        cf._sad_cl_v = cf._version
    return cf


def fused_mlp(mlp: PreparedMLP, layout: Layout, B, N, P, S, feat_cl=None, feat2_cl=None, xyz=None, new_xyz=None,
              idx=None, radius=0.0, radius_t=None, normalize_xyz=False, extra=None, last_relu=True,
              want_cf=True, want_cl=True):
    """Raw launcher of sad_shared_mlp_fwd -> (out_cf (B,Clast,P) f32 | None, out_cl (B,P,Clast) bf16 | None)."""
    imgs, biases, w_ptrs, b_ptrs, c_arr = mlp.packed(layout, S)
    dev = imgs[0].device
    c_last = mlp.c_out[-1]
    out_cf = torch.empty((B, c_last, P), dtype=torch.float32, device=dev) if want_cf else None
    out_cl = torch.empty((B, P, c_last), dtype=torch.bfloat16, device=dev) if want_cl else None
    E = len(layout.extra_cols)
    counter = torch.zeros(1, dtype=torch.int32, device=dev) if DYNAMIC_TILES else None
    opts = _lib.MlpOpts(int(TILES_PER_CTA[0]), int(SUPER_TILES[0]))
    with torch.cuda.device(dev):
        rc = _lib.load().sad_shared_mlp_fwd(
            B, N, P, S, _ptr(feat_cl), layout.c0, _ptr(feat2_cl), layout.c1, _ptr(xyz), _ptr(new_xyz), _ptr(idx),
            float(radius), _ptr(radius_t), int(bool(normalize_xyz)), _ptr(extra), E, len(mlp), w_ptrs, b_ptrs, c_arr,
            int(bool(last_relu)), _ptr(out_cl), _ptr(out_cf), _ptr(counter), ctypes.byref(opts),
            _VP(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "shared_mlp")
    return out_cf, out_cl


# ----------------------------------------------------------------------------- point-wise fast path (csrc/mlp_pw.cu)
FAST_PW = [True]      # False: FP / voting stages through the general kernel (tests compare the two)


def _packed_pw(mlp: PreparedMLP, kind: int):
    key = ("pw", kind)
    if key not in mlp._packed:
        lib = _lib.load()
        dev = mlp.layers[0][0].device
        Ws = [np.ascontiguousarray(W.detach().cpu().numpy(), dtype=np.float32) for (W, _) in mlp.layers]
        bs = [np.ascontiguousarray(b.detach().cpu().numpy(), dtype=np.float32) for (_, b) in mlp.layers]
        c_last = Ws[-1].shape[0]
        img = np.zeros(int(lib.sad_pw_mlp_image_bytes(kind)), dtype=np.uint8)
        W2 = Ws[1] if len(Ws) == 3 else None
        _lib.check(lib.sad_pw_mlp_pack(kind, _VP(Ws[0].ctypes.data), _VP(W2.ctypes.data) if W2 is not None else _VP(0),
                                       _VP(Ws[-1].ctypes.data), int(c_last), _VP(img.ctypes.data)), "pw_mlp_pack")
        bl = np.zeros(384, dtype=np.float32)
        bl[:c_last] = bs[-1]
        mlp._packed[key] = (torch.from_numpy(img).to(dev), torch.from_numpy(bs[0]).to(dev),
                            torch.from_numpy(bs[1]).to(dev) if len(bs) == 3 else None, torch.from_numpy(bl).to(dev))
    return mlp._packed[key]


def _pw_tiles_per_cta():
    return min(2, int(TILES_PER_CTA[0]))      # these stages have 32-64 tiles: keep them spread over the SMs


def fp_interp_mlp_fast(known_cl, skip_cl, idx, weight, mlp: PreparedMLP, B, m, n):
    """FP module in ONE launch: interpolation + concat + MLP -> (out_cf (B,C,n) f32, out_cl (B,n,C) bf16)."""
    img, b1, _, bl = _packed_pw(mlp, 0)
    dev = img.device
    c_last = mlp.c_out[-1]
    out_cf = torch.empty((B, c_last, n), dtype=torch.float32, device=dev)
    out_cl = torch.empty((B, n, c_last), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().sad_pw_mlp_fwd(0, B, n, m, _ptr(skip_cl), _ptr(known_cl), _ptr(idx), _ptr(weight), _ptr(img),
                                        _ptr(b1), _VP(0), _ptr(bl), c_last, _ptr(out_cf), _ptr(out_cl), _VP(0), _VP(0), _VP(0),
                                        _pw_tiles_per_cta(), _VP(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "pw_mlp (FP)")
    return out_cf, out_cl


def vote_mlp_fast(seed_xyz, seed_features, mlp: PreparedMLP):
    """Voting module in ONE launch: 3-layer MLP + (vote = seed + y) -> vote_xyz (B,n,3), vote_features (B,256,n) f32
    carrying its bf16 channel-last twin."""
    B, C, n = seed_features.shape
    img, b1, b2, bl = _packed_pw(mlp, 1)
    dev = img.device
    seed_cl = to_cl_bf16(seed_features)
    sf = seed_features.contiguous()
    vote_xyz = torch.empty((B, n, 3), dtype=torch.float32, device=dev)
    out_cf = torch.empty((B, 256, n), dtype=torch.float32, device=dev)
    out_cl = torch.empty((B, n, 256), dtype=torch.bfloat16, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().sad_pw_mlp_fwd(1, B, n, 0, _ptr(seed_cl), _VP(0), _VP(0), _VP(0), _ptr(img), _ptr(b1), _ptr(b2),
                                        _ptr(bl), 259, _ptr(out_cf), _ptr(out_cl), _ptr(seed_xyz.contiguous()), _ptr(sf),
                                        _ptr(vote_xyz), _pw_tiles_per_cta(),
                                        _VP(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "pw_mlp (voting)")
    return vote_xyz, _attach(out_cf, out_cl)


def vote_fast_ok(seed_features, mlp: PreparedMLP) -> bool:
    B, C, n = seed_features.shape
    return (bool(FAST_PW[0]) and mlp.dtype == "bf16" and C == 256 and n % 128 == 0 and mlp.c_out == [256, 256, 259]
            and seed_features.dtype == torch.float32)


# ----------------------------------------------------------------------------- stage entry points
def sa_group_mlp(xyz, new_xyz, features, idx, radius, mlp: PreparedMLP, use_xyz=True, normalize_xyz=True):
    """Group (relative, optionally radius-normalised xyz ++ features) -> MLP -> max over nsample.
    xyz (B,N,3), new_xyz (B,P,3), features (B,C,N) f32 | None, idx (B,P,S) i32 -> (B,Cout,P) f32."""
    B, P, S = idx.shape
    N = xyz.shape[1]
    c_feat = 0 if features is None else features.shape[1]
    if not mlp.fusable(S) or (c_feat == 0 and not use_xyz):
        _reject("sa_group_mlp", mlp, S)
    if mlp.dtype == "tf32":
        return sa_group_mlp_tf32(xyz, new_xyz, features, idx, radius, mlp, use_xyz, normalize_xyz)
    layout = sa_layout(c_feat, use_xyz or c_feat == 0)
    feat_cl, extra = None, None
    if layout.c0:
        feat_cl = to_cl_bf16(features, pad_to=layout.c0)
    elif c_feat:
        extra = features.contiguous() if c_feat == 1 else features.transpose(1, 2).contiguous()
    radius_t = radius if torch.is_tensor(radius) else None
    inst = _fast_instance(mlp, layout, S, P)
    if inst >= 0:
        out_cf, out_cl = fused_sa_fast(mlp, inst, layout, B, N, P, feat_cl, xyz, new_xyz, idx,
                                       0.0 if radius_t is not None else float(radius), radius_t, normalize_xyz, extra,
                                       want_cf=_WANT_CF[0])
        return _attach(out_cf, out_cl) if out_cf is not None else out_cl
    out_cf, out_cl = fused_mlp(mlp, layout, B, N, P, S, feat_cl=feat_cl, xyz=xyz if layout.xyz_cols else None,
                               new_xyz=new_xyz, idx=idx, radius=0.0 if radius_t is not None else float(radius),
                               radius_t=radius_t, normalize_xyz=normalize_xyz, extra=extra)
    return _attach(out_cf, out_cl)


def fp_interp_mlp(known_feats, unknow_feats, idx, weight, mlp: PreparedMLP):
    """three_interpolate -> concat skip -> MLP.  known_feats (B,C2,m), unknow_feats (B,C1,n) | None,
    idx/weight (B,n,3) -> (B,Cout,n) f32."""
    B, C2, m = known_feats.shape
    n = idx.shape[1]
    c_skip = 0 if unknow_feats is None else unknow_feats.shape[1]
    if not mlp.fusable(1):
        _reject("fp_interp_mlp", mlp, 1)
    if mlp.dtype == "tf32":
        return fp_interp_mlp_tf32(known_feats, unknow_feats, idx, weight, mlp)
    if FAST_PW[0] and C2 == 256 and c_skip == 256 and n % 128 == 0 and mlp.c_out[0] == 256 and len(mlp) == 2 \
            and 8 <= mlp.c_out[1] <= 256:
        out_cf, out_cl = fp_interp_mlp_fast(to_cl_bf16(known_feats), to_cl_bf16(unknow_feats), idx, weight, mlp, B, m, n)
        return _attach(out_cf, out_cl)
    layout = fp_layout(C2, c_skip)
    known_cl = to_cl_bf16(known_feats, pad_to=layout.c0)
    interp_cl = torch.empty((B, n, layout.c0), dtype=torch.bfloat16, device=known_cl.device)
    with torch.cuda.device(known_cl.device):
        _lib.check(_lib.load().sad_three_interpolate_cl_fwd(B, layout.c0, m, n, _ptr(known_cl), _ptr(idx), _ptr(weight),
                                                            _ptr(interp_cl), _stream(known_cl)), "three_interpolate_cl")
    skip_cl = to_cl_bf16(unknow_feats, pad_to=layout.c1) if c_skip else None
    out_cf, out_cl = fused_mlp(mlp, layout, B, n, n, 1, feat_cl=interp_cl, feat2_cl=skip_cl)
    return _attach(out_cf, out_cl)


def pointwise_mlp(x: torch.Tensor, mlp: PreparedMLP, last_relu: bool = True, want_cl: bool = True) -> torch.Tensor:
    """x (B,C,n) -> (B,Cout,n) f32 (voting and other per-point stacks)."""
    B, C, n = x.shape
    if not mlp.fusable(1):
        _reject("pointwise_mlp", mlp, 1)
    if mlp.dtype == "tf32":
        return pointwise_mlp_tf32(x, mlp, last_relu)
    layout = Layout(c0=_round64(C), c0_cols=range(C))
    out_cf, out_cl = fused_mlp(mlp, layout, B, n, n, 1, feat_cl=to_cl_bf16(x, pad_to=layout.c0), last_relu=last_relu,
                               want_cl=want_cl)
    return _attach(out_cf, out_cl)
