"""LCTGenerator forward + backward as one autograd node over the lctgan kernels
(reference models/generator.py:440-632, GRUblockf :31-145, GRUblockt :148-255).

Layout: everything is channels-last [B, T, F, C] and a "row" is one (b, t, f) position with its
C = 64 channels, so LayerNorm / linear layers are plain [M, 64] row operators and the frequency
and time blocks only differ in the stride geometry handed to the GRU and attention kernels.
None of the reference's permute(...).contiguous() copies exist here.

Reference quirks reproduced: the time axis grows 126->129 through the
encoder (k_t = 2, pad 1) and shrinks to 123 through the decoder; skips are cropped from the
low-index corner; the output is ReLU'd, zero-padded back to the input size and only then passed
through the sigmoid, so the last frames are exactly 0.5.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU

SLOPE = 0.2
G = 4          # GRU groups (hard-coded 4 in the reference: generator.py:48, :165)
H = 16         # GRU hidden = input size
HEADS = 4      # attention heads (hard-coded: generator.py:80, :196)
C = 64

BLOCKS = (("GRUf1", True), ("GRUt1", False), ("GRUf2", True))   # (name, bidirectional / frequency block)


def block_param_names(pre: str, bidir: bool) -> List[str]:
    names = []
    for gi in range(1, G + 1):
        sufs = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"]
        if bidir:
            sufs += [s + "_reverse" for s in sufs]
        names += [f"{pre}.gru{gi}.{s}" for s in sufs]
    names += [f"{pre}.attn.in_proj_weight", f"{pre}.attn.in_proj_bias", f"{pre}.attn.out_proj.weight",
              f"{pre}.attn.out_proj.bias", f"{pre}.layernorm1.weight", f"{pre}.layernorm1.bias",
              f"{pre}.layernorm2.weight", f"{pre}.layernorm2.bias", f"{pre}.lin.weight", f"{pre}.lin.bias"]
    return names


def _geom(B: int, T: int, F: int, freq: bool) -> Tuple[int, int, int, int, int, int]:
    """(nseq, L, inner, outer_stride, inner_stride, step_stride) in rows of a [B,T,F,C] buffer."""
    if freq:
        return (B * T, F, 1, F, 0, 1)
    return (B * F, T, F, T * F, 1, F)


def _pack_gru(P: Dict[str, torch.Tensor], pre: str, D: int):
    """Gather the 4*D GRUs' parameters into [GD,48,16] / [GD,48] buffers (gd = group*D + dir)."""
    dev = P[f"{pre}.gru1.weight_ih_l0"].device
    GD = G * D
    wih = torch.empty(GD, 3 * H, H, dtype=torch.float32, device=dev)
    whh = torch.empty(GD, 3 * H, H, dtype=torch.float32, device=dev)
    bih = torch.empty(GD, 3 * H, dtype=torch.float32, device=dev)
    bhh = torch.empty(GD, 3 * H, dtype=torch.float32, device=dev)
    srcs, dsts = [], []
    for g in range(G):
        for d in range(D):
            suf = "_reverse" if d == 1 else ""
            gd = g * D + d
            for name, buf in (("weight_ih_l0", wih), ("weight_hh_l0", whh), ("bias_ih_l0", bih), ("bias_hh_l0", bhh)):
                srcs.append(P[f"{pre}.gru{g + 1}.{name}{suf}"].contiguous())
                dsts.append(buf[gd])
    ops.mt_copy(srcs, dsts)
    return wih, whh, bih, bhh


def _block_fwd(P, pre, x, B, T, F, freq, S):
    """x: [M, 64] rows of a [B,T,F,64] buffer.  Saves what the backward needs into S[pre]."""
    M = x.shape[0]
    D = 2 if freq else 1
    GD = G * D
    geo = _geom(B, T, F, freq)
    dev = x.device
    f32 = dict(dtype=torch.float32, device=dev)
    wih, whh, bih, bhh = _pack_gru(P, pre, D)
    xn, mean1, rstd1 = ops.layernorm_fwd(x, P[f"{pre}.layernorm1.weight"], P[f"{pre}.layernorm1.bias"])
    gi = torch.empty(M, GD, 3 * H, **f32)
    ops.gemm(xn, wih, gi, M, 3 * H, H, lda=C, ldb=H, ldc=GD * 3 * H, bias=bih, nbatch=GD, a_div=D, sA=H,
             sB=3 * H * H, sC=3 * H, sBias=3 * H)
    hs = torch.empty(M, GD, H, **f32)
    # training: every step's gates and starting state are kept, the backward is then a pure chain of gate derivatives
    gsave = torch.empty(M, GD, 4 * H, **f32) if S is not None else None
    hprev = torch.empty(M, GD, H, **f32) if S is not None else None
    ops.call("lct_gru_fwd", gi, whh, bhh, hs, gsave, hprev, geo[0], geo[1], GD, D, geo[2], geo[3], geo[4], geo[5])
    seq = torch.empty(M, C, **f32)
    cat = torch.empty(M, 2 * C, **f32) if freq else None
    ops.call("lct_gru_combine", x, hs, seq, cat, 2 * C, M, G, D)
    sn, mean2, rstd2 = ops.layernorm_fwd(seq, P[f"{pre}.layernorm2.weight"], P[f"{pre}.layernorm2.bias"])
    qkv = torch.empty(M, 3 * C, **f32)
    ops.gemm(sn, P[f"{pre}.attn.in_proj_weight"], qkv, M, 3 * C, C, lda=C, ldb=C, ldc=3 * C,
             bias=P[f"{pre}.attn.in_proj_bias"])
    ao = torch.empty(M, C, **f32)
    lse = torch.empty(M, HEADS, **f32)
    ops.call("lct_attn_fwd", qkv, ao, lse, HEADS, geo[0], geo[1], geo[2], geo[3], geo[4], geo[5])
    mix = torch.empty(M, C, **f32)
    out = torch.empty(M, C, **f32)
    if freq:
        # attention out-projection lands in the right half of the concat buffer: cat = [gru | attn]
        ops.gemm(ao, P[f"{pre}.attn.out_proj.weight"], cat, M, C, C, lda=C, ldb=C, ldc=2 * C,
                 bias=P[f"{pre}.attn.out_proj.bias"], c_off=C)
        ops.gemm(cat, P[f"{pre}.lin.weight"], mix, M, C, 2 * C, lda=2 * C, ldb=2 * C, ldc=C,
                 bias=P[f"{pre}.lin.bias"], act=ACT_LRELU, slope=SLOPE, res=seq, ldr=C, out2=out, ldo=C)
        lin_in = cat
    else:
        a2 = torch.empty(M, C, **f32)
        ops.gemm(ao, P[f"{pre}.attn.out_proj.weight"], a2, M, C, C, lda=C, ldb=C, ldc=C,
                 bias=P[f"{pre}.attn.out_proj.bias"])
        ops.gemm(a2, P[f"{pre}.lin.weight"], mix, M, C, C, lda=C, ldb=C, ldc=C, bias=P[f"{pre}.lin.bias"],
                 act=ACT_LRELU, slope=SLOPE, res=seq, ldr=C, out2=out, ldo=C)
        lin_in = a2
    if S is not None:
        S[pre] = dict(x=x, xn=xn, mean1=mean1, rstd1=rstd1, seq=seq, sn=sn, mean2=mean2, rstd2=rstd2,
                      qkv=qkv, ao=ao, lse=lse, mix=mix, lin_in=lin_in, wih=wih, whh=whh, gsave=gsave, hprev=hprev, geo=geo,
                      D=D)
    return out


def _split_k(M: int) -> int:
    # the weight-gradient GEMMs have 1..3 output tiles: ~133 rows per CTA puts 256+ CTAs on the 148 SMs
    return max(1, min(256, (M + 127) // 128))


class _Side:
    """Runs parameter-gradient kernels of the generator backward on a forked stream: they are off the critical
    data-gradient chain and the chain's kernels are too small to fill the GPU on their own.  `fork()` orders the
    side stream after everything enqueued so far; tensors handed to it are kept alive until `join()`."""

    def __init__(self, device):
        from . import config
        self.cur = torch.cuda.current_stream(device)
        self.side = config.generator_side_stream(device) if config.concurrent_discriminators else None
        self.keep = []

    def fork(self, *tensors):
        self.keep.extend(tensors)
        if self.side is None:
            return torch.cuda.stream(self.cur)
        self.side.wait_stream(self.cur)
        return torch.cuda.stream(self.side)

    def join(self):
        if self.side is not None:
            self.cur.wait_stream(self.side)
        self.keep.clear()


def _block_bwd(P, pre, dout, B, T, F, freq, S, GR, side=None, z=None):
    """dout: [M,64] gradient of the block output.  Fills GR[name] for the block's parameters and
    returns the gradient of the block input."""
    own = side is None
    if own:
        side = _Side(dout.device)
    s = S[pre]
    M = dout.shape[0]
    D = s["D"]
    GD = G * D
    geo = s["geo"]
    dev = dout.device
    f32 = dict(dtype=torch.float32, device=dev)
    ks = _split_k(M)
    if z is None:
        z = lambda *shape: ops.zeros(shape, dev)

    # out = seq + lrelu(lin(lin_in))
    dmix = ops.act_bwd(s["mix"], dout, ACT_LRELU, SLOPE)
    kin = 2 * C if freq else C
    with side.fork(dmix):          # parameter gradients of `lin`
        dlin_w = z(C, kin)
        ops.gemm(dmix, s["lin_in"], dlin_w, C, kin, M, lda=C, ldb=kin, ldc=kin, ta=True, tb=True, ksplit=ks)
        dlin_b = z(C)
        ops.colsum(dmix, dlin_b, M, C, C)
    dlin_in = torch.empty(M, kin, **f32)
    ops.gemm(dmix, P[f"{pre}.lin.weight"], dlin_in, M, kin, C, lda=C, ldb=kin, ldc=kin, tb=True)
    GR[f"{pre}.lin.weight"], GR[f"{pre}.lin.bias"] = dlin_w, dlin_b

    # attention out-projection: its output is lin_in (t block) or the right half of cat (f block)
    a_off = C if freq else 0
    with side.fork(dlin_in):
        dwo = z(C, C)
        ops.gemm(dlin_in, s["ao"], dwo, C, C, M, lda=kin, ldb=C, ldc=C, ta=True, tb=True, ksplit=ks, a_off=a_off)
        dbo = z(C)
        ops.colsum(dlin_in.view(-1)[a_off:], dbo, M, C, kin)
    dao = torch.empty(M, C, **f32)
    ops.gemm(dlin_in, P[f"{pre}.attn.out_proj.weight"], dao, M, C, C, lda=kin, ldb=C, ldc=C, tb=True, a_off=a_off)
    GR[f"{pre}.attn.out_proj.weight"], GR[f"{pre}.attn.out_proj.bias"] = dwo, dbo

    dqkv = torch.empty(M, 3 * C, **f32)
    ops.call("lct_attn_bwd", s["qkv"], s["ao"], s["lse"], dao, dqkv, HEADS, geo[0], geo[1], geo[2], geo[3], geo[4],
             geo[5])
    with side.fork(dqkv):
        dwi = z(3 * C, C)
        ops.gemm(dqkv, s["sn"], dwi, 3 * C, C, M, lda=3 * C, ldb=C, ldc=C, ta=True, tb=True, ksplit=ks)
        dbi = z(3 * C)
        ops.colsum(dqkv, dbi, M, 3 * C, 3 * C)
    dsn = torch.empty(M, C, **f32)
    ops.gemm(dqkv, P[f"{pre}.attn.in_proj_weight"], dsn, M, C, 3 * C, lda=3 * C, ldb=C, ldc=C, tb=True)
    GR[f"{pre}.attn.in_proj_weight"], GR[f"{pre}.attn.in_proj_bias"] = dwi, dbi

    # seq = x + gru: dseq = dout (residual) + LN2 backward
    dg2, db2 = z(C), z(C)
    dseq = ops.layernorm_bwd(dsn, s["seq"], P[f"{pre}.layernorm2.weight"], s["mean2"], s["rstd2"], dg2, db2, dres=dout)
    GR[f"{pre}.layernorm2.weight"], GR[f"{pre}.layernorm2.bias"] = dg2, db2

    # gradient reaching the GRU output: through seq, and (f block) through the left half of cat
    if freq:
        dgru = torch.empty(M, C, **f32)
        # dgru = dseq + dlin_in[:, :64]  (strided read of the concat gradient)
        ops.gemm_free_add(dseq, dlin_in, dgru, M, C, kin)
    else:
        dgru = dseq
    dgi = torch.empty(M, GD, 3 * H, **f32)
    dgh = torch.empty(M, GD, 3 * H, **f32)
    ops.call("lct_gru_bwd", s["gsave"], s["hprev"], s["whh"], dgru, C, dgi, dgh, geo[0], geo[1], GD, D, geo[2], geo[3],
             geo[4], geo[5])
    # dW_ih[gd] = dgi[:, gd, :]^T @ xn[:, g*16:(g+1)*16];  dW_hh[gd] = dgh[:, gd, :]^T @ hprev[:, gd, :];  biases = column sums
    with side.fork(dgi, dgh):
        ksg = max(1, 2 * ks // GD)
        dwih, dwhh, dbih, dbhh = z(GD, 3 * H, H), z(GD, 3 * H, H), z(GD, 3 * H), z(GD, 3 * H)
        ops.gemm(dgi, s["xn"], dwih, 3 * H, H, M, lda=GD * 3 * H, ldb=C, ldc=H, ta=True, tb=True, ksplit=ksg,
                 nbatch=GD, a_div=1, b_div=D, sA=3 * H, sB=H, sC=3 * H * H)
        ops.gemm(dgh, s["hprev"], dwhh, 3 * H, H, M, lda=GD * 3 * H, ldb=GD * H, ldc=H, ta=True, tb=True, ksplit=ksg,
                 nbatch=GD, a_div=1, b_div=1, sA=3 * H, sB=H, sC=3 * H * H)
        ops.colsum(dgi, dbih, M, GD * 3 * H, GD * 3 * H)
        ops.colsum(dgh, dbhh, M, GD * 3 * H, GD * 3 * H)
    # dxn[:, g] = dgi[:, g, (d,48)] @ [W_ih(g,0); W_ih(g,1)]
    dxn = torch.empty(M, C, **f32)
    ops.gemm(dgi, s["wih"], dxn, M, H, D * 3 * H, lda=GD * 3 * H, ldb=H, ldc=C, tb=True, nbatch=G, sA=D * 3 * H,
             sB=D * 3 * H * H, sC=H)
    for g in range(G):
        for d in range(D):
            suf = "_reverse" if d == 1 else ""
            gd = g * D + d
            GR[f"{pre}.gru{g + 1}.weight_ih_l0{suf}"] = dwih[gd]
            GR[f"{pre}.gru{g + 1}.weight_hh_l0{suf}"] = dwhh[gd]
            GR[f"{pre}.gru{g + 1}.bias_ih_l0{suf}"] = dbih[gd]
            GR[f"{pre}.gru{g + 1}.bias_hh_l0{suf}"] = dbhh[gd]
    dg1, db1 = z(C), z(C)
    dx = ops.layernorm_bwd(dxn, s["x"], P[f"{pre}.layernorm1.weight"], s["mean1"], s["rstd1"], dg1, db1, dres=dseq)
    GR[f"{pre}.layernorm1.weight"], GR[f"{pre}.layernorm1.bias"] = dg1, db1
    if own:
        side.join()
    return dx


def conv_out_tf(T: int, F: int) -> Tuple[int, int]:
    return T + 1, (F - 1) // 2 + 1


def deconv_out_tf(T: int, F: int) -> Tuple[int, int]:
    return T - 1, 2 * F


def generator_forward(P: Dict[str, torch.Tensor], mag: torch.Tensor, use_sigmoid: bool, S):
    """mag: [B, T, F] (physical spectrogram layout).  Returns mask [B, T, F]."""
    B, Tm, Fm = mag.shape
    mag4 = mag.view(B, Tm, Fm, 1)
    chans = (P["conv1.weight"].shape[0], P["conv2.weight"].shape[0], P["conv3.weight"].shape[0])
    x = mag4
    enc = []
    T, F = Tm, Fm
    for i, co in enumerate(chans, 1):
        T, F = conv_out_tf(T, F)
        x = ops.gconv(x, P[f"conv{i}.weight"], P[f"conv{i}.bias"], (T, F), co, transposed=False, act=ACT_LRELU,
                      slope=SLOPE)
        enc.append(x)
    T3, F3 = T, F
    M = B * T3 * F3
    x3 = enc[-1]
    h, mean0, rstd0 = ops.layernorm_fwd(x3.view(M, C), P["layernorm.weight"], P["layernorm.bias"])
    for name, freq in BLOCKS:
        h = _block_fwd(P, name, h, B, T3, F3, freq, S)
    h = h.view(B, T3, F3, C)
    dec_in, dec_out = [], []
    for i, act in ((2, ACT_LRELU), (3, ACT_LRELU), (4, ACT_RELU)):
        w = P[f"deconv{i}.weight"]
        d_in = ops.skip_add_fwd(h, mag, P[f"skip{i}.weight"].reshape(-1), P[f"skip{i}.bias"])
        To, Fo = deconv_out_tf(d_in.shape[1], d_in.shape[2])
        y = ops.gconv(d_in, w, P[f"deconv{i}.bias"], (To, Fo), w.shape[1], transposed=True, act=act, slope=SLOPE)
        dec_in.append(d_in)
        dec_out.append(y)
        h = y
    y4 = dec_out[-1]
    mask = ops.final_mask_fwd(y4, Tm, Fm, use_sigmoid)
    if S is not None:
        S["top"] = dict(enc=enc, mean0=mean0, rstd0=rstd0, dec_in=dec_in, dec_out=dec_out, T3=T3, F3=F3)
    return mask


def generator_backward(P, mag, mask, gmask, use_sigmoid, S, need_mag_grad=False, arena_out=None):
    if need_mag_grad:
        raise RuntimeError("LCTGenerator: gradient w.r.t. the input magnitude is not implemented "
                           "(the training step never needs it: the noisy spectrum is data)")
    top = S["top"]
    B, Tm, Fm = mag.shape
    GR: Dict[str, torch.Tensor] = {}
    dev = mag.device
    f32 = dict(dtype=torch.float32, device=dev)
    # every parameter-gradient accumulator of this backward is carved from ONE buffer cleared by one memset (round 1:
    # ~90 separate torch.zeros fill kernels); the buffer is also what the fused gradient clip reduces over
    z = ops.Arena(sum((p.numel() + 3) // 4 * 4 for p in P.values()) + 1024, dev)
    if arena_out is not None:
        arena_out.append(z.buf)
    enc, dec_in, dec_out = top["enc"], top["dec_in"], top["dec_out"]
    T3, F3 = top["T3"], top["F3"]
    M = B * T3 * F3
    side = _Side(dev)

    # ---- decoder
    dpre = ops.final_mask_bwd(dec_out[2], mask, gmask.contiguous(), use_sigmoid, act=ACT_RELU)
    prev_outs = [None, dec_out[0], dec_out[1]]      # what feeds skip-add i (None: the bottleneck output)
    dh3 = None
    for j, i in ((2, 4), (1, 3), (0, 2)):
        w = P[f"deconv{i}.weight"]
        d_in = dec_in[j]
        Bq, Ti, Fi, Ci = d_in.shape
        Co = w.shape[1]
        with side.fork(dpre):
            GR[f"deconv{i}.weight"] = ops.gconv_wgrad(d_in, dpre, w.shape, out=z(*w.shape))
            db = z(Co)
            ops.colsum(dpre, db, dpre.numel() // Co, Co, Co)
            GR[f"deconv{i}.bias"] = db
        g_in = ops.gconv(dpre, w, None, (Ti, Fi), Ci, transposed=False)
        dw, dbs = z(Ci), z(Ci)
        src = prev_outs[j]
        h_shape = (B, T3, F3, C) if src is None else tuple(src.shape)
        dh = ops.skip_add_bwd(g_in, mag, h_shape, dw, dbs, want_dh=True)
        GR[f"skip{i}.weight"] = dw.view(P[f"skip{i}.weight"].shape)
        GR[f"skip{i}.bias"] = dbs
        if src is None:
            dh3 = dh
        else:
            dpre = ops.act_bwd(src, dh, ACT_LRELU, SLOPE)

    # ---- bottleneck
    dh = dh3.view(M, C)
    for name, freq in reversed(BLOCKS):
        dh = _block_bwd(P, name, dh, B, T3, F3, freq, S, GR, side, z)
    dg0, db0 = z(C), z(C)
    dx3 = ops.layernorm_bwd(dh, enc[2].view(M, C), P["layernorm.weight"], top["mean0"], top["rstd0"], dg0, db0)
    GR["layernorm.weight"], GR["layernorm.bias"] = dg0, db0

    # ---- encoder
    dpre = ops.act_bwd(enc[2], dx3.view(enc[2].shape), ACT_LRELU, SLOPE)
    mag4 = mag.view(B, Tm, Fm, 1)
    inputs = [mag4, enc[0], enc[1]]
    for i in (3, 2, 1):
        w = P[f"conv{i}.weight"]
        xin = inputs[i - 1]
        Co = w.shape[0]
        with side.fork(dpre):
            GR[f"conv{i}.weight"] = ops.gconv_wgrad(dpre, xin, w.shape, out=z(*w.shape))
            db = z(Co)
            ops.colsum(dpre, db, dpre.numel() // Co, Co, Co)
            GR[f"conv{i}.bias"] = db
        if i > 1:
            dpre = ops.gconv(dpre, w, None, (xin.shape[1], xin.shape[2]), xin.shape[3], transposed=True, gmul=xin,
                             gact=ACT_LRELU, gslope=SLOPE)
    side.join()
    return GR


def _on_generator_stream(device, fn):
    """Run fn on the high-priority generator stream, forked from and joined to the caller's stream (the generator is a
    serial chain of small kernels: it must not queue behind the wide weight-gradient grids of the discriminators)."""
    from . import config
    if not config.concurrent_discriminators:
        return fn()
    cur = torch.cuda.current_stream(device)
    hp = config.generator_stream(device)
    hp.wait_stream(cur)
    with torch.cuda.stream(hp):
        out = fn()
    cur.wait_stream(hp)
    return out


class GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mag_phys, use_sigmoid, names, owner, keep_state, *params):
        """owner: the LCTGenerator module (receives ``_grad_arena``, the flat buffer holding the gradients of the last
        backward); keep_state: torch.is_grad_enabled() at the call site - ``ctx.needs_input_grad`` alone reflects
        ``requires_grad`` even under no_grad (infer.py, the D step's enhancer pass), where keeping every activation
        until forward returns would cost the training-size memory peak for nothing."""
        if not mag_phys.is_cuda:
            raise RuntimeError("LCTGenerator (lctgan) is CUDA only (sm_100a); there is no CPU fallback")
        P = dict(zip(names, params))
        need = bool(keep_state) and any(ctx.needs_input_grad[5:])
        S = {} if need else None
        mask = _on_generator_stream(mag_phys.device, lambda: generator_forward(P, mag_phys.contiguous(), use_sigmoid, S))
        ctx.S = S
        ctx.names = names
        ctx.owner = owner
        ctx.use_sigmoid = use_sigmoid
        ctx.save_for_backward(mag_phys, mask, *params)
        return mask

    @staticmethod
    def backward(ctx, gmask):
        if ctx.S is None:
            raise RuntimeError("LCTGenerator backward: the saved activations are gone - either the forward ran under "
                               "no_grad, or this is a second backward through the same graph (retain_graph=True is not "
                               "supported: the activations are released after the first backward)")
        mag_phys, mask, *params = ctx.saved_tensors
        P = dict(zip(ctx.names, params))
        arena = []
        if mag_phys.is_cuda:
            # the G step's dead discriminator parameter gradients (lctgan.config.defer) start here: the chain of small
            # kernels below leaves most of the GPU idle
            from . import config
            config.launch_deferred_param_grads()
        GR = _on_generator_stream(mag_phys.device, lambda: generator_backward(
            P, mag_phys.contiguous(), mask, gmask, ctx.use_sigmoid, ctx.S, ctx.needs_input_grad[0], arena))
        ctx.S = None
        if ctx.owner is not None and arena:
            ctx.owner._grad_arena = arena[0]
        grads = [GR.get(n) if ctx.needs_input_grad[5 + i] else None for i, n in enumerate(ctx.names)]
        return (None, None, None, None, None, *grads)
