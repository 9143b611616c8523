"""TEST INFRASTRUCTURE: a torch-CPU interpreter of d3fk op lists (fp32 only).

It executes the exact `d3fk_op` records the product hands to libd3fk, reading and writing the raw host
pointers stored in them, with plain torch ops as the semantics of each op (include/d3fk.h).  It lets
the `-m "not gpu"` suite validate the host logic (plan topology, buffer wiring, backward ordering,
gradient accumulation) against the oracle without a GPU.  It is never imported by the product."""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from denoising_diffusion_deep_fake_b200 import _lib

_CT = {torch.float32: C.c_float, torch.float64: C.c_double, torch.uint8: C.c_uint8, torch.int64: C.c_int64,
       torch.int32: C.c_int32}


def view(ptr, shape, dtype=torch.float32):
    n = 1
    for s in shape:
        n *= int(s)
    if n == 0:
        return torch.empty(shape, dtype=dtype)
    buf = (_CT[dtype] * n).from_address(ptr)
    return torch.frombuffer(buf, dtype=dtype).view(*shape)


def nhwc(ptr, B, H, W, ld, C_):
    return view(ptr, (B, H, W, ld))[..., :C_]


def _gather_input(p):
    Hs, Ws = (p.Hi >> p.up0), (p.Wi >> p.up0)
    a = nhwc(p.src0, p.B, Hs, Ws, p.ld0, p.c0).permute(0, 3, 1, 2)
    if p.up0:
        a = F.interpolate(a, scale_factor=2, mode="nearest")
    if p.c1:
        b = nhwc(p.src1, p.B, p.Hi, p.Wi, p.ld1, p.c1).permute(0, 3, 1, 2)
        a = torch.cat([a, b], dim=1)
    return a.contiguous()


def run_conv(p):
    assert p.dtype == _lib.F32
    ctot = p.c0 + p.c1
    if p.mode == 0:
        a = _gather_input(p)
        w = view(p.w, (p.Cout, p.kh, p.kw, ctot)).permute(0, 3, 1, 2).contiguous()
        v = F.conv2d(a, w, stride=p.stride, padding=p.pad)
    else:
        assert p.c1 == 0 and p.up0 == 0
        dy = nhwc(p.src0, p.B, p.Hi, p.Wi, p.ld0, p.c0).permute(0, 3, 1, 2).contiguous()
        w = view(p.w, (p.Cout, p.kh, p.kw, p.c0)).permute(3, 0, 1, 2).contiguous()   # [in=c0][out][kh][kw]
        oph = p.Ho - ((p.Hi - 1) * p.stride - 2 * p.pad + p.kh)
        opw = p.Wo - ((p.Wi - 1) * p.stride - 2 * p.pad + p.kw)
        v = F.conv_transpose2d(dy, w, stride=p.stride, padding=p.pad, output_padding=(oph, opw))
    assert v.shape[2] == p.Ho and v.shape[3] == p.Wo, (v.shape, p.Ho, p.Wo)
    if p.scale:
        v = v * view(p.scale, (p.Cout,)).view(1, -1, 1, 1) + view(p.shift, (p.Cout,)).view(1, -1, 1, 1)
    elif p.shift:
        v = v + view(p.shift, (p.Cout,)).view(1, -1, 1, 1)
    if p.res:
        v = v + nhwc(p.res, p.B, p.Ho, p.Wo, p.ldr, p.Cout).permute(0, 3, 1, 2)
    if p.relu:
        v = v.clamp_min(0)
    if p.stats and p.bw_x:
        # fused BN-backward reduction: sums of g' and g' * xhat (see d3fk_conv_params.bw_*)
        st = view(p.stats, (2, p.Cout), torch.float64)
        x = nhwc(p.bw_x, p.B, p.Ho, p.Wo, p.bw_ldx, p.Cout).permute(0, 3, 1, 2)
        g = v
        if p.bw_relu:
            act = nhwc(p.bw_act, p.B, p.Ho, p.Wo, p.bw_ldact, p.Cout).permute(0, 3, 1, 2)
            g = torch.where(act > 0, v, torch.zeros_like(v))
        xh = (x - view(p.bw_mean, (p.Cout,)).view(1, -1, 1, 1)) * view(p.bw_invstd, (p.Cout,)).view(1, -1, 1, 1)
        st[0] += g.double().sum(dim=(0, 2, 3))
        st[1] += (g.double() * xh.double()).sum(dim=(0, 2, 3))
    elif p.stats:
        st = view(p.stats, (2, p.Cout), torch.float64)
        st[0] += v.double().sum(dim=(0, 2, 3))
        st[1] += (v.double() ** 2).sum(dim=(0, 2, 3))
    if p.out_nchw:
        view(p.out_nchw, (p.B, p.Cout, p.Ho, p.Wo)).copy_(v)
    else:
        nhwc(p.out, p.B, p.Ho, p.Wo, p.ldo, p.Cout).copy_(v.permute(0, 2, 3, 1))


def run_wgrad(p):
    a = _gather_input(p)
    dy = nhwc(p.dy, p.B, p.Ho, p.Wo, p.ldy, p.Cout).permute(0, 3, 1, 2).contiguous()
    ctot = p.c0 + p.c1
    g = torch.nn.grad.conv2d_weight(a, (p.Cout, ctot, p.kh, p.kw), dy, stride=p.stride, padding=p.pad)
    dw = view(p.dw, (p.cout_real, p.cin_real, p.kh, p.kw))
    dw += g[:p.cout_real, :p.cin_real]


def run_pack(p):
    w = view(p.w, (p.Cout, p.Cin, p.kh, p.kw))
    if p.w_fwd:
        out = view(p.w_fwd, (p.Cout, p.kh, p.kw, p.cin_pad))
        out.zero_()
        out[..., :p.Cin] = w.permute(0, 2, 3, 1)
    if p.w_dgrad:
        out = view(p.w_dgrad, (p.Cin, p.kh, p.kw, p.cout_pad))
        out.zero_()
        out[..., :p.Cout] = w.permute(1, 2, 3, 0)


def run_layout(p):
    src = view(p.src, (p.B, p.C, p.H, p.W))
    dst = view(p.dst, (p.B, p.H, p.W, p.cpad))
    dst.zero_()
    dst[..., :p.C] = src.permute(0, 2, 3, 1)
    if p.chansum:
        view(p.chansum, (p.C,)).add_(dst[..., :p.C].reshape(-1, p.C).sum(0))


def run_bn_finalize(p):
    st = view(p.stats, (2, p.C), torch.float64)
    n = float(p.count)
    mean = st[0] / n
    var = (st[1] / n - mean * mean).clamp_min(0)
    invstd = 1.0 / torch.sqrt(var + p.eps)
    g, b = view(p.gamma, (p.C,)).double(), view(p.beta, (p.C,)).double()
    view(p.scale, (p.C,)).copy_((g * invstd).float())
    view(p.shift, (p.C,)).copy_((b - mean * g * invstd).float())
    view(p.mean, (p.C,)).copy_(mean.float())
    view(p.invstd, (p.C,)).copy_(invstd.float())
    if p.running_mean:
        rm, rv = view(p.running_mean, (p.C,)), view(p.running_var, (p.C,))
        unb = var * n / (n - 1) if n > 1 else var
        rm.copy_(((1 - p.momentum) * rm.double() + p.momentum * mean).float())
        rv.copy_(((1 - p.momentum) * rv.double() + p.momentum * unb).float())
    if p.num_batches_tracked:
        view(p.num_batches_tracked, (1,), torch.int64).add_(1)


def run_bn_fold(p):
    rv, rm = view(p.running_var, (p.C,)), view(p.running_mean, (p.C,))
    sc = view(p.gamma, (p.C,)) / torch.sqrt(rv + p.eps)
    view(p.scale, (p.C,)).copy_(sc)
    view(p.shift, (p.C,)).copy_(view(p.beta, (p.C,)) - rm * sc)


def _rows(ptr, count, ld, C_):
    return view(ptr, (count, ld))[:, :C_]


def run_bn_apply(p):
    x = _rows(p.x, p.count, p.ldx, p.C)
    if p.stats:                     # fused finalize (train forward)
        run_bn_finalize(p)
    v = x * view(p.scale, (p.C,)) + view(p.shift, (p.C,))
    if p.res:
        v = v + _rows(p.res, p.count, p.ldr, p.C)
    if p.relu:
        v = v.clamp_min(0)
    _rows(p.y, p.count, p.ldy, p.C).copy_(v)


def _masked_dy(p):
    dy = _rows(p.dy, p.count, p.lddy, p.C)
    if p.relu:
        dy = dy * (_rows(p.act, p.count, p.ldact, p.C) > 0)
    return dy


def run_bn_bwd_reduce(p):
    dy = _masked_dy(p).double()
    xh = (_rows(p.x, p.count, p.ldx, p.C) - view(p.mean, (p.C,))) * view(p.invstd, (p.C,))
    bs = view(p.bstats, (2, p.C), torch.float64)
    bs[0] += dy.sum(0)
    bs[1] += (dy * xh.double()).sum(0)


def run_bn_bwd_finalize(p):
    bs = view(p.bstats, (2, p.C), torch.float64)
    n = float(p.count)
    if p.dbeta:
        view(p.dbeta, (p.C,)).copy_(bs[0].float())
    if p.dgamma:
        view(p.dgamma, (p.C,)).copy_(bs[1].float())
    coef = view(p.coef, (3, p.C))
    coef[0] = view(p.gamma, (p.C,)) * view(p.invstd, (p.C,))
    coef[1] = (bs[0] / n).float()
    coef[2] = (bs[1] / n).float()


def run_bn_bwd_apply(p):
    run_bn_bwd_finalize(p)          # the kernel derives its coefficients and writes dgamma/dbeta itself
    dy = _masked_dy(p)
    xh = (_rows(p.x, p.count, p.ldx, p.C) - view(p.mean, (p.C,))) * view(p.invstd, (p.C,))
    coef = view(p.coef, (3, p.C))
    dx = coef[0] * (dy - coef[1] - xh * coef[2])
    if p.dres:
        _rows(p.dres, p.count, p.lddres, p.C).copy_(dy)
    _rows(p.dx, p.count, p.lddx, p.C).copy_(dx)


def run_maxpool_fwd(p):
    x = nhwc(p.x, p.B, p.H, p.W, p.ldx, p.C)
    Ho, Wo = p.H // 2, p.W // 2
    xp = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1), value=float("-inf"))
    best = torch.full((p.B, p.C, Ho, Wo), float("-inf"))
    idx = torch.zeros((p.B, p.C, Ho, Wo), dtype=torch.uint8)
    first = torch.ones((p.B, p.C, Ho, Wo), dtype=torch.bool)
    for kh in range(3):
        for kw in range(3):
            v = xp[:, :, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2]
            valid = torch.isfinite(v) | (v == float("inf"))
            take = valid & (first | (v > best))
            best = torch.where(take, v, best)
            idx = torch.where(take, torch.tensor(kh * 3 + kw, dtype=torch.uint8), idx)
            first = first & ~take
    nhwc(p.y, p.B, Ho, Wo, p.ldy, p.C).copy_(best.permute(0, 2, 3, 1))
    if p.idx:
        view(p.idx, (p.B, Ho, Wo, p.C), torch.uint8).copy_(idx.permute(0, 2, 3, 1))


def run_maxpool_bwd(p):
    Ho, Wo = p.H // 2, p.W // 2
    dy = nhwc(p.dy, p.B, Ho, Wo, p.lddy, p.C).permute(0, 3, 1, 2)
    idx = view(p.idx, (p.B, Ho, Wo, p.C), torch.uint8).permute(0, 3, 1, 2)
    g = torch.zeros((p.B, p.C, p.H + 2, p.W + 2))
    for kh in range(3):
        for kw in range(3):
            g[:, :, kh:kh + 2 * Ho:2, kw:kw + 2 * Wo:2] += dy * (idx == kh * 3 + kw)
    g = g[:, :, 1:-1, 1:-1].permute(0, 2, 3, 1)
    dx = nhwc(p.dx, p.B, p.H, p.W, p.lddx, p.C)
    if p.accumulate:
        dx += g
    else:
        dx.copy_(g)


def run_sumpool2(p):
    dy = nhwc(p.dy, p.B, 2 * p.H, 2 * p.W, p.lddy, p.C)
    g = dy[:, 0::2, 0::2] + dy[:, 0::2, 1::2] + dy[:, 1::2, 0::2] + dy[:, 1::2, 1::2]
    dx = nhwc(p.dx, p.B, p.H, p.W, p.lddx, p.C)
    if p.accumulate:
        dx += g
    else:
        dx.copy_(g)


def run_chansum(p):
    x = _rows(p.x, p.count, p.ld, p.C)
    view(p.out, (p.C,)).add_(x.sum(0))


def run_pack_all(p):
    n = (p.n >> 1) & 0xFFFF
    table = (_lib.PackParams * n).from_address(p.p0)
    for i in range(n):
        run_pack(table[i])


def run_wgrad_group(p):
    for i in range(p.count):
        one = _lib.WgradParams.from_buffer_copy(bytes(p.base))
        one.src0, one.dy, one.dw = p.src0[i], p.dy[i], p.dw[i]
        run_wgrad(one)


def run_bn_bwd(p):
    run_bn_bwd_reduce(p)
    run_bn_bwd_apply(p)


def run_upcat(p):
    a = nhwc(p.src0, p.B, p.H // 2, p.W // 2, p.ld0, p.c0)
    out = nhwc(p.out, p.B, p.H, p.W, p.ldo, p.c0 + p.c1)
    out[..., :p.c0] = a.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    if p.c1:
        out[..., p.c0:] = nhwc(p.src1, p.B, p.H, p.W, p.ld1, p.c1)


def run_conv_bn(p):
    run_conv(p.conv)
    run_bn_apply(p.bn)


def run_memset(p):
    view(p.p0, (p.n,), torch.uint8).zero_()


_DISPATCH = {
    _lib.OP_CONV: run_conv, _lib.OP_WGRAD: run_wgrad, _lib.OP_PACK: run_pack, _lib.OP_NCHW2NHWC: run_layout,
    _lib.OP_BN_FINALIZE: run_bn_finalize, _lib.OP_BN_APPLY: run_bn_apply, _lib.OP_BN_FOLD: run_bn_fold,
    _lib.OP_BN_BWD_REDUCE: run_bn_bwd_reduce, _lib.OP_BN_BWD_FINALIZE: run_bn_bwd_finalize,
    _lib.OP_BN_BWD_APPLY: run_bn_bwd_apply, _lib.OP_MAXPOOL_FWD: run_maxpool_fwd, _lib.OP_MAXPOOL_BWD: run_maxpool_bwd,
    _lib.OP_SUMPOOL2: run_sumpool2, _lib.OP_CHANSUM: run_chansum, _lib.OP_MEMSET: run_memset,
    _lib.OP_PACK_ALL: run_pack_all, _lib.OP_CONV_BN: run_conv_bn, _lib.OP_UPCAT: run_upcat, _lib.OP_BN_BWD: run_bn_bwd,
    _lib.OP_WGRAD_GROUP: run_wgrad_group,
}


def run_ops(oplist):
    for op in oplist:
        if op.kind == _lib.OP_JOIN:          # stream bookkeeping only: the interpreter runs every lane in list order
            continue
        _DISPATCH[op.kind](_lib.op_params(op))
