"""CPU suite: the host logic of the hot path.  The exact op lists the product hands to libd3fk are executed by
the torch-CPU interpreter (tests/op_interpreter.py) and compared with the oracle: forward (train / eval),
BN running statistics, every parameter gradient, plan bookkeeping."""
import copy

import pytest
import torch

import oracle
import denoising_diffusion_deep_fake_b200 as d3
from denoising_diffusion_deep_fake_b200 import _lib
from denoising_diffusion_deep_fake_b200.plan import UnetPlan, all_convs, backward_param_order, unet_layers
import op_interpreter as I


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def make_pair(seed=0):
    torch.manual_seed(seed)
    ref = oracle.Unet()
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if p.dim() == 1 and "segmentation_head" not in n:
                p.uniform_(0.5, 1.5) if n.endswith("weight") else p.uniform_(-0.3, 0.3)
    m = d3.Unet(precision="fp32")
    m.load_state_dict(ref.state_dict())
    return ref, m


def cpu_plan(m, B, H, W, training):
    m._ensure_grad_arena(torch.device("cpu"))
    return UnetPlan(dict(m.named_parameters()), dict(m.named_buffers()), B, H, W, _lib.F32, "cpu", training,
                    grad_arena=m._grad_arena, grad_offsets=m._grad_offsets)


def run_forward(plan, x):
    y = torch.empty_like(x)
    I.run_ops(plan.pack_ops)
    _lib.op_params(plan.fwd_ops.array[plan.in_op_index]).src = x.data_ptr()
    _lib.op_params(plan.fwd_ops.array[plan.out_op_index]).out_nchw = y.data_ptr()
    I.run_ops(plan.fwd_ops)
    return y


def test_topology_tables():
    convs = all_convs()
    assert len(convs) == 47 and sum(1 for c in convs if c.bn) == 46
    ref, m = make_pair()
    names = backward_param_order()
    assert sorted(names) == sorted(n for n, _ in m.named_parameters())
    assert names[0] == "segmentation_head.0.weight" and names[-1] == "encoder.bn1.bias"
    buckets = m.grad_buckets()
    assert buckets[0][0] == 0 and buckets[-1][1] == m._grad_numel
    assert all(a[1] == b[0] for a, b in zip(buckets, buckets[1:])) and len(buckets) == 5
    sizes = [e - s for s, e in buckets]
    assert sizes[1] >= 13_114_368 and sizes[2] >= 6_822_400      # layer4 / layer3 parameters
    stem, stages, dec, head = unet_layers()
    assert [len(s) for s in stages] == [3, 4, 6, 3] and dec[0]["conv1"].cin == 768 and head.cout == 3


@pytest.mark.parametrize("B,H,W", [(2, 64, 64), (3, 32, 96)])
def test_train_plan_matches_oracle(B, H, W):
    ref, m = make_pair()
    x = torch.randn(B, 3, H, W)
    dy = torch.randn(B, 3, H, W)
    plan = cpu_plan(m, B, H, W, True)
    # weight packing: one launch per gradient bucket (= backward segment), 47 layers in total
    assert len(plan.pack_ops) == 5 and len(plan.pack_bucket_ops) == 5 and len(plan.bwd_segments) == 5
    assert sum((_lib.op_params(o).n >> 1) & 0xFFFF for o in plan.pack_ops) == 47
    y = run_forward(plan, x)
    ref.train()
    y_ref = ref(x)
    assert rel(y, y_ref.detach()) < 1e-4        # tiny-batch BN in the 1x3 bottleneck amplifies fp32 rounding
    sd, sd_ref = m.state_dict(), ref.state_dict()
    for k in sd_ref:
        if "running_" in k:
            assert torch.allclose(sd[k], sd_ref[k], atol=2e-5, rtol=1e-4), k
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == 1
    y_ref.backward(dy)
    _lib.op_params(plan.bwd_segments[0].array[plan.dy_op_index]).src = dy.data_ptr()
    for seg in plan.bwd_segments:
        I.run_ops(seg)
    errs = {}
    for n, p in ref.named_parameters():
        off = m._grad_offsets[n]
        errs[n] = rel(m._grad_arena[off:off + p.numel()].view(p.shape), p.grad)
    # exact up to ReLU-mask flips (a single flipped element moves a layer's gradient by ~1e-3)
    assert errs["segmentation_head.0.weight"] < 1e-5 and errs["segmentation_head.0.bias"] < 1e-5
    assert max(errs.values()) < 2e-2, max(errs, key=errs.get)
    vals = sorted(errs.values())
    assert vals[len(vals) // 2] < 5e-3


def test_eval_plan_matches_oracle():
    ref, m = make_pair(seed=1)
    ref.train()
    with torch.no_grad():
        ref(torch.randn(4, 3, 64, 64))
    m.load_state_dict(ref.state_dict())
    ref.eval()
    x = torch.randn(2, 3, 64, 32)
    plan = cpu_plan(m, 2, 64, 32, False)
    assert sum(1 for op in plan.fwd_ops if op.kind in (_lib.OP_CONV, _lib.OP_CONV_BN)) == 47
    n_join = sum(1 for op in plan.fwd_ops if op.kind == _lib.OP_JOIN)   # branch-lane joins of the 3 downsample blocks
    n_upcat = sum(1 for op in plan.fwd_ops if op.kind == _lib.OP_UPCAT)   # blocks 3-4 (slab path) + blocks 0-2 (TMA-fed generic path)
    assert n_upcat in (2, 5)
    assert len(plan.fwd_ops) - n_join - n_upcat == 49               # 47 convs (BN folded) + layout + maxpool
    assert n_join in (0, 3)
    y = run_forward(plan, x)
    with torch.no_grad():
        assert rel(y, ref(x)) < 1e-5


def test_module_contract_on_cpu():
    ref, m = make_pair()
    assert set(m.state_dict()) == set(ref.state_dict())
    assert sum(p.numel() for p in m.parameters()) == 24436659
    m2 = copy.deepcopy(m)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))
    assert m2._plans == {} and m2 is not m
    with pytest.raises(ValueError):
        d3.Unet(encoder_name="resnet50")
    with pytest.raises(RuntimeError):
        cpu_plan(m, 1, 48, 64, False)
    # init distributions follow torchvision (encoder) / smp (decoder, head)
    fresh = d3.Unet()
    w = fresh.encoder.layer1[0].conv1.weight
    assert abs(w.std().item() - (2.0 / (64 * 9)) ** 0.5) < 5e-3       # kaiming_normal_(fan_out)
    assert float(fresh.segmentation_head[0].bias.abs().max()) == 0.0


def test_pack_buckets_follow_backward_segments():
    """train.StepOverlap re-packs bucket i (and updates its master weights) while backward segments > i still run:
    no op of a later segment may read a packed operand of bucket i, nor write a gradient outside buckets >= its own."""
    _, m = make_pair()
    plan = cpu_plan(m, 2, 64, 64, True)
    buckets = m.grad_buckets()
    packed_of_bucket = []
    for ol in plan.pack_bucket_ops:
        p = _lib.op_params(ol.array[0])
        n = (p.n >> 1) & 0xFFFF
        table = (_lib.PackParams * n).from_address(p.p0)
        packed_of_bucket.append({t.w_fwd for t in table} | {t.w_dgrad for t in table if t.w_dgrad})
    assert sum(len(s) for s in packed_of_bucket) >= 47
    base = m._grad_arena.data_ptr()
    for j, seg in enumerate(plan.bwd_segments):
        for op in seg:
            p = _lib.op_params(op)
            if op.kind == _lib.OP_CONV:
                for i in range(j):
                    assert p.w not in packed_of_bucket[i], f"segment {j} reads a packed weight of bucket {i}"
            written = []
            if op.kind == _lib.OP_WGRAD:
                written.append(p.dw)
            if op.kind == _lib.OP_WGRAD_GROUP:
                written += [p.dw[i] for i in range(p.count)]
            if op.kind in (_lib.OP_BN_BWD, _lib.OP_BN_BWD_APPLY, _lib.OP_BN_BWD_FINALIZE):
                written += [q for q in (p.dgamma, p.dbeta) if q]
            for q in written:
                off = (q - base) // 4
                assert buckets[j][0] <= off < buckets[j][1], f"segment {j} writes a gradient of another bucket ({off})"


def test_bf16_plan_structure_space_to_depth_stem_and_materialised_concats():
    """Host logic only (no kernel runs): the bf16 plans route the stem through the space-to-depth operands (forward AND weight
    gradient, conv / wgrad mode 2), pack its weights in the stem's own gradient bucket, read the caller's input through ONE
    layout op, and materialise upsample + concat for all five decoder blocks; a size whose stem tile is not a box keeps the
    gather-form stem."""
    m = d3.Unet(precision="bf16")
    m._ensure_grad_arena(torch.device("cpu"))
    plan = UnetPlan(dict(m.named_parameters()), dict(m.named_buffers()), 2, 64, 64, _lib.BF16, "cpu", True,
                    grad_arena=m._grad_arena, grad_offsets=m._grad_offsets)
    assert plan.stem_s2d and len(plan.in_op_indices) == 1
    kinds = [op.kind for op in plan.fwd_ops]
    assert kinds.count(_lib.OP_NCHW2S2D) == 1 and kinds.count(_lib.OP_NCHW2NHWC) == 0 and kinds.count(_lib.OP_UPCAT) == 5
    stem = next(op for op in plan.fwd_ops if op.kind == _lib.OP_CONV_BN)
    cv = _lib.op_params(stem).conv
    assert (cv.mode, cv.c0, cv.ld0, cv.kh, cv.kw, cv.stride, cv.pad, cv.Hi, cv.Wi, cv.Ho, cv.Wo, cv.Cout) == (2, 64, 16, 4, 1, 1, 2, 32, 32, 32, 32, 64)
    assert cv.src0 == plan.xs2d.data_ptr() and cv.w == plan.w_stem_s2d.data_ptr() and tuple(plan.xs2d.shape) == (2, 32, 35, 16)
    wg = [op for seg in plan.bwd_segments for op in seg if op.kind == _lib.OP_WGRAD]
    stem_wg = [op for op in wg if _lib.op_params(op).mode == 2]
    assert len(stem_wg) == 1 and _lib.op_params(stem_wg[0]).cin_real == 3 and stem_wg[0] is not None
    assert _lib.op_params(stem_wg[0]).dw == m._grad_arena.data_ptr() + 4 * m._grad_offsets["encoder.conv1.weight"]
    # the stem's pack op lives in the last bucket (layer1 + stem), behind that bucket's pack_all
    per_bucket = [[op.kind for op in ol] if ol is not None else [] for ol in plan.pack_bucket_ops]
    assert [k.count(_lib.OP_PACK_STEM) for k in per_bucket] == [0, 0, 0, 0, 1] and per_bucket[-1][0] == _lib.OP_PACK_ALL
    # eval plan: same stem, one layout op, no weight gradient
    plan_e = UnetPlan(dict(m.named_parameters()), dict(m.named_buffers()), 2, 64, 64, _lib.BF16, "cpu", False)
    assert plan_e.stem_s2d and [op.kind for op in plan_e.fwd_ops].count(_lib.OP_NCHW2S2D) == 1
    assert plan_e.in_op_index == plan_e.in_op_indices[0] and plan_e.bwd_segments is None
    # 96 x 96: the stem's 128-pixel tile is not a (w, h, n) box -> gather form on the 8-channel NHWC image
    plan_g = UnetPlan(dict(m.named_parameters()), dict(m.named_buffers()), 1, 96, 96, _lib.BF16, "cpu", False)
    assert not plan_g.stem_s2d and [op.kind for op in plan_g.fwd_ops].count(_lib.OP_NCHW2NHWC) == 1
    assert _lib.op_params(next(op for op in plan_g.fwd_ops if op.kind == _lib.OP_CONV)).mode == 0


def test_slab_weight_gradient_contract_halves_and_alignment():
    """What the slab weight gradient's coalesced epilogue (csrc/wgrad_tc.cu) relies on, checked on the host: every gradient of
    the flat arena starts on a 16-byte boundary (so dW rows take the 16-byte `red.global.add.v4.f32` form), and the 128 -> 32
    channel layer of decoder block 3 is emitted as two 64-channel launches that accumulate into disjoint channel ranges of ONE
    dW (cin_real = 128 for both, dW offset 64 * 9 floats, source offset 64 channels, pixel stride unchanged)."""
    m = d3.Unet(precision="bf16")
    m._ensure_grad_arena(torch.device("cpu"))
    assert all(off % 4 == 0 for off in m._grad_offsets.values()) and m._grad_arena.data_ptr() % 16 == 0
    B = 76                                   # 76 * 32 * 32 = 77 824 pixels >= 128 * 148 * 4: the slab path's threshold
    plan = UnetPlan(dict(m.named_parameters()), dict(m.named_buffers()), B, 64, 64, _lib.BF16, "cpu", True,
                    grad_arena=m._grad_arena, grad_offsets=m._grad_offsets)
    dw0 = m._grad_arena.data_ptr() + 4 * m._grad_offsets["decoder.blocks.3.conv1.0.weight"]
    halves = [_lib.op_params(op) for seg in plan.bwd_segments for op in seg
              if op.kind == _lib.OP_WGRAD and dw0 <= _lib.op_params(op).dw < dw0 + 4 * 32 * 128 * 9]
    assert len(halves) == 2
    a, b = sorted(halves, key=lambda p: p.dw)
    assert (a.c0, b.c0, a.c1, b.c1, a.cin_real, b.cin_real, a.ld0, b.ld0) == (64, 64, 0, 0, 128, 128, 128, 128)
    assert a.dw == dw0 and b.dw - a.dw == 64 * 9 * 4 and b.src0 - a.src0 == 64 * 2 and a.dy == b.dy
    assert (a.Cout, a.cout_real, a.kh, a.kw, a.stride, a.pad, a.Hi, a.Wi) == (32, 32, 3, 3, 1, 1, 32, 32)
    assert a.dw % 16 == 0 and b.dw % 16 == 0
