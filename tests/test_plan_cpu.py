"""Host-logic tests (no GPU): the plan compiler's op list, executed by the CPU interpreter in oracle/, must
reproduce the reference goldens within the bf16 tolerance north_star states (abs 2e-2, rel 1e-2)."""
import numpy as np
import pytest
import torch

from test_oracle_cpu import GOLDEN_CASES, compare_x_maps, golden_state_dict, load_golden

SMALL = [c for c in GOLDEN_CASES if "640" not in c]


@pytest.mark.parametrize("name", SMALL + ["n640_stress"])
def test_plan_interpreted_on_cpu_matches_reference_golden(name):
    from dcfa_b200.plan import Plan
    from oracle import forward as O
    from oracle import plan_interp
    z, meta, keys = load_golden(name)
    sd = golden_state_dict(meta, keys)
    rgb, nir = O.synth_inputs(meta["B"], meta["H"], meta["W"], meta["seed"] + 1000)
    plan = Plan(sd, meta["phi"], meta["nc"], meta["B"], meta["H"], meta["W"])
    dbox, cls, x = plan_interp.run_plan(plan, rgb, nir)
    compare_x_maps(z, x, 2e-2, 1e-2)
    np.testing.assert_allclose(cls.numpy(), z["cls"], atol=2e-2, rtol=1e-2)
    np.testing.assert_allclose(dbox.numpy(), z["dbox"], atol=2e-2, rtol=1e-2)
    assert plan.A == z["dbox"].shape[-1]


def test_plan_flop_count_matches_survey():
    """algorithmic conv FLOPs per image pair (2*MAC over every nn.Conv2d), nc=1, 640x640."""
    from dcfa_b200.plan import Plan
    from oracle import forward as O
    import json, os
    keys = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "state_dict_keys.json")))["n_nc1"]
    sd = O.synth_state_dict({k: torch.empty(tuple(s), device="meta") for k, s in keys.items()}, 1, "default")
    plan = Plan(sd, "n", 1, 1, 640, 640)
    assert abs(plan.conv_flops / 7.361e9 - 1.0) < 0.01, plan.conv_flops


def test_plan_uint8_input_matches_oracle_on_preprocessed_pixels():
    """input_u8 plans (uint8 NHWC images, /255 folded into the stem) vs the fp32 oracle fed preprocess_input's
    output (utils/utils.py:76-79 + the HWC->CHW transpose of yolo_mul.py:76)."""
    from dcfa_b200.plan import Plan
    from oracle import forward as O
    from oracle import plan_interp
    z, meta, keys = load_golden("s128_stress")
    sd = golden_state_dict(meta, keys)
    g = torch.Generator().manual_seed(3)
    rgb8 = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8)
    nir8 = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8)
    plan = Plan(sd, meta["phi"], meta["nc"], 2, 128, 128, input_u8=True)
    dbox, cls, x = plan_interp.run_plan(plan, rgb8, nir8)
    ref = O.yolo_forward(sd, meta["phi"], rgb8.permute(0, 3, 1, 2).float() / 255.0, nir8.permute(0, 3, 1, 2).float() / 255.0,
                         meta["nc"])
    np.testing.assert_allclose(dbox.numpy(), ref[0].numpy(), atol=2e-2, rtol=1e-2)
    np.testing.assert_allclose(cls.numpy(), ref[1].numpy(), atol=2e-2, rtol=1e-2)


def test_arena_liveness_reuse_is_sound_and_exact():
    """Arena addresses are shared between tensors with disjoint lifetimes (Plan._assign_arena_by_liveness).  Checked
    independently of the allocator: (1) at every op, the byte ranges of the tensors referenced within the fusion span
    around it never overlap unless they are the same tensor; (2) the interpreted plan gives bit-identical outputs with
    and without reuse; (3) the arena actually shrinks."""
    from dcfa_b200 import plan as P
    from oracle import forward as O
    from oracle import plan_interp
    z, meta, keys = load_golden("s128_stress")
    sd = golden_state_dict(meta, keys)
    rgb, nir = O.synth_inputs(meta["B"], meta["H"], meta["W"], meta["seed"] + 1000)
    args = (sd, meta["phi"], meta["nc"], meta["B"], meta["H"], meta["W"])
    bump, reuse = P.Plan(*args, reuse_arena=False), P.Plan(*args)
    assert reuse.arena_bytes < 0.6 * bump.arena_bytes and bump.arena_bytes == bump.arena_bytes_bump
    out_b, out_r = plan_interp.run_plan(bump, rgb, nir), plan_interp.run_plan(reuse, rgb, nir)
    assert torch.equal(out_b[0], out_r[0]) and torch.equal(out_b[1], out_r[1])
    for a, b in zip(out_b[2], out_r[2]):
        assert torch.equal(a, b)

    # (1): map every arena reference back to its bump allocation through the un-reused twin plan
    import bisect
    starts = [a[0] for a in bump._allocs]
    sizes = [a[1] for a in bump._allocs]
    per_op = []
    for ob, orr in zip(bump.ops, reuse.ops):
        cur = {}
        for f in P.Plan._VIEW_FIELDS:
            vb, vr = getattr(ob, f), getattr(orr, f)
            if vb.buf != P.BUF_ARENA:
                continue
            a = bisect.bisect_right(starts, vb.off) - 1
            cur[a] = vr.off - (vb.off - starts[a])
        per_op.append(cur)
    span = P.Plan._FUSION_SPAN
    first, last, where = {}, {}, {}
    for i, cur in enumerate(per_op):
        for a, o in cur.items():
            first.setdefault(a, i)
            last[a] = i
            assert where.setdefault(a, o) == o          # one address per tensor
    ids = sorted(where)
    for x in range(len(ids)):
        for y in range(x + 1, len(ids)):
            a, b = ids[x], ids[y]
            if last[a] + span < first[b] or last[b] + span < first[a]:
                continue                                 # never alive (nor inside one fused kernel) together
            assert where[a] + sizes[a] <= where[b] or where[b] + sizes[b] <= where[a], (a, b)
