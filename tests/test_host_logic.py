"""CPU checks of host-side scheduling arithmetic (no GPU).

* `bench.pick_inflight`: the number of in-flight contexts for a K-step timed region.
* the run / owner arithmetic of the persistent fused cross-attention (`csrc/xattn.cu`): restated here line for
  line (integer arithmetic only) and checked for the invariants the kernel's merge protocol relies on.
"""
import importlib.util
import os
import random

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(REPO, 'bench.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_pick_inflight_fills_rounds():
    b = _bench()
    for k in range(1, 65):
        c = b.pick_inflight(k)
        assert 1 <= c <= 6 and c <= k
        if k <= 12:
            tail = k % c
            assert tail == 0 or tail >= c // 2, (k, c)  # short runs: never a nearly empty last round
    assert b.pick_inflight(8) == 4 and b.pick_inflight(5) == 5 and b.pick_inflight(1) == 1


def _plan(npairs, nkb, num_sms, ctas_per_sm):
    """cross_attn_mq_plan (csrc/xattn.cu)."""
    nb = npairs * nkb
    g = min(nb, num_sms * max(1, ctas_per_sm))
    bpc = nb // g
    return g, (nkb + bpc - 1) // bpc + 1


def test_xattn_runs_partition_the_work_and_bound_the_partials():
    rnd = random.Random(0)
    cases = [(128, 64, 148, 2), (16, 4, 148, 2), (8, 1, 148, 2), (1, 1, 148, 3), (256, 64, 148, 2)]
    cases += [(rnd.randint(1, 300), rnd.randint(1, 70), rnd.choice([16, 132, 148]), rnd.randint(1, 3)) for _ in range(400)]
    for npairs, nkb, sms, cps in cases:
        nb = npairs * nkb
        g, max_parts = _plan(npairs, nkb, sms, cps)
        owner = lambda b: ((b + 1) * g - 1) // nb            # the kernel's `owner` lambda
        seen = [None] * nb
        for cta in range(g):
            lo, hi = cta * nb // g, (cta + 1) * nb // g     # the kernel's run [b, b_end)
            assert hi > lo                                  # no empty CTA: every counted CTA arrives at the merge
            for b in range(lo, hi):
                assert seen[b] is None
                seen[b] = cta
        assert all(x is not None for x in seen)
        for b in range(nb):
            assert owner(b) == seen[b]
        for p in range(npairs):
            first, last = owner(p * nkb), owner(p * nkb + nkb - 1)
            assert 1 <= last - first + 1 <= max_parts      # partial slots reserved per pair suffice
            for cta in range(first, last + 1):              # every CTA in first..last owns a block of this pair
                assert any(seen[b] == cta for b in range(p * nkb, (p + 1) * nkb))
