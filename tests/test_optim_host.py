"""Host logic of the optimizer: the reference's learning-rate schedule
`tf.train.piecewise_constant(global_step, step_boundary, learning_rate)` (run_dnn.py:119-126, dmt.conf
`learning_rate = 0.001,0.0001`, `step_boundary = 300000000`)."""
import pytest
import torch

from conftest import make_plan


class _FakeModel(object):
    """What TFAdam needs of a model on a box without a GPU: the parameter store and a plan."""

    def __init__(self, plan, store):
        self.plan, self.params, self.lib = plan, store, None
        self.launches = 0

    def invalidate_prepared(self):
        pass


def test_piecewise_constant_matches_tf_semantics_and_oracle():
    from cikm2020_dmt_b200.optim import piecewise_constant
    from oracle import dmt_oracle as O
    b, v = [10, 20], [1e-3, 1e-4, 1e-5]
    for step, want in [(0, 1e-3), (10, 1e-3), (11, 1e-4), (20, 1e-4), (21, 1e-5), (10 ** 9, 1e-5)]:
        assert piecewise_constant(step, b, v) == want == O.piecewise_constant(step, b, v)
    with pytest.raises(ValueError):
        piecewise_constant(0, [10], [1e-3])            # TF: len(values) must be len(boundaries) + 1


def test_tfadam_follows_the_conf_schedule_across_the_boundary():
    from cikm2020_dmt_b200.optim import TFAdam
    from cikm2020_dmt_b200.params import ParamStore
    conf, plan = make_plan("dmt.conf", overrides={("model", "step_boundary"): "3"})
    assert plan.learning_rate == [0.001, 0.0001] and plan.step_boundary == [3]
    store = ParamStore(plan, device="cpu", seed=1)
    opt = TFAdam(_FakeModel(plan, store), plan.learning_rate)          # boundary taken from the model's plan
    seen = []
    for _ in range(6):
        opt.begin_step()
        seen.append(opt._cfg(None).lr)
    # global_step 0..3 -> 1e-3 (values[0] while step <= boundary), then 1e-4
    assert seen == pytest.approx([1e-3] * 4 + [1e-4] * 2)
    assert opt.t == 6 and opt.global_step == 6
    # a resume past the boundary trains at the decayed rate although the Adam powers restart
    opt2 = TFAdam(_FakeModel(plan, store), plan.learning_rate, global_step=100)
    opt2.begin_step()
    assert opt2._cfg(None).lr == pytest.approx(1e-4) and opt2.t == 1
    # a list without boundaries is an error, never a silent lr[0]
    with pytest.raises(ValueError):
        TFAdam(_FakeModel(None, store), [1e-3, 1e-4])
    with pytest.raises(TypeError):
        opt._cfg([1e-3, 1e-4])


def test_shipped_conf_has_the_reference_schedule():
    conf, plan = make_plan("dmt.conf")
    assert plan.learning_rate == [0.001, 0.0001] and plan.step_boundary == [300000000]
