"""checkpoints under the reference's variable names + the serving-side feature normalisation."""
import os

import numpy as np
import torch

from conftest import make_plan


class _Opt(object):
    def __init__(self, store):
        self.t = 7
        self.m_dense = torch.randn_like(store.dense)
        self.v_dense = torch.rand_like(store.dense)
        self.m_tab = {k: torch.randn_like(v) for k, v in store.tables.items()}
        self.v_tab = {k: torch.rand_like(v) for k, v in store.tables.items()}


def test_save_load_roundtrip_with_tf_names_and_done_marker(tmp_path):
    from cikm2020_dmt_b200 import checkpoint as CK
    from cikm2020_dmt_b200.params import ParamStore
    conf, plan = make_plan("dmt_d64.conf")
    store = ParamStore(plan, device="cpu", seed=1).randomize_(2)
    opt = _Opt(store)
    d = str(tmp_path / "ckpt")
    assert CK.latest(d) is None
    path = CK.save(d, 150, store, optimizer=opt, extra={"train_calls": 12})
    CK.save(d, 100, store)
    assert os.path.exists(os.path.join(d, "step-150.model.DONE")) and CK.latest(d) == 150      # run_dnn.py:385
    with np.load(path) as z:
        keys = set(z.files)
    assert "DnnModel/embedding_trans/Sku/embedding" in keys and "DnnModel/layer_bias0/kernel" in keys
    assert "DnnModel/mmoe_layers/expert-0/expert-layer-0/weights/Adam_1" in keys             # TF slot names
    other = ParamStore(plan, device="cpu", seed=9)
    opt2 = _Opt(other)
    extra = CK.load(d, 150, other, optimizer=opt2)
    for name, v in store.named_parameters():
        assert torch.equal(v, other[name]), name
    assert opt2.t == 7 and torch.equal(opt2.m_dense[:1000], opt.m_dense[:1000])
    for k in store.tables:
        assert torch.equal(opt2.v_tab[k], opt.v_tab[k])
    assert int(extra["train_calls"]) == 12
    # an incomplete checkpoint (no marker) is ignored
    os.remove(os.path.join(d, "step-150.model.DONE"))
    assert CK.latest(d) == 100


def test_sharded_tables_reassemble(tmp_path):
    from cikm2020_dmt_b200 import checkpoint as CK
    from cikm2020_dmt_b200.params import ParamStore
    conf, plan = make_plan("dmt_d64.conf")
    full = ParamStore(plan, device="cpu", seed=3)
    name = plan.tables["Sku"].scope
    rows = full.tables[name].shape[0]
    half = (rows + 1) // 2
    d = str(tmp_path / "ck")
    for rank, (lo, hi) in enumerate([(0, half), (half, rows)]):
        part = ParamStore(plan, device="cpu", seed=3, row_shards={name: (lo, hi)})
        CK.save(d, 5, part, shards={name: (lo, hi)}, rank=rank)
    # a single-GPU reader gets the whole table back, a 2-rank reader its own shard
    one = ParamStore(plan, device="cpu", seed=8)
    CK.load(d, 5, one)
    assert torch.equal(one.tables[name], full.tables[name])
    part1 = ParamStore(plan, device="cpu", seed=8, row_shards={name: (half, rows)})
    CK.load(d, 5, part1, shards={name: (half, rows)}, rank=1)
    assert torch.equal(part1.tables[name], full.tables[name][half:])


def test_sharded_tables_resume_with_adam_slots(tmp_path):
    """Data-parallel exact resume: the Adam slots of a row-sharded table are cut by the variable's row range
    (2 ranks save with optimizer; each rank, and a single-GPU reader, gets its own rows of m / v back), and the
    model's cached weight images are invalidated by `load`."""
    from cikm2020_dmt_b200 import checkpoint as CK
    from cikm2020_dmt_b200.params import ParamStore
    conf, plan = make_plan("dmt_d64.conf")
    name = plan.tables["Sku"].scope
    full = ParamStore(plan, device="cpu", seed=3)
    rows = full.tables[name].shape[0]
    half = (rows + 1) // 2
    d = str(tmp_path / "ck")
    opts = []
    for rank, (lo, hi) in enumerate([(0, half), (half, rows)]):
        part = ParamStore(plan, device="cpu", seed=3, row_shards={name: (lo, hi)})
        torch.manual_seed(100 + rank)
        opt = _Opt(part)
        opts.append(opt)
        CK.save(d, 9, part, optimizer=opt, shards={name: (lo, hi)}, rank=rank)

    class _Model(object):
        invalidated = 0

        def invalidate_prepared(self):
            self.invalidated += 1

    for rank, (lo, hi) in enumerate([(0, half), (half, rows)]):
        part = ParamStore(plan, device="cpu", seed=8, row_shards={name: (lo, hi)})
        opt = _Opt(part)
        opt.model = _Model()
        CK.load(d, 9, part, optimizer=opt, shards={name: (lo, hi)}, rank=rank)
        assert opt.t == 7 and opt.model.invalidated == 1
        assert torch.equal(part.tables[name], full.tables[name][lo:hi])
        assert torch.equal(opt.m_tab[name], opts[rank].m_tab[name])
        assert torch.equal(opt.v_tab[name], opts[rank].v_tab[name])
    one = ParamStore(plan, device="cpu", seed=8)
    opt = _Opt(one)
    m = _Model()
    CK.load(d, 9, one, optimizer=opt, model=m)
    assert m.invalidated == 1
    assert torch.equal(opt.m_tab[name], torch.cat([opts[0].m_tab[name], opts[1].m_tab[name]], 0))
    assert torch.equal(opt.v_tab[name], torch.cat([opts[0].v_tab[name], opts[1].v_tab[name]], 0))


def test_serving_feature_normalisation_formula():
    """export_model.py:88-96 / preprocess.py:17-43 restated in numpy fp64."""
    from cikm2020_dmt_b200 import checkpoint as CK
    rng = np.random.default_rng(1)
    mean, std = rng.random(615) * 5, rng.random(615) * 3
    std[:5] = 0.0
    x = torch.from_numpy(rng.normal(1.0, 2.0, (7, 615)).astype(np.float32))
    got = CK.serving_features(x, mean, std).double().numpy()
    eps = 1e-7
    c = mean * std / ((std + eps) ** 2 * 3) + mean * std / (std + eps) - mean
    want = np.clip(np.clip(x.double().numpy(), 0, None) * std / ((std + eps) ** 2 * 3.0) - c, -0.99, 0.99)
    assert np.allclose(got, want, atol=2e-5)
    assert got.min() >= -0.99 - 1e-6 and got.max() <= 0.99 + 1e-6      # fp32(0.99)
