import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONF_DIR = os.path.join(ROOT, "conf", "settings") + "/"
SMALL_ROWS = {"Sku": 2000, "Brand": 700, "Shopid": 900, "Cid3": 300, "Cid2": 60}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def make_plan(conf_file="dmt_d64.conf", rows=None, overrides=None):
    from cikm2020_dmt_b200.conf import Conf
    from cikm2020_dmt_b200.plan import build_plan
    conf = Conf(CONF_DIR, conf_file, overrides=overrides)
    plan = build_plan(conf)
    rows = SMALL_ROWS if rows is None else rows
    for t in list(plan.tables.values()) + list(plan.bias_tables.values()):
        if t.name in rows:
            t.rows = rows[t.name]
    return conf, plan


@pytest.fixture(scope="session")
def small_d64():
    return make_plan("dmt_d64.conf")


@pytest.fixture(scope="session")
def small_d80():
    return make_plan("dmt.conf")
