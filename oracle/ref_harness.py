"""Import harness for the REAL reference (AmazingDD/daisyRec at /root/reference).

TEST INFRASTRUCTURE ONLY.  Nothing under ``daisyrec_b200/`` may import this file.
Its jobs: (1) validating the restatements in ``oracle/`` against the reference itself,
(2) generating the golden fixtures in ``tests/golden/`` via ``oracle/gen_golden.py`` (both in
the build container, from ``/root/reference``), and (3) letting ``bench.py --impl reference`` /
the ``cpu_baseline`` leg time the reference's own ``fit`` on the GPU box's host cores from the
installed copy ``oracle/_ref`` (``oracle/build_ref.py``).

The reference does not run unmodified on numpy 2.3 / pandas 3.0 / scipy 1.18
(SURVEY.md facts 10, Appendix B).  Three behavioural shims are applied *here*, in
process, before ``daisy.*`` is imported -- the reference tree is never edited:

1. pandas>=2: ``Series.agg(callable)`` stopped mapping element-wise
   (daisy/utils/sampler.py:91 relies on it)            -> fall back to ``Series.map``.
2. numpy>=2:  ``np.asfarray`` removed (daisy/utils/metrics.py:206).
3. scipy:     ``dok_matrix._update`` removed (daisy/model/LightGCNRecommender.py:89).

``daisy.utils.config`` needs colorlog/colorama (absent) so the config dict is
assembled by hand from the YAML assets exactly as ``init_config`` does
(daisy/utils/config.py:48-75).
"""
import logging
import os
import random
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.join(_HERE, "_ref")          # oracle/build_ref.py: the reference package, installed unmodified


def _default_root():
    env = os.environ.get("DAISY_REF_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/daisy"):         # build container: the source tree (has data/ml-100k too)
        return "/root/reference"
    return INSTALLED_ROOT                              # GPU box: only the installed copy travels


REF_ROOT = _default_root()


def use_root(path):
    """Point the harness at another copy of the reference (bench.py times the installed copy everywhere)."""
    global REF_ROOT
    REF_ROOT = path


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "daisy"))


def _apply_shims():
    import numpy as np
    import pandas as pd
    import scipy.sparse as sp

    if not getattr(pd.Series.agg, "_drb_shim", False):
        _agg = pd.Series.agg

        def agg(self, func=None, *a, **k):
            if callable(func):
                return self.map(func)
            return _agg(self, func, *a, **k)

        agg._drb_shim = True
        pd.Series.agg = agg
    if not hasattr(np, "asfarray"):
        np.asfarray = lambda a: np.asarray(a, dtype=np.float64)
    if not hasattr(sp.dok_matrix, "_update"):
        def _upd(self, d):
            for key, v in d.items():
                self[key] = v
        sp.dok_matrix._update = _upd


def import_reference():
    """Put the reference on sys.path (after the shims) and return the ``daisy`` package."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _apply_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import daisy  # noqa: F401
    return daisy


def make_config(algo="mf", **overrides):
    """basic.yaml (+) <algo>.yaml (+) overrides, the same merge as daisy/utils/config.py:48-75.

    Mirrors SURVEY fact 9: the real CLI yields early_stop=False, init_method='default'.
    """
    import yaml
    assets = os.path.join(REF_ROOT, "daisy", "assets")
    cfg = yaml.safe_load(open(os.path.join(assets, "basic.yaml")))
    cfg.update(yaml.safe_load(open(os.path.join(assets, f"{algo}.yaml"))) or {})
    cfg["algo_name"] = algo
    cfg["early_stop"] = False
    cfg["gpu"] = ""            # force CPU (AbstractRecommender.py:99 overwrites CUDA_VISIBLE_DEVICES)
    cfg["data_path"] = os.path.join(REF_ROOT, "data") + "/"
    cfg.update(overrides)
    logger = logging.getLogger("daisy_ref")
    logger.setLevel(logging.WARNING)
    cfg["logger"] = logger
    return cfg


def seed_everything(seed):
    """daisy/utils/config.py:32-36 (CPU part)."""
    import numpy as np
    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def load_ml100k(cfg):
    """test.py:55-71: read -> preprocess -> tsbr split -> get_ur.  Returns a dict of artefacts."""
    import_reference()
    from daisy.utils.loader import RawDataReader, Preprocessor
    from daisy.utils.splitter import TestSplitter
    from daisy.utils.utils import get_ur
    reader, proc = RawDataReader(cfg), Preprocessor(cfg)
    df = proc.process(reader.get_data())
    cfg["user_num"], cfg["item_num"] = proc.user_num, proc.item_num
    tr_idx, te_idx = TestSplitter(cfg).split(df)
    train_set, test_set = df.iloc[tr_idx, :].copy(), df.iloc[te_idx, :].copy()
    test_ur, train_ur = get_ur(test_set), get_ur(train_set)
    cfg["train_ur"] = train_ur
    return dict(df=df, train_set=train_set, test_set=test_set, test_ur=test_ur, train_ur=train_ur)
