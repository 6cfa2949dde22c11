"""Import the UNMODIFIED reference (`/root/reference/src`) in this container.

Only `tests/golden/make_golden.py` uses this, and only here (the GPU box has no
`/root/reference`).  Nothing is copied: the reference is put on `sys.path` and
four import-time breakages of its environment pins (torch 1.11 / numpy 1.21 /
scipy 1.7, `requirements.txt`) are shimmed, as listed in SURVEY.md Appendix A:

* `lmdb`, `matplotlib` -- imported but unused on this path
  (`src/utils/dataset.py:18`, `src/common/trainer.py:14`) -> empty stub modules;
* `scipy.sparse.dok_matrix._update` -- private API removed from scipy
  (`src/models/freedom.py:111`) -> dict update;
* `np.float` -- removed alias (`src/utils/metrics.py:51`);
* `torch_scatter.scatter_add` (`src/utils/utils.py:140`, MGCN only) -> index_add_;
* MGCN hard-codes `.cuda()` (`src/models/mgcn.py:59,69`) -> identity on a CPU box;
* LayerGCN imports `models.common.*` (`src/models/layergcn.py:12-13`) -> alias.
"""
import os
import sys
import types

import numpy as np
import scipy.sparse as sp
import torch

REF_SRC = "/root/reference/src"


def install():
    if not os.path.isdir(REF_SRC):
        raise RuntimeError("reference tree not present (golden vectors are generated in the build container only)")
    for name in ("lmdb", "matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ts = types.ModuleType("torch_scatter")
    ts.scatter_add = lambda src, index, dim=0, dim_size=None: torch.zeros(
        dim_size, dtype=src.dtype, device=src.device).index_add_(0, index, src)
    sys.modules.setdefault("torch_scatter", ts)
    if not hasattr(sp.dok_matrix, "_update"):
        sp.dok_matrix._update = lambda self, d: self._dict.update({k: np.float32(v) for k, v in d.items()})
    if not hasattr(np, "float"):
        np.float = float
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import common
    import common.abstract_recommender
    import common.loss
    for m in ("", ".abstract_recommender", ".loss"):
        sys.modules["models.common" + m] = sys.modules["common" + m]


def run_dir(tmp: str) -> str:
    """`Config` resolves `./configs` against CWD (`src/utils/configurator.py:72-73`) and the data
    path is `../data/` (`src/configs/overall.yaml:7`): build `<tmp>/run/configs -> reference` and
    return `<tmp>/data`, leaving CWD in `<tmp>/run`."""
    run = os.path.join(tmp, "run")
    os.makedirs(run, exist_ok=True)
    link = os.path.join(run, "configs")
    if not os.path.exists(link):
        os.symlink(os.path.join(REF_SRC, "configs"), link)
    os.makedirs(os.path.join(tmp, "data"), exist_ok=True)
    os.chdir(run)
    return os.path.join(tmp, "data")
