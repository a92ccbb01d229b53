"""Compatibility layer for running the reference DRIVERS (search.py / train.py / prediction.py)
unchanged on this image.  Nothing here is on the compute path.

    from nas_3d_unet_b200 import compat
    compat.install()            # before importing the reference's search / train / prediction modules

install() applies, idempotently:
  * `np.int` alias (patches.py:75,188,194 use it; removed in numpy 1.24).  `np.bool` is left alone.
  * `tqdm.notebook.tqdm` -> plain `tqdm.tqdm` when ipywidgets is absent (search.py:17).
  * `torch.optim.lr_scheduler.ReduceLROnPlateau` accepting and dropping `verbose=`
    (search.py:105-106, train.py:50; the kwarg was removed from torch).
  * `torch.load` defaulting to `weights_only=False` (search.py:112, train.py:57, prediction.py:56
    un-pickle Counter / defaultdict / Genotype objects).
  * optionally `torch.optim.Adam` -> optim.FlatAdam (`flat_adam=True`) so the drivers' optimiser
    lines pick up the one-launch update without an edit.
The geno_file of the drivers - a pickle of (str(Genotype), count), search.py:190-195, read back with
eval() in train.py:38 - goes through load_genotype / save_genotype / parse_genotype (no eval).
"""
import functools
import pickle
import sys
import types

import torch

_installed = {}


def _make_plateau(base):
    class ReduceLROnPlateau(base):
        """torch's scheduler, tolerant of the removed `verbose` argument"""

        def __init__(self, optimizer, *args, verbose=None, **kwargs):
            if len(args) > 5:      # old positional order: mode, factor, patience, verbose, threshold, ...
                args = args[:3] + args[4:]
            super().__init__(optimizer, *args, **kwargs)
            self.verbose = bool(verbose)

        def step(self, metrics, *a, **k):
            before = [g["lr"] for g in self.optimizer.param_groups]
            super().step(metrics, *a, **k)
            if self.verbose:
                for i, (b, g) in enumerate(zip(before, self.optimizer.param_groups)):
                    if g["lr"] != b:
                        print("Epoch %5d: reducing learning rate of group %d to %.4e."
                              % (self.last_epoch, i, g["lr"]))
    ReduceLROnPlateau.__qualname__ = "ReduceLROnPlateau"
    return ReduceLROnPlateau


def install(flat_adam=False):
    import numpy as np
    if not hasattr(np, "int"):
        np.int = int
        _installed["np.int"] = True
    try:
        import ipywidgets  # noqa: F401
    except ImportError:
        import tqdm
        mod = sys.modules.get("tqdm.notebook")
        if mod is None or getattr(mod, "_nas3d_shim", False) is False:
            shim = types.ModuleType("tqdm.notebook")
            shim.tqdm = tqdm.tqdm
            shim.trange = tqdm.trange
            shim._nas3d_shim = True
            sys.modules["tqdm.notebook"] = shim
            tqdm.notebook = shim
            _installed["tqdm.notebook"] = True
    sched = torch.optim.lr_scheduler
    if not getattr(sched.ReduceLROnPlateau, "_nas3d_shim", False):
        cls = _make_plateau(sched.ReduceLROnPlateau)
        cls._nas3d_shim = True
        _installed["ReduceLROnPlateau"] = sched.ReduceLROnPlateau
        sched.ReduceLROnPlateau = cls
    if not getattr(torch.load, "_nas3d_shim", False):
        orig = torch.load

        @functools.wraps(orig)
        def load(*args, **kwargs):
            kwargs.setdefault("weights_only", False)
            return orig(*args, **kwargs)
        load._nas3d_shim = True
        _installed["torch.load"] = orig
        torch.load = load
    if flat_adam and not getattr(torch.optim.Adam, "_nas3d_shim", False):
        from .optim import FlatAdam
        _installed["Adam"] = torch.optim.Adam
        FlatAdam._nas3d_shim = True
        torch.optim.Adam = FlatAdam
    return sorted(_installed)


def uninstall():
    """undo install() (tests)"""
    import numpy as np
    if _installed.pop("np.int", None) and hasattr(np, "int"):
        del np.int
    if "ReduceLROnPlateau" in _installed:
        torch.optim.lr_scheduler.ReduceLROnPlateau = _installed.pop("ReduceLROnPlateau")
    if "torch.load" in _installed:
        torch.load = _installed.pop("torch.load")
    if "Adam" in _installed:
        torch.optim.Adam = _installed.pop("Adam")
    if _installed.pop("tqdm.notebook", None):
        sys.modules.pop("tqdm.notebook", None)


def parse_genotype(text):
    """'Genotype(down=[(name, edge), ...], up=[...])' -> Genotype, without eval() (the drivers do
    `eval(pickle.load(f)[0])`, train.py:38, prediction.py:48)"""
    import ast
    from .genotype import Genotype
    node = ast.parse(text.strip(), mode="eval").body
    if not (isinstance(node, ast.Call) and getattr(node.func, "id", None) == "Genotype" and not node.args):
        raise ValueError("not a Genotype(...) expression: %.60r" % text)
    kw = {k.arg: ast.literal_eval(k.value) for k in node.keywords}
    if set(kw) != {"down", "up"}:
        raise ValueError("Genotype needs exactly down= and up=")
    return Genotype(down=[(str(n), int(e)) for n, e in kw["down"]],
                    up=[(str(n), int(e)) for n, e in kw["up"]])


def load_genotype(path):
    """read the reference's geno_file: a pickle of (str(Genotype), count) (search.py:190-195);
    returns (Genotype, count)"""
    with open(path, "rb") as f:
        text, count = pickle.load(f)
    return parse_genotype(text), int(count)


def save_genotype(gene, path, count=1):
    """write a geno_file the unchanged drivers can read back (train.py:37-38)"""
    from .genotype import Genotype
    text = str(Genotype(down=[(str(n), int(e)) for n, e in gene.down],
                        up=[(str(n), int(e)) for n, e in gene.up]))
    with open(path, "wb") as f:
        pickle.dump((text, int(count)), f)
