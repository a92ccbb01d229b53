"""The reference's own drivers, unchanged, end to end ("search.py, train.py and
prediction.py run unchanged against it").

  * CPU (not gpu): the drivers over the REFERENCE's modules - proves the harness (fake h5py, stubs,
    compat shims, tiny config) drives the real code: Searching().search() -> Training().main_run()
    -> Prediction().predict() complete and write what the reference writes.
  * GPU: the very same harness with nas_3d_unet_b200/dropin first on sys.path, so the drivers'
    `from nas import ShellNet`, `from searched import SearchedNet`, `from loss import
    WeightedDiceLoss` resolve to THIS package and every step runs on the sm_100a kernels.
"""
import os

import numpy as np
import pytest

from conftest import ROOT
import driver_harness as H

needs_reference = pytest.mark.skipif(H.reference_dir() is None,
                                     reason="reference drivers unavailable (no /root/reference, no oracle/_ref)")


def _check(res):
    gene_text, count = res.gene
    assert gene_text.startswith("Genotype(down=[") and count >= 1
    assert res.files == ["best_genotype.pkl", "best_search.pt", "best_train.pt", "last_search.pt",
                         "last_train.pt"] or set(res.files) >= {"best_genotype.pkl", "last_search.pt",
                                                                "last_train.pt"}
    for k in ("shell_loss", "kernel_loss", "val_loss"):
        assert len(res.search_history[k]) == 1 and 0.0 <= res.search_history[k][0] <= 1.0
    for k in ("loss", "val_loss"):
        assert len(res.train_history[k]) == 1 and 0.0 <= res.train_history[k][0] <= 1.0
    assert res.resumed_search_epoch == 1 and res.resumed_train_epoch == 1      # checkpoints reload
    assert len(res.nifti) == 3
    for path, vol in res.nifti.items():
        assert path.endswith(".nii.gz") and vol.dtype == np.uint8 and vol.shape == (48, 48, 40)
        assert set(np.unique(vol)) <= {0, 1, 2, 4}


@needs_reference
def test_reference_drivers_run_on_reference_modules_cpu(tmp_path, monkeypatch):
    """harness self-check on the host: nothing of this package on the model path"""
    import torch
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)     # search.py:66 -> cpu
    with H.driver_environment(str(tmp_path), impl_dir=None) as env:
        res = H.run_drivers(env)
    assert os.path.dirname(res.model_file) == env.reference
    _check(res)


@needs_reference
@pytest.mark.gpu
def test_reference_drivers_run_unchanged_on_this_package(tmp_path):
    from nas_3d_unet_b200 import _lib
    n0 = _lib.launch_count()
    dropin = os.path.join(ROOT, "nas_3d_unet_b200", "dropin")
    with H.driver_environment(str(tmp_path), impl_dir=dropin) as env:
        res = H.run_drivers(env)
    # the drivers' `from nas import ShellNet` went through the dropin shim to THIS package's class
    assert os.path.dirname(res.model_file) == os.path.join(ROOT, "nas_3d_unet_b200"), res.model_file
    assert _lib.launch_count() - n0 > 1000, "the drivers' steps did not run on the nas3d kernels"
    _check(res)
