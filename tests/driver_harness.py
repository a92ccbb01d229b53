"""Harness that runs the reference's UNCHANGED drivers (search.py / train.py / prediction.py) end to
end on a tiny synthetic dataset (SURVEY.md App. F shims + App. I recipe):

  * an in-memory stand-in for h5py (absent in this image): File(path, mode) -> {subject: {name: array}}
  * stubs for nibabel / nilearn.image (absent; augmentation is off in the shipped config)
  * nas_3d_unet_b200.compat.install()  (np.int, tqdm.notebook, ReduceLROnPlateau(verbose), torch.load)
  * a working directory with config.yml (the reference's keys, small shapes), data/*.pkl, data/affine.npy
  * sys.path = [<impl dir>, <reference dir>]: with impl = nas_3d_unet_b200/dropin the drivers import
    THIS package's nas / searched / loss / cell / prim_ops / genotype; with impl = None they import
    the reference's own modules (CPU self-check of the harness).

The reference directory is /root/reference where it exists, else the bytecode compiled from it by
oracle/build_ref.py (oracle/_ref, which travels to the GPU box).
"""
import contextlib
import os
import pickle
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODULES = ("search", "train", "prediction", "generator", "patches", "augment", "helper", "adabound",
               "nas", "searched", "loss", "cell", "prim_ops", "genotype", "plot", "preprocess")


def reference_dir():
    sys.path.insert(0, ROOT) if ROOT not in sys.path else None
    from oracle import build_ref
    return build_ref.reference_dir()


class FakeH5Store:
    """h5py.File stand-in over one dict per path"""

    def __init__(self):
        self.files = {}

    def module(self):
        store = self

        class File:
            def __init__(self, path, mode="r"):
                self.path = path
                if "w" in mode:
                    store.files[path] = {}
                if path not in store.files:
                    raise OSError("no such (fake) h5 file: %s" % path)
                self.d = store.files[path]

            def __enter__(self):
                return self.d

            def __exit__(self, *exc):
                return False

        mod = types.ModuleType("h5py")
        mod.File = File
        return mod


def synthetic_subjects(n=3, shape=(48, 48, 40), seed=0):
    """SURVEY App. H at small scale: ellipsoid brain, U(10,110) inside, 0 outside; nested tumour
    blobs 2 > 1 > 4; brain_width = bounding box of the brain"""
    rng = np.random.default_rng(seed)
    out = {}
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    c = [(s - 1) / 2 for s in shape]
    brain = (((zz - c[0]) / (shape[0] * 0.42)) ** 2 + ((yy - c[1]) / (shape[1] * 0.45)) ** 2
             + ((xx - c[2]) / (shape[2] * 0.4)) ** 2) < 1.0
    nz = np.nonzero(brain)
    bw = np.array([[a.min() for a in nz], [a.max() for a in nz]])
    for i in range(n):
        sub = {}
        for mod in ("t1", "t1ce", "flair", "t2"):
            sub["BraTS19_%03d_%s.nii.gz" % (i, mod)] = (
                (rng.random(shape, dtype=np.float32) * 100 + 10) * brain).astype(np.float32)
        centre = np.array(c) + rng.integers(-4, 5, size=3)
        r2 = (zz - centre[0]) ** 2 + (yy - centre[1]) ** 2 + (xx - centre[2]) ** 2
        seg = np.zeros(shape, dtype=np.int16)
        seg[r2 < 11 ** 2] = 2
        seg[r2 < 7 ** 2] = 1
        seg[r2 < 4 ** 2] = 4
        sub["BraTS19_%03d_seg.nii.gz" % i] = seg * brain
        sub["brain_width"] = bw.copy()
        out["BraTS19_%03d" % i] = sub
    return out


CONFIG = """data:
  affine_file: data/affine.npy
  all_mods: [t1, t1ce, flair, t2]
  aug_distort: 0.25
  augment: false
  augment_distortion_factor: 0.25
  augment_flip: true
  batch_size_train: 1
  batch_size_val: 1
  cross_val_indices: data/cross_val_indices.pkl
  img_shape: [%(d)d, %(h)d, %(w)d]
  labels: [1, 2, 4]
  mean_std_file: data/mean_std.pkl
  patch_overlap: 8
  permute: true
  skip_health: true
  spe_file: data/spe.pkl
  testing_h5: data/testing.h5
  training_h5: data/training.h5
  validation_h5: data/validation.h5
  inclusive_label: true
  both_ps: false
predict:
  output_folder: data/predicted
search:
  patch_shape: 32
  best_geno_count: 40
  channel_change: true
  depth: 4
  epochs: 1
  geno_file: log/best_genotype.pkl
  gpu: true
  grad_clip: 5
  init_n_kernels: 4
  last_save: log/last_search.pt
  best_shot: log/best_search.pt
  log_path: log
  multi_gpus: false
  n_nodes: 3
  normal_w_share: false
train:
  patch_shape: 32
  best_shot: log/best_train.pt
  epochs: 1
  last_save: log/last_train.pt
"""


@contextlib.contextmanager
def driver_environment(workdir, impl_dir, shape=(48, 48, 40), n_subjects=3):
    """everything the unchanged drivers need, torn down afterwards (sys.path, sys.modules, cwd)"""
    ref = reference_dir()
    if ref is None:
        raise RuntimeError("no reference drivers: neither /root/reference nor oracle/_ref exists")
    from nas_3d_unet_b200 import compat
    store = FakeH5Store()
    store.files["data/training.h5"] = synthetic_subjects(n_subjects, shape)
    saved_nifti = {}

    nib = types.ModuleType("nibabel")

    class Nifti1Image:
        def __init__(self, data, affine):
            self.data, self.affine = np.asarray(data), affine

        def to_filename(self, path):
            saved_nifti[path] = self.data
    nib.Nifti1Image = Nifti1Image
    nilearn = types.ModuleType("nilearn")
    nilearn_image = types.ModuleType("nilearn.image")
    nilearn_image.new_img_like = nilearn_image.resample_to_img = None
    nilearn.image = nilearn_image

    saved_modules = {k: sys.modules.get(k) for k in REF_MODULES + ("h5py", "nibabel", "nilearn", "nilearn.image")}
    saved_path = list(sys.path)
    cwd = os.getcwd()
    os.makedirs(os.path.join(workdir, "data", "predicted"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "log"), exist_ok=True)
    with open(os.path.join(workdir, "config.yml"), "w") as f:
        f.write(CONFIG % dict(d=shape[0], h=shape[1], w=shape[2]))
    with open(os.path.join(workdir, "data", "cross_val_indices.pkl"), "wb") as f:
        pickle.dump({"train_list_0": list(range(n_subjects - 1)), "val_list_0": [n_subjects - 1]}, f)
    np.save(os.path.join(workdir, "data", "affine.npy"), np.eye(4))
    try:
        for k in REF_MODULES:
            sys.modules.pop(k, None)
        sys.modules["h5py"] = store.module()
        sys.modules["nibabel"] = nib
        sys.modules["nilearn"] = nilearn
        sys.modules["nilearn.image"] = nilearn_image
        sys.path[:] = ([impl_dir] if impl_dir else []) + [ref] + [p for p in saved_path if p not in ("", ".")]
        compat.install()
        os.chdir(workdir)
        yield types.SimpleNamespace(store=store, nifti=saved_nifti, reference=ref)
    finally:
        os.chdir(cwd)
        compat.uninstall()
        sys.path[:] = saved_path
        for k, v in saved_modules.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def run_drivers(env, seed=0):
    """Searching().search() -> Training().main_run() -> Prediction().predict(): the three entry
    points of the reference, verbatim.  Returns what a user would look at afterwards."""
    import random

    import torch
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    import search
    import train
    import prediction
    s = search.Searching(jupyter=True)
    model_module = type(s.model).__module__
    model_file = sys.modules[model_module].__file__
    gene = s.search()
    search_hist = dict(s.history)
    # a second Searching() resumes from log/last_search.pt (check_resume, search.py:108-127)
    os.remove("log/best_genotype.pkl")
    s2 = search.Searching(jupyter=True)
    resumed_epoch = s2.epoch
    with open("log/best_genotype.pkl", "wb") as f:
        pickle.dump(gene, f)
    t = train.Training(jupyter=True)
    t.main_run()
    train_hist = dict(t.history)
    t2 = train.Training(jupyter=True)
    p = prediction.Prediction(jupyter=True)
    p.predict("data/training.h5")
    return types.SimpleNamespace(gene=gene, search_history=search_hist, resumed_search_epoch=resumed_epoch,
                                 train_history=train_hist, resumed_train_epoch=t2.epoch,
                                 nifti=dict(env.nifti), model_file=model_file,
                                 files=sorted(os.listdir("log")))
