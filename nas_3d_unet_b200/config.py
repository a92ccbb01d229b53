"""Host-side execution options of the engine, read ONCE per process.

The defaults are the measured-best settings; the environment (NAS3D_<NAME>) is consulted a single
time at import for A/B runs of the bench, and tests change options with `override(...)`.  Nothing
on the per-call path reads the environment.  (The kernel-selection options of the C-ABI library
live in csrc/options.cu and are changed with `_lib.option(name, value)`.)
"""
import contextlib
import os


class EngineConfig:
    __slots__ = ("wgrad_stream", "wgrad_streams", "lanes", "candidate_lanes", "virtual_cat", "gn_fold",
                 "umma", "umma_min_c", "pw_fused_bwd")

    def __init__(self, env=os.environ):
        def flag(name, default):
            v = env.get(name)
            return default if v is None or v == "" else v != "0"

        def integer(name, default):
            v = env.get(name)
            return default if v is None or v == "" else int(v)

        # weight-gradient kernels on side stream(s), concurrent with the dgrad / node chain
        self.wgrad_stream = flag("NAS3D_WGRAD_STREAM", True)
        self.wgrad_streams = max(1, min(4, integer("NAS3D_WGRAD_STREAMS", 1)))
        # stream lanes the independent edges of a cell node are spread over (1 = caller's stream)
        self.lanes = max(1, min(4, integer("NAS3D_LANES", 4)))
        # supernet: the K candidate ops of a MixedOp run their FORWARD on streams of their own (measured
        # on B200: search step 26.1 -> 19.6 ms at 64^3, 42.4 -> 36.6 ms at 128^3, profiles/r3f_*)
        self.candidate_lanes = flag("NAS3D_CANDIDATE_LANES", True)
        # cell outputs stay a virtual concat of their node buffers (cell.py:82 never copies)
        self.virtual_cat = flag("NAS3D_VIRTUAL_CAT", True)
        # GroupNorm coefficient kernels folded into the affine kernels' prologues.  Off: measured
        # SLOWER on B200 (406 vs 414 patches/s, profiles/r1f_ab_gn_fold_*.json)
        self.gn_fold = flag("NAS3D_GN_FOLD", False)
        # tcgen05 path: on from 32 channels (16 for stride-2 dil-1), tools/umma_micro.py
        self.umma = not flag("NAS3D_DISABLE_UMMA", False)
        self.umma_min_c = integer("NAS3D_UMMA_MIN_C", None)
        # stride-1 1x1x1 convs: dgrad + wgrad (+ sigmoid backward) in one pass
        self.pw_fused_bwd = flag("NAS3D_PW_FUSED_BWD", True)


cfg = EngineConfig()


@contextlib.contextmanager
def override(**kw):
    """with override(gn_fold=True): ...   (tests / A-B tools only)"""
    prev = {k: getattr(cfg, k) for k in kw}
    for k, v in kw.items():
        setattr(cfg, k, v)
    try:
        yield cfg
    finally:
        for k, v in prev.items():
            setattr(cfg, k, v)
