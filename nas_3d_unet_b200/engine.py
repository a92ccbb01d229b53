"""Host-side executor of the B200 compute path.

The reference builds its graph out of torch.nn leaf modules and lets torch autograd replay
it (prim_ops.py:68-83, cell.py:24-33,66-82).  Here the module tree (same classes, same
state_dict) *records a tape of C-ABI kernel launches* instead: one torch.autograd.Function
wraps the outermost module call, its forward walks the tree launching the sm_100a kernels of
libnas3d_b200.so on torch's current stream, its backward replays the tape in reverse.  Torch
only supplies device memory, streams and the autograd edge to the caller.

Vocabulary
  Act   an activation: logical (N,C,D,H,W) torch tensor, physically NDHWC with voxel pitch ld
        (possibly a channel slice of a wider buffer = zero-copy concat).
  Term  a lazily evaluated per-(n,c) affine view  act(a[n,c]*x + b[n,c])  of an Act: GroupNorm
        apply, ReLU, SE channel scale and the softmax(alpha) weight all fold into a Term, and
        affine_sum() evaluates a whole list of Terms in one pass over memory.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ConvDesc, check, int_array, ptr_array
from .config import cfg


class Nas3dDeviceError(RuntimeError):
    pass


def _stream():
    return torch.cuda.current_stream().cuda_stream


_side_streams = {}


def wgrad_stream_enabled():
    """weight-gradient kernels run on a second stream, concurrent with the dgrad / node-backward
    chain that does not depend on them (config.cfg.wgrad_stream)"""
    from . import profiling
    if profiling._active is not None:     # per-kernel event timing wants one stream
        return False
    return cfg.wgrad_stream


def _side_stream(device, which=0):
    key = (device.index if device.index is not None else torch.cuda.current_device(), which)
    s = _side_streams.get(key)
    if s is None:
        s = _side_streams[key] = torch.cuda.Stream(device=device)
    return s


def n_lanes():
    """streams the independent edges of a cell node are spread over (config.cfg.lanes, default 4;
    1 = everything on the caller's stream)"""
    from . import profiling
    if profiling._active is not None:
        return 1
    return cfg.lanes


class _Laned:
    """tape entry recorded on lane k > 0: its backward runs on the same lane"""
    __slots__ = ("fn", "lane")

    def __init__(self, fn, lane):
        self.fn = fn
        self.lane = lane


class _JoinMarker:
    """tape entry: all lanes rejoin the caller's stream here"""
    __slots__ = ()


class _Lane:
    """context manager: run the enclosed launches (ours and torch's) on lane k.  The lane first
    waits for everything enqueued on the caller's stream so far; ExecCtx.join_lanes() makes the
    caller's stream wait for the lanes.  Tensors allocated inside belong to the lane's allocator
    pool; they are only freed when the tape dies, after every stream has been joined."""

    def __init__(self, ctx, k, forward_only=False):
        self.ctx, self.k = ctx, k
        self.forward_only = forward_only     # tape entries keep the ENCLOSING lane (see on_sublane)
        self.tctx = None

    def __enter__(self):
        ctx, k = self.ctx, self.k
        if k == 0 or ctx.lanes <= 1:
            return self
        main = torch.cuda.current_stream(ctx.device)
        lane = _side_stream(ctx.device, k)
        ev = torch.cuda.Event()
        ev.record(main)
        lane.wait_event(ev)
        self.saved = (ctx.stream, ctx.lane, ctx.tape_lane)
        self.tctx = torch.cuda.stream(lane)
        self.tctx.__enter__()
        ctx.stream = lane.cuda_stream
        ctx.lane = k
        if not self.forward_only:
            ctx.tape_lane = k
        ctx.dirty_lanes.add(k)
        return self

    def __exit__(self, *exc):
        if self.tctx is not None:
            self.ctx.stream, self.ctx.lane, self.ctx.tape_lane = self.saved
            self.tctx.__exit__(*exc)
        return False


def get_lib():
    """the C-ABI library (or bench.py's event-timing proxy around it)"""
    from . import profiling
    return profiling.active() or _lib.load()


def _require_cuda(t, what):
    if not t.is_cuda:
        raise Nas3dDeviceError(
            "%s is on %s: the nas3d_b200 path runs on CUDA (sm_100a) only, there is no CPU fallback"
            % (what, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (what, t.dtype))


# ----------------------------------------------------------------------------------------
# activations
# ----------------------------------------------------------------------------------------
class Act:
    __slots__ = ("t", "g", "N", "C", "D", "H", "W", "ld", "requires_grad", "S", "bias_param",
                 "bias_done")

    def __init__(self, t, ld, requires_grad=True):
        self.t = t
        self.g = None
        self.S = None      # per-(n,c) fp64 {sum, sum^2} if a producer kernel already computed them
        self.bias_param = None   # bias of the conv that produced this tensor (GN-bwd yields its grad)
        self.bias_done = False
        self.N, self.C, self.D, self.H, self.W = t.shape
        self.ld = ld
        self.requires_grad = requires_grad

    @property
    def V(self):
        return self.D * self.H * self.W

    @property
    def ptr(self):
        return self.t.data_ptr()

    def grad_slot(self):
        """(grad Act-like tensor, accumulate flag) for a producer of d(this)."""
        if self.g is None:
            self.g = alloc(self.N, self.C, self.D, self.H, self.W, self.t.device)
            return self.g, 0
        return self.g, 1

    def slice(self, c0, c1):
        a = Act(self.t[:, c0:c1], self.ld, self.requires_grad)
        return a


def virtual_cat_enabled():
    return cfg.virtual_cat


class CatAct:
    """virtual concat along channels (cell.py:82): equal-width dense parts that are never copied
    into one buffer.  Only 1x1x1 convolutions consume a cell output, and they read the parts."""
    __slots__ = ("parts", "N", "C", "D", "H", "W", "requires_grad", "S")

    def __init__(self, parts):
        self.parts = parts
        p0 = parts[0]
        self.N, self.D, self.H, self.W = p0.N, p0.D, p0.H, p0.W
        self.C = sum(p.C for p in parts)
        self.requires_grad = any(p.requires_grad for p in parts)
        self.S = None

    @property
    def V(self):
        return self.D * self.H * self.W


def materialize_cat(ctx, cat):
    """copy the parts into one dense (N,C,D,H,W) buffer (module boundary only)"""
    out = new_act(cat.N, cat.C, cat.D, cat.H, cat.W, ctx.device)
    c0 = 0
    slices = []
    for p in cat.parts:
        sl = out.slice(c0, c0 + p.C)
        affine_sum(ctx, [Term(p)], sl)
        slices.append(sl)
        c0 += p.C
    w = cat.parts[0].C

    def bwd():
        if out.g is None:
            return
        for j, sl in enumerate(slices):
            sl.g = out.g[:, j * w:(j + 1) * w]
    ctx.push(bwd)
    return out


def alloc(N, Cc, D, H, W, device):
    """logical (N,C,D,H,W) tensor, physical NDHWC dense"""
    return torch.empty((N, D, H, W, Cc), device=device, dtype=torch.float32).permute(0, 4, 1, 2, 3)


def new_act(N, Cc, D, H, W, device):
    return Act(alloc(N, Cc, D, H, W, device), Cc)


def new_act_padded(N, Cc, D, H, W, device):
    """Cc channels stored at the next multiple-of-4 voxel pitch (the 3-channel head output,
    nas.py:50-52): every access to it is then one aligned 128-bit vector.  The pad lanes hold
    unspecified values and are never read as data."""
    ld = (Cc + 3) // 4 * 4
    if ld == Cc:
        return new_act(N, Cc, D, H, W, device)
    buf = torch.empty((N, D, H, W, ld), device=device, dtype=torch.float32)
    return Act(buf.permute(0, 4, 1, 2, 3)[:, :Cc], ld)


def _ndhwc_pitch(t):
    """voxel pitch if t is laid out NDHWC (possibly channel-sliced), else None"""
    N, Cc, D, H, W = t.shape
    sn, sc, sd, sh, sw = t.stride()
    ld = sw
    if ld < Cc:
        return None
    ok = (sc == 1 or Cc == 1) and (sh == W * ld or H == 1) and (sd == H * W * ld or D == 1) and (
        sn == D * H * W * ld or N == 1)
    return ld if ok else None


def as_act(t, requires_grad):
    """wrap a caller tensor; NCDHW input is re-laid out by our own kernel"""
    _require_cuda(t, "input tensor")
    if t.dim() != 5:
        raise ValueError("expected a 5-D (N,C,D,H,W) tensor, got shape %s" % (tuple(t.shape),))
    ld = _ndhwc_pitch(t)
    if ld is not None and ld % 4 == 0 and t.data_ptr() % 16 == 0:
        return Act(t, ld, requires_grad)
    if not t.is_contiguous():
        t = t.contiguous()
    N, Cc, D, H, W = t.shape
    ldd = (Cc + 3) // 4 * 4
    buf = torch.empty((N, D, H, W, ldd), device=t.device, dtype=torch.float32)
    if ldd != Cc:
        buf.zero_()
    check(get_lib().nas3d_ncdhw_to_ndhwc(t.data_ptr(), buf.data_ptr(), N, Cc, D * H * W, ldd,
                                           _stream()), "ncdhw_to_ndhwc")
    return Act(buf.permute(0, 4, 1, 2, 3)[:, :Cc], ldd, requires_grad)


def owned_grad(g, like, force_copy=False):
    """normalise an incoming gradient tensor to like's NDHWC layout"""
    ld = _ndhwc_pitch(g)
    # an output stored at a padded pitch (the 3-channel head, new_act_padded) takes its gradient at
    # the same pitch: the fused 1x1 backward reads both as aligned vectors
    padded = like.ld != like.C
    if (ld is not None and not force_copy and g.data_ptr() % 16 == 0 and (ld % 4 == 0 or like.C % 4)
            and (not padded or ld == like.ld)):
        return g, ld
    if padded:
        buf = torch.empty((like.N, like.D, like.H, like.W, like.ld), device=g.device, dtype=torch.float32)
        out = buf.permute(0, 4, 1, 2, 3)[:, :like.C]
        out.copy_(g)
        return out, like.ld
    out = alloc(like.N, like.C, like.D, like.H, like.W, g.device)
    out.copy_(g)
    return out, like.C


# ----------------------------------------------------------------------------------------
# execution context (one per outermost module call)
# ----------------------------------------------------------------------------------------
class Alpha:
    """softmax(alpha) matrix or row handed in by the caller (cell.py:24-33)"""
    __slots__ = ("t", "g", "ncol")

    def __init__(self, t):
        _require_cuda(t, "alpha")
        self.t = t.contiguous()
        self.g = None
        self.ncol = t.shape[-1]

    def ptr(self, row, k):
        off = (row * self.ncol + k) if self.t.dim() == 2 else k
        return self.t.data_ptr() + 4 * off

    def gptr(self, row, k):
        if self.g is None:
            self.g = torch.zeros_like(self.t)
        off = (row * self.ncol + k) if self.t.dim() == 2 else k
        return self.g.data_ptr() + 4 * off


class ExecCtx:
    def __init__(self, record, training, device):
        self.record = record
        self.training = training
        self.device = device
        self.tape = []
        self.params = {}        # id(param) -> param (those touched in forward)
        self.bucket = None
        self.views = {}
        self.lib = get_lib()
        self.stream = _stream()
        self.packed = {}
        self.pending_gn = []
        self.lanes = n_lanes()
        self.lane = 0                # stream lane the forward is running on (tags GroupNorm jobs)
        self.tape_lane = 0           # lane the backward of the entries being recorded will run on
        self.dirty_lanes = set()

    def use(self, *params):
        for p in params:
            if p is not None:
                self.params[id(p)] = p

    def push(self, fn):
        if self.record:
            self.tape.append(fn if self.tape_lane == 0 else _Laned(fn, self.tape_lane))

    def on_lane(self, k):
        return _Lane(self, k % self.lanes)

    def on_sublane(self, idx):
        """FORWARD-only concurrency for the candidate ops of one MixedOp (cell.py:30): they read
        the same input, so their dgrads accumulate into ONE gradient buffer and must stay serial -
        the tape entries recorded inside keep the enclosing lane - but their forward kernels are
        independent and run on a stream of their own (joined with the other lanes at the node)."""
        if self.lanes <= 1 or not cfg.candidate_lanes:
            return _Lane(self, 0)
        return _Lane(self, 8 + (self.lane % 8) * 8 + (idx % 8), forward_only=True)

    def join_lanes(self, mark=False):
        """the caller's stream waits for every lane used since the last join; with mark=True the
        backward pass joins at the mirror position too"""
        if self.dirty_lanes:
            main = torch.cuda.current_stream(self.device)
            for k in sorted(self.dirty_lanes):
                ev = torch.cuda.Event()
                ev.record(_side_stream(self.device, k))
                main.wait_event(ev)
            self.dirty_lanes.clear()
        if mark and self.record and self.lanes > 1:
            self.tape.append(_JoinMarker())

    # ---- backward side -------------------------------------------------------------
    def begin_backward(self):
        self.stream = _stream()
        nside = cfg.wgrad_streams if wgrad_stream_enabled() else 0
        # weight-gradient streams (round-robin): keys 100.. keep them apart from the branch lanes
        self.side = [_side_stream(self.device, 100 + i) for i in range(max(0, min(4, nside)))]
        self.side_next = 0
        self.side_used = set()
        self.keep = []
        plist = list(self.params.values())
        total = sum((p.numel() + 3) // 4 * 4 for p in plist)
        self.bucket = torch.zeros(max(total, 4), device=self.device, dtype=torch.float32)
        off = 0
        for p in plist:
            n = p.numel()
            self.views[id(p)] = self.bucket[off:off + n].view(p.shape)
            off += (n + 3) // 4 * 4

    def gptr(self, p):
        return self.views[id(p)].data_ptr() if p is not None else None

    def fork_wgrad(self, *keep):
        """stream for a weight-gradient launch: the side stream, ordered after everything enqueued
        on the main stream so far (its inputs dy / x are complete there).  Nothing on the main
        stream waits for it until join_wgrad(); `keep` tensors are held until then."""
        if not self.side:
            return self.stream
        main = torch.cuda.current_stream(self.device)
        i = self.side_next
        self.side_next = (i + 1) % len(self.side)
        ev = torch.cuda.Event()
        ev.record(main)
        self.side[i].wait_event(ev)
        self.side_used.add(i)
        self.keep.extend(keep)
        return self.side[i].cuda_stream

    def join_wgrad(self):
        main = torch.cuda.current_stream(self.device)
        for i in sorted(self.side_used):
            ev = torch.cuda.Event()
            ev.record(self.side[i])
            main.wait_event(ev)
        self.side_used = set()
        self.keep = []


# ----------------------------------------------------------------------------------------
# Terms and the fused affine-sum
# ----------------------------------------------------------------------------------------
MAX_TERMS = 32      # NAS3D_MAX_TERMS (include/nas3d_b200.h:224)


class Term:
    __slots__ = ("x", "a", "b", "relu", "kind", "alpha", "aux")

    def __init__(self, x, a=None, b=None, relu=False, kind="plain", aux=None):
        self.x = x
        self.a = a
        self.b = b
        self.relu = relu
        self.kind = kind      # 'plain' | 'gn' | 'se' | 'coef' (fixed per-(n,c) scale, no grad to it)
        self.alpha = None     # (Alpha, row, k)
        self.aux = aux

    @property
    def is_identity(self):
        return self.a is None and self.b is None and not self.relu and self.alpha is None


def _tp(t):
    return t.data_ptr() if t is not None else None


def gn_fold_enabled():
    """cfg.gn_fold: the GroupNorm coefficient kernels are folded into the affine kernels'
    prologues (94 fewer launches per searched-net step).  Off by default: measured on B200 it is
    SLOWER (406 vs 414 patches/s, profiles/r1f_ab_gn_fold_*.json) - the separate 5 us coefficient
    kernels overlap with other stream lanes, while a prologue of dependent fp64 loads / rsqrt /
    two block barriers delays the first load of every CTA of the one-wave streaming kernels."""
    return cfg.gn_fold


def _take_gn_jobs(ctx, terms, out):
    """pending GroupNorm coefficient jobs of `terms` that the affine kernel can compute itself
    (same G and eps): {id(term): job}; they leave ctx.pending_gn"""
    if not (gn_fold_enabled() and ctx.pending_gn):
        return {}
    pend = {id(j) for j in ctx.pending_gn}
    cand = [(t, t.aux["job"]) for t in terms
            if t.kind == "gn" and t.aux.get("job") is not None and id(t.aux["job"]) in pend]
    if not cand or len({(j[4], j[6]) for _, j in cand}) != 1:
        return {}
    if any((j[2], j[3], j[5]) != (out.N, out.C, out.V) for _, j in cand):
        return {}
    taken = {id(j) for _, j in cand}
    ctx.pending_gn = [j for j in ctx.pending_gn if id(j) not in taken]
    return {id(t): j for t, j in cand}


def affine_sum(ctx, terms, out):
    """out = sum_k w_k act_k(a_k x_k + b_k)   (cell.py:30,32,81 / prim_ops.py:75-80,152)

    The kernels take at most MAX_TERMS terms per launch; a longer list (a supernet node with
    n_nodes >= 8: 9 states x 4 candidates) is evaluated as a chain - the sum of the first
    MAX_TERMS terms becomes an identity term of the next launch - so the tape differentiates it
    like any other node."""
    while len(terms) > MAX_TERMS:
        part = new_act(out.N, out.C, out.D, out.H, out.W, ctx.device)
        affine_sum(ctx, terms[:MAX_TERMS], part)
        terms = [Term(part)] + list(terms[MAX_TERMS:])
    n = len(terms)
    lib = ctx.lib
    fold = _take_gn_jobs(ctx, terms, out)
    flush_gn(ctx)
    xs, lds = ptr_array([t.x.ptr for t in terms]), int_array([t.x.ld for t in terms])
    aa, bb = ptr_array([_tp(t.a) for t in terms]), ptr_array([_tp(t.b) for t in terms])
    ww = ptr_array([t.alpha[0].ptr(t.alpha[1], t.alpha[2]) if t.alpha else None for t in terms])
    rl = int_array([1 if t.relu else 0 for t in terms])
    if fold:
        jobs = [fold.get(id(t)) for t in terms]
        j0 = next(j for j in jobs if j is not None)
        rc = lib.nas3d_affine_sum_fwd_gn(
            n, xs, lds, aa, bb, ww, rl,
            ptr_array([j[0].data_ptr() if j else None for j in jobs]),
            ptr_array([j[1].weight.data_ptr() if j else None for j in jobs]),
            ptr_array([j[1].bias.data_ptr() if j else None for j in jobs]),
            ptr_array([j[8].data_ptr() if j else None for j in jobs]),
            j0[4], j0[6], out.ptr, out.ld, out.N, out.V, out.C, ctx.stream)
        check(rc, "affine_sum_fwd_gn")
    else:
        rc = lib.nas3d_affine_sum_fwd(n, xs, lds, aa, bb, ww, rl, out.ptr, out.ld, out.N, out.V,
                                      out.C, ctx.stream)
        check(rc, "affine_sum_fwd")
    ctx.push(lambda: _affine_sum_bwd(ctx, terms, out))
    return out


def _affine_sum_bwd(ctx, terms, out):
    if out.g is None:
        return
    lib = ctx.lib
    dout = out.g
    ld_dout = _ndhwc_pitch(dout)
    N, Cc, V = out.N, out.C, out.V
    dev = dout.device
    st = ctx.stream
    # pass 1: reductions for the terms that need them
    need = [t for t in terms if t.kind in ("gn", "se") or (t.alpha is not None)]
    R = {}
    if need:
        Rbuf = torch.empty((len(need), N, Cc, 2), device=dev, dtype=torch.float64)
        for i, t in enumerate(need):
            R[id(t)] = Rbuf[i]
        rc = lib.nas3d_affine_sum_bwd_reduce(
            len(need), ptr_array([t.x.ptr for t in need]), int_array([t.x.ld for t in need]),
            ptr_array([_tp(t.a) for t in need]), ptr_array([_tp(t.b) for t in need]),
            int_array([1 if t.relu else 0 for t in need]),
            dout.data_ptr(), ld_dout, ptr_array([R[id(t)].data_ptr() for t in need]),
            N, V, Cc, st)
        check(rc, "affine_sum_bwd_reduce")
    # per-term coefficient kernels (all GroupNorm terms of the node in one batched launch)
    P, Q, Rr = {}, {}, {}
    gn_terms = [t for t in terms if t.kind == "gn"]
    if gn_terms:
        pq_all = torch.empty((len(gn_terms), 3, N, Cc), device=dev, dtype=torch.float32)
        wps, dwps, bps = [], [], []
        for i, t in enumerate(gn_terms):
            P[id(t)], Q[id(t)], Rr[id(t)] = pq_all[i, 0], pq_all[i, 1], pq_all[i, 2]
            wps.append(t.alpha[0].ptr(t.alpha[1], t.alpha[2]) if t.alpha else None)
            dwps.append(t.alpha[0].gptr(t.alpha[1], t.alpha[2]) if t.alpha else None)
            # the conv that produced x (if any) gets its bias gradient from the same sums
            bp = t.x.bias_param if (t.x.bias_param is not None and not t.x.bias_done) else None
            if bp is not None:
                t.x.bias_done = True
            bps.append(ctx.gptr(bp) if bp is not None else None)
        by_g = {}
        for i, t in enumerate(gn_terms):
            by_g.setdefault(t.aux["G"], []).append(i)
        # one GroupNorm geometry and every GroupNorm term receives a dx: the apply kernel below
        # derives p, q, r (and the parameter gradients) itself
        fold_gn = (gn_fold_enabled() and len(by_g) == 1 and all(t.x.requires_grad for t in gn_terms))
        for G, idxs in ({} if fold_gn else by_g).items():
            for c0 in range(0, len(idxs), 32):
                ii = idxs[c0:c0 + 32]
                tt = [gn_terms[i] for i in ii]
                check(lib.nas3d_gn_bwd_coef_batch(
                    len(tt), ptr_array([R[id(t)].data_ptr() for t in tt]),
                    ptr_array([t.aux["mean_rstd"].data_ptr() for t in tt]),
                    ptr_array([t.aux["gamma"].data_ptr() for t in tt]),
                    ptr_array([t.a.data_ptr() for t in tt]), ptr_array([t.b.data_ptr() for t in tt]),
                    ptr_array([wps[i] for i in ii]), N, Cc, G, V,
                    ptr_array([pq_all[i, 0].data_ptr() for i in ii]),
                    ptr_array([pq_all[i, 1].data_ptr() for i in ii]),
                    ptr_array([pq_all[i, 2].data_ptr() for i in ii]),
                    ptr_array([ctx.gptr(t.aux["gamma"]) for t in tt]),
                    ptr_array([ctx.gptr(t.aux["beta"]) for t in tt]),
                    ptr_array([dwps[i] for i in ii]),
                    ptr_array([t.aux["S"].data_ptr() for t in tt]),
                    ptr_array([bps[i] for i in ii]), st), "gn_bwd_coef_batch")
    for t in terms:
        wptr = t.alpha[0].ptr(t.alpha[1], t.alpha[2]) if t.alpha else None
        dwptr = t.alpha[0].gptr(t.alpha[1], t.alpha[2]) if t.alpha else None
        if t.kind == "gn":
            continue
        elif t.kind == "se":
            aux = t.aux
            pq = torch.empty((2, N, Cc), device=dev, dtype=torch.float32)
            P[id(t)], Rr[id(t)] = pq[0], pq[1]
            fc0, fc2 = aux["fc0"], aux["fc2"]
            rc = lib.nas3d_se_bwd_coef(
                R[id(t)].data_ptr(), aux["S"].data_ptr(), t.a.data_ptr(), aux["hz"].data_ptr(),
                fc0.weight.data_ptr(), fc2.weight.data_ptr(), wptr, N, Cc, V,
                pq[0].data_ptr(), pq[1].data_ptr(),
                ctx.gptr(fc0.weight), ctx.gptr(fc0.bias), ctx.gptr(fc2.weight), ctx.gptr(fc2.bias),
                dwptr, st)
            check(rc, "se_bwd_coef")
        elif t.alpha is not None:
            # plain / fixed-coefficient term under a softmax weight: d alpha = <dout, y>
            if t.a is None and t.b is None and not t.relu:
                check(lib.nas3d_plain_bwd_coef(R[id(t)].data_ptr(), wptr, N, Cc, dwptr, st),
                      "plain_bwd_coef")
            else:
                raise NotImplementedError("alpha-weighted term of kind %r with coefficients" % t.kind)
    # pass 2: dx_k
    live = [t for t in terms if t.x.requires_grad]
    if not live:
        return
    seen = set()
    dxs, accs = [], []
    for t in live:
        g, acc = t.x.grad_slot()
        key = id(t.x)
        if key in seen:
            acc = 1
        seen.add(key)
        dxs.append(g)
        accs.append(acc)

    def pcoef(t):
        if id(t) in P:
            return P[id(t)].data_ptr()
        if t.kind == "coef" and t.a is not None:
            return t.a.data_ptr()     # y = a*x with constant a:  dx = a*dout
        return None

    common = (
        len(live), ptr_array([t.x.ptr for t in live]), int_array([t.x.ld for t in live]),
        ptr_array([_tp(t.a) for t in live]), ptr_array([_tp(t.b) for t in live]),
        int_array([1 if t.relu else 0 for t in live]),
        ptr_array([pcoef(t) for t in live]),
        ptr_array([Q[id(t)].data_ptr() if id(t) in Q else None for t in live]),
        ptr_array([Rr[id(t)].data_ptr() if id(t) in Rr else None for t in live]),
        ptr_array([t.alpha[0].ptr(t.alpha[1], t.alpha[2]) if t.alpha else None for t in live]),
        ptr_array([g.data_ptr() for g in dxs]), int_array([_ndhwc_pitch(g) for g in dxs]),
        int_array(accs), dout.data_ptr(), ld_dout)
    if gn_terms and fold_gn:
        gi = {id(t): i for i, t in enumerate(gn_terms)}

        def gp(fn):
            return ptr_array([fn(t, gi[id(t)]) if id(t) in gi else None for t in live])
        rc = lib.nas3d_affine_sum_bwd_apply_gn(
            *common,
            gp(lambda t, i: R[id(t)].data_ptr()), gp(lambda t, i: t.aux["mean_rstd"].data_ptr()),
            gp(lambda t, i: t.aux["gamma"].data_ptr()), gp(lambda t, i: ctx.gptr(t.aux["gamma"])),
            gp(lambda t, i: ctx.gptr(t.aux["beta"])), gp(lambda t, i: dwps[i]),
            gp(lambda t, i: t.aux["S"].data_ptr()), gp(lambda t, i: bps[i]),
            gn_terms[0].aux["G"], N, V, Cc, st)
        check(rc, "affine_sum_bwd_apply_gn")
    else:
        check(lib.nas3d_affine_sum_bwd_apply(*common, N, V, Cc, st), "affine_sum_bwd_apply")


def materialize(ctx, term):
    """evaluate a Term into an Act (no-op for identity terms)"""
    if term.is_identity:
        return term.x
    x = term.x
    out = new_act(x.N, x.C, x.D, x.H, x.W, ctx.device)
    return affine_sum(ctx, [term], out)


def bind_concat(ctx, out, nodes, c_node):
    """zero-copy concat (cell.py:82): nodes were written into channel slices of `out`; in the
    backward the node gradients alias the matching slices of d(out)."""
    def bwd():
        if out.g is None:
            return
        for j, node in enumerate(nodes):
            node.g = out.g[:, j * c_node:(j + 1) * c_node]
    ctx.push(bwd)


# ----------------------------------------------------------------------------------------
# GroupNorm / SE coefficient producers
# ----------------------------------------------------------------------------------------
def moments(ctx, x):
    S = torch.empty((x.N, x.C, 2), device=ctx.device, dtype=torch.float64)
    check(ctx.lib.nas3d_moments_nc(x.ptr, x.N, x.V, x.C, x.ld, S.data_ptr(), ctx.stream),
          "moments_nc")
    return S


def gn_term(ctx, x, norm, relu):
    """GroupNorm (+ReLU) of x as a lazy Term (prim_ops.py:56-58,75-80)"""
    if norm.num_channels != x.C:
        raise ValueError("GroupNorm expects %d channels, got %d" % (norm.num_channels, x.C))
    ctx.use(norm.weight, norm.bias)
    S = x.S if x.S is not None else moments(ctx, x)
    x.S = S        # identity / se_conv candidates of other MixedOps reuse the statistics of x
    G = norm.num_groups
    ab = torch.empty((2, x.N, x.C), device=ctx.device, dtype=torch.float32)
    mr = torch.empty((x.N, G, 2), device=ctx.device, dtype=torch.float32)
    # the coefficient kernel is deferred: affine_sum() flushes all pending ones of a node at once
    job = (S, norm, x.N, x.C, G, x.V, float(norm.eps), ab, mr, ctx.lane)
    ctx.pending_gn.append(job)
    aux = {"mean_rstd": mr, "gamma": norm.weight, "beta": norm.bias, "G": G, "S": S, "job": job}
    return Term(x, ab[0], ab[1], relu, "gn", aux)


def flush_gn(ctx):
    """launch the deferred GroupNorm coefficient kernels, batched per (N, C, G, V, eps)"""
    if not ctx.pending_gn:
        return
    # while lanes are running concurrently only this lane's statistics are known to be complete on
    # this stream; after a join everything pending is
    if ctx.dirty_lanes:
        mine = [j for j in ctx.pending_gn if j[9] == ctx.lane]
        ctx.pending_gn = [j for j in ctx.pending_gn if j[9] != ctx.lane]
    else:
        mine, ctx.pending_gn = ctx.pending_gn, []
    groups = {}
    for job in mine:
        groups.setdefault(job[2:7], []).append(job)
    for (N, Cc, G, V, eps), jobs in groups.items():
        for i in range(0, len(jobs), 32):
            chunk = jobs[i:i + 32]
            check(ctx.lib.nas3d_gn_coef_batch(
                len(chunk), ptr_array([j[0].data_ptr() for j in chunk]),
                ptr_array([j[1].weight.data_ptr() for j in chunk]),
                ptr_array([j[1].bias.data_ptr() for j in chunk]), N, Cc, G, V, eps,
                ptr_array([j[7][0].data_ptr() for j in chunk]),
                ptr_array([j[7][1].data_ptr() for j in chunk]),
                ptr_array([j[8].data_ptr() for j in chunk]), ctx.stream), "gn_coef_batch")


def se_term(ctx, x, fc):
    """x * sigmoid(fc(mean(x))) as a lazy Term (prim_ops.py:133-139,149-152)"""
    fc0, fc2 = fc[0], fc[2]
    ctx.use(fc0.weight, fc0.bias, fc2.weight, fc2.bias)
    S = x.S if x.S is not None else moments(ctx, x)
    x.S = S
    s = torch.empty((x.N, x.C), device=ctx.device, dtype=torch.float32)
    hz = torch.empty((x.N, 2), device=ctx.device, dtype=torch.float32)
    check(ctx.lib.nas3d_se_excite(S.data_ptr(), fc0.weight.data_ptr(), fc0.bias.data_ptr(),
                                  fc2.weight.data_ptr(), fc2.bias.data_ptr(), x.N, x.C, x.V,
                                  s.data_ptr(), hz.data_ptr(), ctx.stream), "se_excite")
    return Term(x, s, None, False, "se", {"S": S, "hz": hz, "fc0": fc0, "fc2": fc2})


# ----------------------------------------------------------------------------------------
# convolution
# ----------------------------------------------------------------------------------------
class ConvSpec:
    """static description of an nn.Conv3d / nn.ConvTranspose3d (prim_ops.py:93-110)"""
    __slots__ = ("k", "stride", "dil", "pad", "out_pad", "transposed", "depthwise", "cin", "cout")

    def __init__(self, m):
        self.transposed = isinstance(m, torch.nn.ConvTranspose3d)
        ks, st, dl, pd = m.kernel_size, m.stride, m.dilation, m.padding
        if len(set(ks)) != 1 or len(set(st)) != 1 or len(set(dl)) != 1 or len(set(pd)) != 1:
            raise NotImplementedError("anisotropic convolution parameters are not used by this network")
        self.k, self.stride, self.dil, self.pad = ks[0], st[0], dl[0], pd[0]
        self.out_pad = m.output_padding[0] if self.transposed else 0
        self.cin, self.cout = m.in_channels, m.out_channels
        if m.groups == 1:
            self.depthwise = False
        elif m.groups == m.in_channels == m.out_channels:
            self.depthwise = True
        else:
            raise NotImplementedError("groups must be 1 or C")
        if self.k not in (1, 3):
            raise NotImplementedError("kernel size %d (1 or 3 supported)" % self.k)

    def out_extent(self, n):
        if self.transposed:
            return (n - 1) * self.stride - 2 * self.pad + self.dil * (self.k - 1) + self.out_pad + 1
        return (n + 2 * self.pad - self.dil * (self.k - 1) - 1) // self.stride + 1


def _desc(spec, big, small):
    d = ConvDesc()
    d.N = big.N
    d.Db, d.Hb, d.Wb, d.Cb, d.ld_big = big.D, big.H, big.W, big.C, big.ld
    d.Ds, d.Hs, d.Ws, d.Cs, d.ld_small = small.D, small.H, small.W, small.C, small.ld
    d.k, d.stride, d.dil, d.pad = spec.k, spec.stride, spec.dil, spec.pad
    d.depthwise = 1 if spec.depthwise else 0
    return d


class _GradView:
    """lets the conv backward address a gradient tensor like an Act"""
    __slots__ = ("t", "N", "C", "D", "H", "W", "ld")

    def __init__(self, t, like):
        self.t = t
        self.N, self.C, self.D, self.H, self.W = like.N, like.C, like.D, like.H, like.W
        self.ld = _ndhwc_pitch(t)

    @property
    def ptr(self):
        return self.t.data_ptr()


def _umma_enabled(d):
    """tcgen05 path policy (measured, tools/umma_micro.py): it wins 2.5-5x over the CUDA-core
    kernels from 32 channels up; at 16 channels the tiled FFMA kernel is still ahead for
    stride 1, while the stride-2 dilation-1 tiled kernels stop at 8 channels."""
    if not cfg.umma:
        return False
    if cfg.umma_min_c is not None:
        return d.Cb >= cfg.umma_min_c
    return d.Cb >= (16 if (d.stride == 2 and d.dil == 1) else 32)


def _umma_packed(ctx, d, m, produce_big):
    """packed [W_hi|W_lo] operand of the tcgen05 path, cached per (weight, direction) for the
    lifetime of one forward/backward (weights only change in optimizer.step)"""
    n = ctx.lib.nas3d_umma_packed_floats(C.byref(d), produce_big)
    if n <= 0 or not _umma_enabled(d):
        return None
    key = (id(m.weight), produce_big, ctx.lib.nas3d_umma_pack_mode(C.byref(d), produce_big))
    wp = ctx.packed.get(key)
    if wp is None:
        wp = torch.empty(int(n), device=ctx.device, dtype=torch.float32)
        check(ctx.lib.nas3d_umma_pack_weights(C.byref(d), m.weight.data_ptr(), produce_big,
                                              wp.data_ptr(), ctx.stream), "umma_pack_weights")
        ctx.packed[key] = wp
    return wp


def _umma_prepack_list(module):
    """(conv module, channels) of every conv under `module` that the tcgen05 path serves; cached on
    the module (the structure of these networks is static)"""
    lst = module.__dict__.get("_nas3d_umma_list")
    if lst is None:
        lst = []
        for m in module.modules():
            if not isinstance(m, (torch.nn.Conv3d, torch.nn.ConvTranspose3d)):
                continue
            if m.kernel_size != (3, 3, 3) or m.groups != 1 or m.in_channels != m.out_channels:
                continue
            d = ConvDesc()
            d.Cb = d.Cs = m.in_channels
            d.k, d.stride, d.dil, d.pad, d.depthwise = 3, m.stride[0], m.dilation[0], m.padding[0], 0
            if d.Cb in (16, 32, 64) and len(set(m.stride)) == 1 and _umma_enabled(d):
                lst.append(m)
        module.__dict__["_nas3d_umma_list"] = lst
    return lst


def _prepack_umma(ctx, module):
    """pack the weight operands of all tcgen05 convs of this module call in ONE launch (forward
    direction, plus the dgrad direction when recording); conv() / _conv_bwd() find them in
    ctx.packed.  The stride-2 transposed direction is packed in parity-class order, which the
    kernel uses for even extents (always the case in these networks; otherwise _umma_packed
    falls back to packing that operand on demand)."""
    convs = _umma_prepack_list(module)
    if not convs:
        return
    lib = ctx.lib
    entries = []
    for m in convs:
        transposed = isinstance(m, torch.nn.ConvTranspose3d)
        dirs = [1 if transposed else 0]
        if ctx.record:
            dirs.append(0 if transposed else 1)
        for pb in dirs:
            mode = 0 if not pb else (2 if (m.stride[0] == 2 and m.dilation[0] == 1 and m.padding[0] == 1) else 1)
            entries.append((m, pb, mode))
    d = ConvDesc()
    d.k, d.depthwise = 3, 0
    sizes = []
    for m, pb, mode in entries:
        d.Cb = d.Cs = m.in_channels
        sizes.append(int(lib.nas3d_umma_packed_floats(C.byref(d), pb)))
    arena = torch.empty(sum(sizes), device=ctx.device, dtype=torch.float32)
    off = 0
    ptrs = []
    for (m, pb, mode), n in zip(entries, sizes):
        t = arena[off:off + n]
        off += n
        ctx.packed[(id(m.weight), pb, mode)] = t
        ptrs.append(t.data_ptr())
    check(lib.nas3d_umma_pack_weights_batch(
        len(entries), int_array([m.in_channels for m, _, _ in entries]),
        int_array([mode for _, _, mode in entries]),
        ptr_array([m.weight.data_ptr() for m, _, _ in entries]), ptr_array(ptrs), ctx.stream),
        "umma_pack_weights_batch")


def conv(ctx, x, m, spec, in_relu=False, in_scale=None, sigmoid=False, stats=False):
    """y = conv(f(x)) with f = optional relu / per-(n,c) scale prologue; returns the raw Act.

    Conv3d forward is the small-from-big gather, ConvTranspose3d forward the big-from-small
    gather of the same conv view (include/nas3d_b200.h)."""
    if x.C != spec.cin:
        raise ValueError("conv expects %d input channels, got %d" % (spec.cin, x.C))
    if isinstance(x, CatAct):
        widths = {p.C for p in x.parts}
        if (spec.k == 1 and not spec.transposed and not spec.depthwise and len(widths) == 1
                and len(x.parts) <= 4 and x.parts[0].C % 4 == 0):
            return _conv_cat(ctx, x, m, spec, in_relu, in_scale, sigmoid, stats)
        x = materialize_cat(ctx, x)
    lib = ctx.lib
    ctx.use(m.weight, m.bias)
    y = new_act(x.N, spec.cout, spec.out_extent(x.D), spec.out_extent(x.H), spec.out_extent(x.W),
                ctx.device)
    bias = m.bias.data_ptr() if m.bias is not None else None
    plain = not (in_relu or in_scale is not None or sigmoid)
    S = None
    if stats:    # GroupNorm follows: let the conv epilogue produce its statistics
        S = torch.empty((y.N, y.C, 2), device=ctx.device, dtype=torch.float64)
        y.S = S
        y.bias_param = m.bias        # GroupNorm is y's only consumer: its backward emits d(bias)
    Sp = _tp(S)
    if not spec.transposed:
        d = _desc(spec, x, y)
        wp = _umma_packed(ctx, d, m, 0) if plain else None
        if wp is not None:
            check(lib.nas3d_umma_conv(C.byref(d), 0, x.ptr, wp.data_ptr(), bias, y.ptr, 0, Sp,
                                      ctx.stream), "umma_conv fwd")
        else:
            check(lib.nas3d_conv_small_from_big(C.byref(d), x.ptr, m.weight.data_ptr(), bias,
                                                _tp(in_scale), 1 if in_relu else 0,
                                                1 if sigmoid else 0, y.ptr, 0, Sp, ctx.stream),
                  "conv_small_from_big")
    else:
        if not plain:
            raise NotImplementedError("prologue/epilogue on a transposed convolution")
        d = _desc(spec, y, x)
        wp = _umma_packed(ctx, d, m, 1)
        if wp is not None:
            check(lib.nas3d_umma_conv(C.byref(d), 1, x.ptr, wp.data_ptr(), bias, y.ptr, 0, Sp,
                                      ctx.stream), "umma_conv convT fwd")
        else:
            check(lib.nas3d_conv_big_from_small(C.byref(d), x.ptr, m.weight.data_ptr(), bias, None,
                                                0, None, y.ptr, 0, Sp, ctx.stream),
                  "conv_big_from_small")
    ctx.push(lambda: _conv_bwd(ctx, x, y, m, spec, in_relu, in_scale, sigmoid))
    return y


def _cat_desc(spec, x, small):
    d = ConvDesc()
    d.N = x.N
    d.Db, d.Hb, d.Wb, d.Cb, d.ld_big = x.D, x.H, x.W, x.C, x.C
    d.Ds, d.Hs, d.Ws, d.Cs, d.ld_small = small.D, small.H, small.W, small.C, small.ld
    d.k, d.stride, d.dil, d.pad = spec.k, spec.stride, spec.dil, spec.pad
    d.depthwise = 0
    return d


def _conv_cat(ctx, x, m, spec, in_relu, in_scale, sigmoid, stats):
    """1x1x1 conv whose input is a virtual concat (preprocess convs of the next cell, the head)"""
    lib = ctx.lib
    ctx.use(m.weight, m.bias)
    mk = new_act_padded if (spec.stride == 1 and fused_pw_bwd_enabled()) else new_act
    y = mk(x.N, spec.cout, spec.out_extent(x.D), spec.out_extent(x.H), spec.out_extent(x.W), ctx.device)
    S = None
    if stats:
        S = torch.empty((y.N, y.C, 2), device=ctx.device, dtype=torch.float64)
        y.S = S
        y.bias_param = m.bias
    d = _cat_desc(spec, x, y)
    parts = x.parts
    check(lib.nas3d_conv1x1_cat_fwd(
        C.byref(d), len(parts), ptr_array([p.ptr for p in parts]), int_array([p.ld for p in parts]),
        m.weight.data_ptr(), m.bias.data_ptr() if m.bias is not None else None, _tp(in_scale),
        1 if in_relu else 0, 1 if sigmoid else 0, y.ptr, _tp(S), ctx.stream), "conv1x1_cat_fwd")

    def bwd():
        if y.g is None:
            return
        st = ctx.stream
        dy_t = y.g
        if spec.stride == 1 and _pw_bwd_fused(ctx, d, parts, y, m, in_relu, in_scale, sigmoid):
            return
        if sigmoid:
            if y.ld != y.C or _ndhwc_pitch(dy_t) != y.C:
                raise Nas3dDeviceError("sigmoid backward of a pitched head output needs the fused 1x1 backward")
            dl = torch.empty_like(dy_t)
            check(lib.nas3d_sigmoid_bwd(y.ptr, dy_t.data_ptr(), dl.data_ptr(), y.N * y.V * y.C, st),
                  "sigmoid_bwd")
            dy_t = dl
        dy = _GradView(dy_t, y)
        d2 = _cat_desc(spec, x, dy)
        check(lib.nas3d_conv1x1_cat_wgrad(
            C.byref(d2), len(parts), ptr_array([p.ptr for p in parts]),
            int_array([p.ld for p in parts]), dy.ptr, _tp(in_scale), 1 if in_relu else 0,
            ctx.gptr(m.weight),
            ctx.gptr(m.bias) if (m.bias is not None and not y.bias_done) else None,
            ctx.fork_wgrad(dy_t)), "conv1x1_cat_wgrad")
        live = [p for p in parts if p.requires_grad]
        if not live:
            return
        if len(live) != len(parts):
            raise NotImplementedError("partial gradient through a virtual concat")
        gs, accs = [], []
        for p in parts:
            g, acc = p.grad_slot()
            gs.append(g)
            accs.append(acc)
        check(lib.nas3d_conv1x1_cat_dgrad(
            C.byref(d2), len(parts), ptr_array([g.data_ptr() for g in gs]),
            int_array([_ndhwc_pitch(g) for g in gs]), int_array(accs), dy.ptr, m.weight.data_ptr(),
            ptr_array([p.ptr if in_relu else None for p in parts]), int_array([p.ld for p in parts]),
            _tp(in_scale), st), "conv1x1_cat_dgrad")
    ctx.push(bwd)
    return y


def fused_pw_bwd_enabled():
    """1x1x1 stride-1 convs: dgrad + wgrad (+ sigmoid backward) in one pass (cfg.pw_fused_bwd =
    False: the separate kernels)"""
    return cfg.pw_fused_bwd


def _pw_bwd_fused(ctx, d, parts, y, m, in_relu, in_scale, sigmoid):
    """backward of a dense 1x1x1 stride-1 conv y = conv(f(cat(parts))) through
    nas3d_conv1x1_bwd_fused; returns False when the shape is not covered (caller falls back)"""
    lib = ctx.lib
    if not fused_pw_bwd_enabled() or not lib.nas3d_conv1x1_bwd_fused_supported(C.byref(d), len(parts)):
        return False
    dy_t = y.g
    ld_dy = _ndhwc_pitch(dy_t)
    if ld_dy != d.ld_small or dy_t.data_ptr() % 16:
        return False
    live = [p for p in parts if p.requires_grad]
    if live and len(live) != len(parts):
        return False
    if any(p.ld % 4 or p.ptr % 16 for p in parts):
        return False
    gs = lds = accs = None
    if live:
        gs, accs = [], []
        for p in parts:
            g, acc = p.grad_slot()
            gs.append(g)
            accs.append(acc)
        lds = [_ndhwc_pitch(g) for g in gs]
        if any(l is None or l % 4 or g.data_ptr() % 16 for l, g in zip(lds, gs)):
            raise Nas3dDeviceError("gradient buffer of a 1x1 conv input is not a dense NDHWC tensor")
    db = ctx.gptr(m.bias) if (m.bias is not None and not y.bias_done) else None
    check(lib.nas3d_conv1x1_bwd_fused(
        C.byref(d), len(parts), ptr_array([p.ptr for p in parts]), int_array([p.ld for p in parts]),
        ptr_array([g.data_ptr() for g in gs]) if gs else None, int_array(lds) if gs else None,
        int_array(accs) if gs else None, dy_t.data_ptr(), y.ptr if sigmoid else None,
        m.weight.data_ptr(), _tp(in_scale), 1 if in_relu else 0, ctx.gptr(m.weight), db,
        ctx.stream), "conv1x1_bwd_fused")
    return True


def _wgrad_workspace(ctx, d, prologue):
    """scratch for the split-K partial sums of the tcgen05 weight gradient (wide dense 3x3x3 convs);
    (None, 0) for every other shape"""
    n = int(ctx.lib.nas3d_conv_wgrad_workspace_floats(C.byref(d), 1 if prologue else 0))
    if n <= 0:
        return None, 0
    return torch.empty(n, device=ctx.device, dtype=torch.float32), n


def _conv_bwd(ctx, x, y, m, spec, in_relu, in_scale, sigmoid):
    if y.g is None:
        return
    lib = ctx.lib
    st = ctx.stream
    dy_t = y.g
    if (spec.k == 1 and spec.stride == 1 and not spec.transposed and not spec.depthwise
            and x.requires_grad and _pw_bwd_fused(ctx, _desc(spec, x, y), [x], y, m, in_relu,
                                                  in_scale, sigmoid)):
        return
    if sigmoid:
        # y holds probabilities; dlogit = dprob * p * (1-p)   (nn.Sigmoid, nas.py:52)
        dl = torch.empty_like(dy_t)
        assert _ndhwc_pitch(dy_t) == y.ld == y.C
        check(lib.nas3d_sigmoid_bwd(y.ptr, dy_t.data_ptr(), dl.data_ptr(), y.N * y.V * y.C, st),
              "sigmoid_bwd")
        dy_t = dl
    dy = _GradView(dy_t, y)
    dW = ctx.gptr(m.weight)
    db = ctx.gptr(m.bias) if (m.bias is not None and not y.bias_done) else None
    if not spec.transposed:
        d = _desc(spec, x, dy)
        ws, nws = _wgrad_workspace(ctx, d, in_relu or in_scale is not None)
        wst = ctx.fork_wgrad(dy_t, ws)
        check(lib.nas3d_conv_wgrad_ws(C.byref(d), dy.ptr, x.ptr, _tp(in_scale), 1 if in_relu else 0,
                                      dW, db, None, _tp(ws), nws, wst), "conv_wgrad")
        if x.requires_grad:
            g, acc = x.grad_slot()
            gv = _GradView(g, x)
            d2 = _desc(spec, gv, dy)
            wp = _umma_packed(ctx, d2, m, 1) if not (in_relu or in_scale is not None) else None
            if wp is not None:
                check(lib.nas3d_umma_conv(C.byref(d2), 1, dy.ptr, wp.data_ptr(), None, gv.ptr, acc,
                                          None, st), "umma_conv dgrad")
            else:
                check(lib.nas3d_conv_big_from_small(C.byref(d2), dy.ptr, m.weight.data_ptr(), None,
                                                    x.ptr if in_relu else None, x.ld,
                                                    _tp(in_scale), gv.ptr, acc, None, st),
                      "conv dgrad")
    else:
        d = _desc(spec, dy, x)
        ws, nws = _wgrad_workspace(ctx, d, False)
        wst = ctx.fork_wgrad(dy_t, ws)
        check(lib.nas3d_conv_wgrad_ws(C.byref(d), x.ptr, dy.ptr, None, 0, dW, None, db, _tp(ws), nws, wst),
              "convT wgrad")
        if x.requires_grad:
            g, acc = x.grad_slot()
            gv = _GradView(g, x)
            d2 = _desc(spec, dy, gv)
            wp = _umma_packed(ctx, d2, m, 0)
            if wp is not None:
                check(lib.nas3d_umma_conv(C.byref(d2), 0, dy.ptr, wp.data_ptr(), None, gv.ptr, acc,
                                          None, st), "umma_conv convT dgrad")
            else:
                check(lib.nas3d_conv_small_from_big(C.byref(d2), dy.ptr, m.weight.data_ptr(), None,
                                                    None, 0, 0, gv.ptr, acc, None, st),
                      "convT dgrad")


# ----------------------------------------------------------------------------------------
# pooling
# ----------------------------------------------------------------------------------------
def pool2(ctx, x, kind):
    if x.D % 2 or x.H % 2 or x.W % 2:
        raise ValueError("2x2x2 pooling needs even extents, got %s" % ((x.D, x.H, x.W),))
    y = new_act(x.N, x.C, x.D // 2, x.H // 2, x.W // 2, ctx.device)
    check(ctx.lib.nas3d_pool2_fwd(kind, x.ptr, x.ld, y.ptr, y.ld, x.N, y.D, y.H, y.W, x.C,
                                  ctx.stream), "pool2_fwd")

    def bwd():
        if y.g is None or not x.requires_grad:
            return
        g, acc = x.grad_slot()
        check(ctx.lib.nas3d_pool2_bwd(kind, x.ptr, x.ld, y.g.data_ptr(), _ndhwc_pitch(y.g),
                                      g.data_ptr(), _ndhwc_pitch(g), acc, x.N, y.D, y.H, y.W, x.C,
                                      ctx.stream), "pool2_bwd")
    ctx.push(bwd)
    return y


# ----------------------------------------------------------------------------------------
# the autograd bridge
# ----------------------------------------------------------------------------------------
_dp_state = {"enabled": False, "group": None}


class _ModuleFn(torch.autograd.Function):
    """One autograd node for a whole module call (prim op, MixedOp, Cell or a full net)."""

    @staticmethod
    def forward(fctx, module, n_acts, n_extras, record, *tensors):
        acts_t = tensors[:n_acts]
        extras_t = tensors[n_acts:n_acts + n_extras]
        dev = acts_t[0].device
        ctx = ExecCtx(record, module.training, dev)
        with torch.cuda.device(dev):
            acts = [as_act(t, bool(record and t.requires_grad)) for t in acts_t]
            extras = [Alpha(t) for t in extras_t]
            _prepack_umma(ctx, module)
            out = module._run(ctx, *acts, *extras)
            if isinstance(out, Term):
                out = materialize(ctx, out)
            if isinstance(out, CatAct):
                out = materialize_cat(ctx, out)
        fctx.ectx = ctx
        fctx.acts = acts
        fctx.extras = extras
        fctx.out = out
        fctx.module = module
        fctx.n_acts, fctx.n_extras = n_acts, n_extras
        fctx.param_ids = [id(p) for p in tensors[n_acts + n_extras:]]
        res = out.t
        if any(res.data_ptr() == a.t.data_ptr() for a in acts):
            res = res.clone()   # never hand an input back as the output
            fctx.out = Act(res, _ndhwc_pitch(res))
            out_src = out

            def bwd_alias():
                if fctx.out.g is not None and out_src.requires_grad:
                    g, acc = out_src.grad_slot()
                    if acc:
                        g.add_(fctx.out.g)
                    else:
                        g.copy_(fctx.out.g)
            ctx.push(bwd_alias)
        return res

    @staticmethod
    def backward(fctx, gout):
        ctx = fctx.ectx
        if ctx is None or not ctx.record:
            raise RuntimeError("backward through a nas3d module call that was not recorded")
        out = fctx.out
        with torch.cuda.device(ctx.device):
            ctx.begin_backward()
            g, _ = owned_grad(gout, out, force_copy=getattr(fctx.module, "_mutates_out_grad", False))
            out.g = g
            for fn in reversed(ctx.tape):
                if isinstance(fn, _Laned):
                    with ctx.on_lane(fn.lane):
                        fn.fn()
                elif isinstance(fn, _JoinMarker):
                    ctx.join_lanes()
                else:
                    fn()
            ctx.join_lanes()
            ctx.join_wgrad()
            ctx.tape = []
            if _dp_state["enabled"]:
                _dp_allreduce(ctx, fctx.extras)
        grads = [None, None, None, None]
        for a in fctx.acts:
            grads.append(a.g if a.requires_grad else None)
        for e in fctx.extras:
            grads.append(e.g if e.g is not None else (torch.zeros_like(e.t)))
        for pid in fctx.param_ids:
            grads.append(ctx.views.get(pid))
        fctx.ectx = None
        fctx.acts = fctx.extras = fctx.out = None
        return tuple(grads)


def _dp_allreduce(ctx, extras):
    """patch-batch data parallelism: average the flat gradient bucket over ranks (NCCL)"""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return
    world = dist.get_world_size(_dp_state["group"])
    if world == 1:
        return
    dist.all_reduce(ctx.bucket, group=_dp_state["group"])
    ctx.bucket.mul_(1.0 / world)
    for e in extras:
        if e.g is not None:
            dist.all_reduce(e.g, group=_dp_state["group"])
            e.g.mul_(1.0 / world)


def enable_data_parallel(group=None, enabled=True):
    """patch-batch data parallelism (one process per GPU): every module backward ends with an
    NCCL all-reduce (mean) of its flat gradient bucket, so the unchanged reference drivers
    (search.py / train.py step loops) train replicas in lock-step."""
    _dp_state["enabled"] = bool(enabled)
    _dp_state["group"] = group


def run_module(module, acts, extras=()):
    """entry used by every nn.Module.forward of this package"""
    acts = tuple(acts)
    extras = tuple(extras)
    for t in acts:
        _require_cuda(t, "%s input" % type(module).__name__)
    params = [p for p in module.parameters()]
    record = torch.is_grad_enabled() and (
        any(t.requires_grad for t in acts) or any(t.requires_grad for t in extras)
        or any(p.requires_grad for p in params))
    return _ModuleFn.apply(module, len(acts), len(extras), record, *acts, *extras, *params)
