"""B200-native WaveFe encoder engine: geometry, HBM buffer plan and the
forward/backward orchestration over the C-ABI kernels.

Formulation.  Activations are channel-last (n, t, c).
Every strided 1-D convolution of the reference (FeBlock, modules.py:1058-1077;
SincConv_fast, modules.py:920-934) becomes ONE GEMM whose A operand is an
*overlapping-row view* of the reflect-padded activation: row (n, t) is the
k*Cin contiguous floats starting at time t*stride, so no im2col is ever
materialised and all strides (1, 2, 10) use the same kernel.  The 251-tap
Cin=1 sinc layer is folded 4x in time (polyphase) so that it has the same
shape.  Train-mode BatchNorm statistics are accumulated in the GEMM epilogue;
BN+PReLU+reflect-pad+dense-skip pooling are one elementwise pass that writes
the next layer's padded operand.  Dense skips are pooled BEFORE their 1x1
projection (mean-pool commutes with a bias-free 1x1 conv, frontend.py:182,
213-232), which removes the reference's (N,256,T) intermediates; the pooled
skips and the QRNN output are concatenated so W + all skips are one GEMM.
"""
import math
import torch

from . import ops

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SINC_FOLD = 4            # polyphase fold of the sinc layer on the fp32 FFMA path
TC_SINC_FOLD = 32        # on the tensor-core path one folded row = one 128-byte TMA row
# numerics of the GEMMs: None = fp32 FFMA; 1 = 3xTF32 tcgen05 (fp32-equivalent); 0 = TF32
PRECISIONS = {"fp32": None, "3xtf32": 1, "tf32": 0}


def _cdiv(a, b):
    return -(-a // b)


def _ru(a, m):
    return _cdiv(a, m) * m


def conv_pads(k, stride, sinc):
    """Reflect pad (left, right): SincConv_fast modules.py:922-928, FeBlock
    modules.py:1058-1071 (dilation 1)."""
    if sinc:
        return (k // 2 - 1, k // 2) if stride > 1 else (k // 2, k // 2)
    if k <= 1:
        return (0, 0)
    if stride > 1 or k % 2 == 0:
        return (k // 2 - 1, k // 2)
    return (k // 2, k // 2)


class ConvGeom(object):
    """Static geometry of one encoder block for a given input length."""

    def __init__(self, idx, sinc, Cin, Cout, k, s, L_in, sinc_fold=4):
        self.idx, self.sinc = idx, sinc
        self.Cin, self.Cout, self.k, self.s, self.L_in = Cin, Cout, k, s, L_in
        self.padL, self.padR = conv_pads(k, s, sinc)
        if self.padL >= L_in or self.padR >= L_in:
            raise ValueError("block %d: reflect pad (%d,%d) needs an input longer than "
                             "%d samples" % (idx, self.padL, self.padR, L_in))
        self.Tpad = L_in + self.padL + self.padR
        self.T_out = (self.Tpad - k) // s + 1
        if sinc:
            # polyphase fold: row u of the GEMM produces times fold*u .. fold*u+fold-1
            self.fold = sinc_fold
            self.lda = sinc_fold                       # floats between consecutive rows
            self.K = _ru(k + sinc_fold - 1, 32 if sinc_fold >= 32 else 4)
            self.Nn = sinc_fold * Cout
            self.P = _cdiv(self.Tpad, sinc_fold)       # rows per sample
            self.rows_out = _cdiv(self.T_out, sinc_fold)
        else:
            self.fold = 1
            self.lda = s * Cin
            self.K = k * Cin
            self.Nn = Cout
            self.P = _cdiv(self.Tpad, s)
            self.rows_out = self.T_out
        self.Ty = self.rows_out * self.fold            # time pitch of the raw output
        self.apad_floats = self.P * self.lda           # per sample
        # backward-data: taps of the polyphase transposed convolution
        self.taps = _cdiv(k, s)
        self.Pd = self.P + self.taps - 1               # rows per sample of zero-padded dy

    @property
    def count(self):
        return self.T_out


def build_geometry(cfg, T, sinc_fold=4):
    geoms, L, cin = [], T, cfg["num_inputs"]
    for i, (k, s, f) in enumerate(zip(cfg["kwidths"], cfg["strides"], cfg["fmaps"])):
        sinc = bool(cfg["sincnet"]) and i == 0
        if sinc and k % 2 == 0:
            k += 1                                     # modules.py:835-836
        g = ConvGeom(i, sinc, cin, f, k, s, L, sinc_fold)
        geoms.append(g)
        L, cin = g.T_out, f
    return geoms


def frame_counts(cfg, T):
    return [g.T_out for g in build_geometry(cfg, T)]


class EncoderPlan(object):
    """All HBM buffers of one (N, T) problem, allocated once and reused every
    step (180 GB HBM: nothing is recomputed or re-allocated).  Padded operand
    buffers are zero-initialised once; kernels only ever write their valid
    region, so halos / slack stay finite."""

    def __init__(self, cfg, N, T, device, prec="fp32"):
        self.cfg, self.N, self.T, self.device = cfg, N, T, device
        if prec not in PRECISIONS:
            raise ValueError("precision must be one of %s" % (sorted(PRECISIONS),))
        self.prec, self.mode = prec, PRECISIONS[prec]
        self._twins = {}
        self._wtables = {}
        self.geoms = build_geometry(cfg, T, TC_SINC_FOLD if self.mode is not None else SINC_FOLD)
        G = self.geoms
        self.nblk = len(G)
        self.Tq = G[-1].T_out
        self.emb = cfg["emb_dim"]
        self.rnn = bool(cfg["rnn_pool"])
        self.skips = bool(cfg["denseskips"])
        self.norm_out = bool(cfg["norm_out"])
        self.Clast = G[-1].Cout
        self.H = (cfg["rnn_dim"] // 2) * 2 if self.rnn else 0
        f32 = dict(dtype=torch.float32, device=device)
        f64 = dict(dtype=torch.float64, device=device)
        z = lambda n: torch.zeros(int(n), **f32)
        e = lambda n: torch.empty(int(n), **f32)

        # concatenated [rnn-out | pooled skips] operand of the output projection
        self.col_off = []
        off = self.H if self.rnn else self.Clast
        for g in G[:-1]:
            self.col_off.append(off)
            if self.skips:
                off += g.Cout
        self.Kc = off
        self.pool_d = [max(g.T_out // self.Tq, 1) for g in G[:-1]]

        self.apad, self.y, self.bn = [], [], []
        self.Wt, self.dWt, self.Wd, self.dyz, self.dxpad = [], [], [], [], []
        for g in G:
            self.apad.append(z(N * g.apad_floats + g.K + 64))
            self.y.append(e(N * g.rows_out * g.Nn))
            self.bn.append(torch.zeros(4, g.Cout, **f32))       # mean, invstd, scale, shift
            self.Wt.append(e(g.Nn * g.K))
            self.dWt.append(e(g.Nn * g.K))
            if g.sinc:
                self.Wd.append(None)
                self.dyz.append(z(N * g.rows_out * g.Nn))
                self.dxpad.append(None)
            else:
                self.Wd.append(e(g.s * g.Cin * g.taps * g.Cout))
                self.dyz.append(z(N * g.Pd * g.Cout + g.taps * g.Cout + 64))
                self.dxpad.append(e(N * g.apad_floats) if g.idx > 0 else None)
        rows = N * self.Tq
        self.rows = rows
        if self.rnn:
            Cq, H = self.Clast, self.H
            self.xq = z(N * (self.Tq + 1) * Cq + 2 * Cq + 64)
            self.Yg, self.Cst, self.dYg = e(rows * 3 * H), e(rows * H), e(rows * 3 * H)
            self.dsrc = e(rows * 2 * Cq)
            self.WqT = e(2 * Cq * 3 * H)
            self.dWq = e(3 * H * 2 * Cq)
        self.cat = z(rows * self.Kc)
        self.dcat = e(rows * self.Kc)
        self.WcatT = e(self.Kc * self.emb)
        self.dWcat = e(self.emb * self.Kc)
        self.yout = e(rows * self.emb)
        self.g = torch.zeros(rows * self.emb + 64, **f32)     # +slack: TC wgrad reads 32-col blocks
        self.bn_out = torch.zeros(4, self.emb, **f32)
        self.bn_out[1].fill_(1.0)
        self.bn_out[2].fill_(1.0)

        # double-precision accumulators: forward batch statistics ...
        self.fs_off, n = [], 0
        for g in G:
            self.fs_off.append(n)
            n += 2 * g.Nn
        self.fs_out = n
        n += 2 * self.emb
        self.stats_f = torch.zeros(n, **f64)
        # ... and backward reductions: per block S1,S2,dalpha,dbias; output S1,S2;
        # qrnn bias; W bias.
        self.bs_off, n = [], 0
        for g in G:
            self.bs_off.append(n)
            n += 4 * g.Cout
        self.bs_out = n
        n += 2 * self.emb
        self.bs_bq = n
        n += 3 * self.H
        self.bs_bw = n
        n += self.emb
        self.stats_b = torch.zeros(n, **f64)
        self.grad_vec = torch.zeros(n, **f32)
        self.zeros64 = torch.zeros(max(2 * max(g.Cout for g in G), 2 * self.emb), **f64)
        self.generation = 0

    # -- GEMM dispatch: fp32 FFMA kernels or tcgen05 tensor-core kernels -------------
    def split(self, name, buf, fresh=True, weights=False):
        """mode 1 (3xTF32): (hi, lo) parts of an operand.
        Activations: kind::tf32 reads only the upper 19 bits of an fp32 operand, so the
        operand itself is "hi" and only lo = rn(x - trunc(x)) needs a twin buffer.
        Weights (small): explicit hi = rn(x), lo = rn(x - hi), so that the dropped lo*lo and
        residual terms have random sign (no coherent bias in strongly cancelling sums).
        `fresh=False` reuses the split computed earlier in the same step."""
        if self.mode != 1:
            return buf, None
        tw = self._twins.get(name)
        if tw is None or tw[1].numel() != buf.numel():
            tw = (torch.zeros_like(buf) if weights else None, torch.zeros_like(buf))
            self._twins[name] = tw
            fresh = True
        if fresh:
            ops.call("pase_split_tf32", buf, tw[0], tw[1], buf.numel())
        return (tw[0] if weights else buf), tw[1]

    def twins_w(self, name, buf):
        """(hi, lo) twin buffers of a weight operand for producers that write the split
        themselves (pase_conv_w_batch); (None, None) outside 3xTF32 mode."""
        if self.mode != 1:
            return None, None
        tw = self._twins.get(name)
        if tw is None or tw[0] is None or tw[1].numel() != buf.numel():
            tw = (torch.zeros_like(buf), torch.zeros_like(buf))
            self._twins[name] = tw
        return tw

    def w_batch(self, op, jobs, dst_base=None):
        """One launch for the weight re-layouts of every conv block (pase_conv_w_batch).
        jobs: [(src, dst tensor | element offset, hi, lo, Cout, Cin, k, s, taps, count)].
        The device job table is rebuilt only when a pointer changed."""
        ptr = lambda t: 0 if t is None else (t if isinstance(t, int) else t.data_ptr())
        # shared-memory tiled kernels (ops 3..5) when every job fits their tile shapes
        tiled = all(Cout % 32 == 0 and Cin % 8 == 0 and (Cin + 1) * k <= 12000
                    and 32 * (8 * k + 1) <= 12000
                    for (_, _, _, _, Cout, Cin, k, _, _, _) in jobs)
        rows, start, blocks = [], 0, 0
        for (src, dst, hi, lo, Cout, Cin, k, sd, taps, count) in jobs:
            rows.append([ptr(src), ptr(dst), ptr(hi), ptr(lo), Cout, Cin, k, sd, taps, start,
                         count, blocks])
            start += count
            blocks += (Cout // 32) * (Cin // 8) if op == 1 else Cout
        ent = self._wtables.get(op)
        if ent is None or ent[0] != rows:
            table = torch.tensor(rows, dtype=torch.int64).reshape(-1).to(self.device)
            ent = (rows, table)
            self._wtables[op] = ent
        if tiled:
            ops.call("pase_conv_w_batch", ent[1], len(rows), blocks, op + 3, dst_base)
        else:
            ops.call("pase_conv_w_batch", ent[1], len(rows), start, op, dst_base)

    def lo_of(self, name, buf):
        """Residual twin of an activation operand, for producers that write it themselves
        (mode 1 only; zero-initialised so halos / slack stay valid)."""
        if self.mode != 1:
            return None
        tw = self._twins.get(name)
        if tw is None or tw[1].numel() != buf.numel():
            tw = (None, torch.zeros_like(buf))
            self._twins[name] = tw
        return tw[1]

    def nt(self, an, A, lda, afresh, bn, B, ldb, bfresh, C, ldc, M, N, K, alpha, bias,
           rows_in, t_valid, rows_out, fold, cs, cq, acc):
        if self.mode is None or lda % 32 != 0 or K % 4 != 0 or ldb % 4 != 0:
            return ops.call("pase_gemm_nt", A, lda, B, ldb, C, ldc, M, N, K, alpha, bias,
                            rows_in, t_valid, rows_out, fold, cs, cq, acc)
        Ah, Al = self.split(an, A, afresh)
        Bh, Bl = self.split(bn, B, bfresh, weights=True)
        return ops.call("pase_tc_gemm_nt", Ah, Al, A.numel() // lda, lda, Bh, Bl, ldb, C, ldc,
                        M, N, K, alpha, bias, rows_in, t_valid, rows_out, fold, cs, cq, acc,
                        self.mode)

    def tn(self, an, A, lda, pitchA, offA, afresh, bn, B, ldb, pitchB, bfresh, C, ldc, I, J,
           groups, rpg, alpha, acc):
        if self.mode is None or ldb % 32 != 0 or lda % 4 != 0 or I % 4 != 0 or J % 32 != 0:
            return ops.call("pase_gemm_tn", A, lda, pitchA, offA, B, ldb, pitchB, 0, C, ldc, I, J,
                            groups, rpg, alpha, acc)
        Ah, Al = self.split(an, A, afresh)
        Bh, Bl = self.split(bn, B, bfresh)
        return ops.call("pase_tc_gemm_tn", Ah, Al, lda, pitchA, offA, Bh, Bl, ldb, pitchB,
                        B.numel() // ldb, C, ldc, I, J, groups, rpg, alpha, acc, self.mode)

    def nbytes(self):
        tot = 0
        for v in self.__dict__.values():
            vs = v if isinstance(v, list) else [v]
            for t in vs:
                if isinstance(t, torch.Tensor):
                    tot += t.numel() * t.element_size()
        return tot


def sinc_constants(k, sr, device):
    """window_ and n_ exactly as SincConv_fast builds them (modules.py:868-876)."""
    n_lin = torch.linspace(0, (k / 2) - 1, steps=int(k / 2))
    window = 0.54 - 0.46 * torch.cos(2 * math.pi * n_lin / k)
    n = (k - 1) / 2.0
    n_ = 2 * math.pi * torch.arange(-n, 0).view(1, -1) / sr
    return n_.reshape(-1).contiguous().to(device), window.contiguous().to(device)


class ParamPack(object):
    """Flat, ordered view of the encoder parameters/buffers handed to the engine."""

    def __init__(self, module):
        self.module = module
        self.names, self.tensors = [], []
        for name, p in module.named_parameters():
            self.names.append(name)
            self.tensors.append(p)
        self.index = {n: i for i, n in enumerate(self.names)}


def _flat(t):
    return t.detach().reshape(-1)


def encoder_forward(plan, mod, x, params, training, save_for_backward):
    """x: (N,1,T) fp32 CUDA tensor.  params: dict name -> tensor.  Returns
    (out (N,emb,Tq), out_ntc (N*Tq, emb))."""
    call = ops.call
    cfg, G, N = plan.cfg, plan.geoms, plan.N
    Tq, Kc, rows, emb = plan.Tq, plan.Kc, plan.rows, plan.emb
    plan.generation += 1
    P = lambda name: _flat(params[name])
    buf = lambda name: _flat(mod.get_buffer(name))

    g0 = G[0]
    call("pase_reflect_pad_wave", x.reshape(-1), plan.apad[0], N, plan.T, g0.padL, g0.padR,
         g0.apad_floats)
    if training:
        plan.stats_f.zero_()
    if plan.skips:
        plan.cat.zero_()

    # GEMM operands (+ 3xTF32 split) of every conv block's weight: one launch
    jobs = []
    for l, g in enumerate(G):
        if not g.sinc:
            hi, lo = plan.twins_w(("Wt", l), plan.Wt[l])
            jobs.append((P("blocks.%d.conv.weight" % l), plan.Wt[l], hi, lo, g.Cout, g.Cin, g.k,
                         1, 1, g.Cout * g.Cin * g.k))
    if jobs:
        plan.w_batch(0, jobs)
    counters = []

    for l, g in enumerate(G):
        pre = "blocks.%d." % l
        if g.sinc:
            call("pase_sinc_make", P(pre + "conv.low_hz_"), P(pre + "conv.band_hz_"),
                 mod._sinc_n, mod._sinc_win, None, plan.Wt[l], g.Cout, g.k, g.fold, g.K,
                 50.0, 50.0, float(cfg["sr"]))
            bias = None
        else:
            bias = P(pre + "conv.bias")
        if training:
            o = plan.fs_off[l]
            cs, cq = plan.stats_f[o:o + g.Nn], plan.stats_f[o + g.Nn:o + 2 * g.Nn]
        else:
            cs = cq = None
        plan.nt(("apad", l), plan.apad[l], g.lda, l == 0, ("Wt", l), plan.Wt[l], g.K, g.sinc,
                plan.y[l], g.Nn, N * g.P, g.Nn, g.K, 1.0, bias, g.P, g.T_out, g.rows_out,
                g.fold, cs, cq, 0)
        mean, invstd, scale, shift = plan.bn[l][0], plan.bn[l][1], plan.bn[l][2], plan.bn[l][3]
        if training:
            call("pase_bn_finalize", cs, cq, g.Cout, g.fold, float(N * g.T_out),
                 P(pre + "norm.weight"), P(pre + "norm.bias"),
                 buf(pre + "norm.running_mean"), buf(pre + "norm.running_var"),
                 BN_MOMENTUM, BN_EPS, mean, invstd, scale, shift)
            counters.append(mod.get_buffer(pre + "norm.num_batches_tracked"))
        else:
            call("pase_bn_eval_affine", buf(pre + "norm.running_mean"),
                 buf(pre + "norm.running_var"), P(pre + "norm.weight"), P(pre + "norm.bias"),
                 g.Cout, BN_EPS, mean, invstd, scale, shift)
        C = g.Cout
        dst_lo = None            # 3xTF32: the producer also writes the operand's tf32 residual
        if l + 1 < plan.nblk:
            nx = G[l + 1]
            dst, d_ss, d_rs, pl, pr = plan.apad[l + 1], nx.apad_floats, C, nx.padL, nx.padR
            dst_lo = plan.lo_of(("apad", l + 1), plan.apad[l + 1])
        elif plan.rnn:
            dst, d_ss, d_rs, pl, pr = plan.xq[C:], (Tq + 1) * C, C, 0, 0
            lo = plan.lo_of("xq", plan.xq)
            dst_lo = None if lo is None else lo[C:]
        else:
            dst, d_ss, d_rs, pl, pr = plan.cat, Tq * Kc, Kc, 0, 0
        if plan.skips and l + 1 < plan.nblk:
            pool, pd = plan.cat[plan.col_off[l]:], plan.pool_d[l]
        else:
            pool, pd = None, 0
        call("pase_bn_prelu_pad_fwd", plan.y[l], g.Ty * C, N, g.T_out, C, scale, shift,
             P(pre + "act.weight"), dst, d_ss, d_rs, pl, pr, pool, Tq * Kc, Kc, pd, Tq, dst_lo)

    if plan.rnn:
        Cq, H = plan.Clast, plan.H
        Wl = params["rnn.layers.0.linear.weight"].detach()
        Wq = torch.cat([Wl[:, Cq:], Wl[:, :Cq]], 1).contiguous()      # [x_{t-1} | x_t] order
        plan.Wq = Wq
        plan.nt("xq", plan.xq, Cq, False, "Wq", Wq.reshape(-1), 2 * Cq, True, plan.Yg, 3 * H,
                N * (Tq + 1), 3 * H, 2 * Cq, 1.0, P("rnn.layers.0.linear.bias"),
                Tq + 1, Tq, Tq, 1, None, None, 0)
        call("pase_qrnn_scan_fwd", plan.Yg, plan.cat, Kc, plan.Cst, N, Tq, H)

    parts = [params["W.weight"].detach().reshape(emb, -1)]
    if plan.skips:
        parts += [params["denseskips.%d.weight" % i].detach().reshape(emb, -1)
                  for i in range(plan.nblk - 1)]
    Wcat = torch.cat(parts, 1).contiguous() if len(parts) > 1 else parts[0].contiguous()
    plan.Wcat = Wcat
    use_stats = plan.norm_out and training
    if use_stats:
        o = plan.fs_out
        cs, cq = plan.stats_f[o:o + emb], plan.stats_f[o + emb:o + 2 * emb]
    else:
        cs = cq = None
    plan.nt("cat", plan.cat, Kc, True, "Wcat", Wcat.reshape(-1), Kc, True, plan.yout, emb,
            rows, emb, Kc, 1.0, P("W.bias"), rows, rows, rows, 1, cs, cq, 0)
    bo = plan.bn_out
    if plan.norm_out:
        if training:
            call("pase_bn_finalize", cs, cq, emb, 1, float(rows), None, None,
                 buf("norm_out.running_mean"), buf("norm_out.running_var"),
                 BN_MOMENTUM, BN_EPS, bo[0], bo[1], bo[2], bo[3])
            counters.append(mod.get_buffer("norm_out.num_batches_tracked"))
        else:
            call("pase_bn_eval_affine", buf("norm_out.running_mean"),
                 buf("norm_out.running_var"), None, None, emb, BN_EPS,
                 bo[0], bo[1], bo[2], bo[3])
    if counters:                                 # BatchNorm step counters: one launch
        torch._foreach_add_(counters, 1)
    out = torch.empty(N, emb, Tq, dtype=torch.float32, device=x.device)
    out_ntc = torch.empty(rows, emb, dtype=torch.float32, device=x.device)
    call("pase_out_affine_nct", plan.yout, bo[2], bo[3], out.reshape(-1), out_ntc.reshape(-1),
         N, Tq, emb)
    plan.last_training = training
    return out, out_ntc


def encoder_backward(plan, mod, params, gout, gntc, training):
    """Returns dict name -> gradient tensor (fp32, parameter shape)."""
    call = ops.call
    cfg, G, N = plan.cfg, plan.geoms, plan.N
    Tq, Kc, rows, emb = plan.Tq, plan.Kc, plan.rows, plan.emb
    P = lambda name: _flat(params[name])
    grads = {}
    sb = plan.stats_b
    sb.zero_()
    # dgrad operands (+ split) of every conv block with an input gradient: one launch
    jobs = []
    for l, g in enumerate(G):
        if not g.sinc and l > 0:
            hi, lo = plan.twins_w(("Wd", l), plan.Wd[l])
            jobs.append((P("blocks.%d.conv.weight" % l), plan.Wd[l], hi, lo, g.Cout, g.Cin, g.k,
                         g.s, g.taps, g.s * g.Cin * g.taps * g.Cout))
    if jobs:
        plan.w_batch(1, jobs)
    zeros = plan.zeros64
    bo = plan.bn_out
    o = plan.bs_out
    S1o, S2o = sb[o:o + emb], sb[o + emb:o + 2 * emb]
    call("pase_out_bwd_reduce", None if gout is None else gout.reshape(-1),
         None if gntc is None else gntc.reshape(-1), plan.yout, bo[0], bo[1], N, Tq, emb,
         plan.g, S1o, S2o)
    use_stats = 1 if (plan.norm_out and training) else 0
    call("pase_out_bwd_apply", plan.g, plan.yout, bo[0], bo[1], bo[2], S1o, S2o, float(rows),
         use_stats, rows, emb)
    call("pase_colsum", plan.g, emb, rows, emb, sb[plan.bs_bw:plan.bs_bw + emb])
    plan.tn("g", plan.g, emb, rows, 0, True, "cat", plan.cat, Kc, rows, False, plan.dWcat, Kc,
            emb, Kc, 1, rows, 1.0, 0)
    call("pase_transpose_pad", plan.Wcat.reshape(-1), Kc, plan.WcatT, emb, emb, Kc)
    plan.nt("g", plan.g, emb, False, "WcatT", plan.WcatT, emb, True, plan.dcat, Kc, rows, Kc,
            emb, 1.0, None, rows, rows, rows, 1, None, None, 0)
    dWcat = plan.dWcat.view(emb, Kc)
    first = plan.H if plan.rnn else plan.Clast
    # gradients handed to autograd must not alias plan-owned buffers (the next backward
    # would overwrite them): copy the slices out
    grads["W.weight"] = dWcat[:, :first].clone().reshape(params["W.weight"].shape)
    if plan.skips:
        for i in range(plan.nblk - 1):
            c0 = plan.col_off[i]
            grads["denseskips.%d.weight" % i] = \
                dWcat[:, c0:c0 + G[i].Cout].clone().reshape(
                    params["denseskips.%d.weight" % i].shape)

    Cl = plan.Clast
    if plan.rnn:
        Cq, H = Cl, plan.H
        call("pase_qrnn_scan_bwd", plan.Yg, plan.Cst, plan.dcat, Kc, plan.dYg, N, Tq, H)
        call("pase_colsum", plan.dYg, 3 * H, rows, 3 * H, sb[plan.bs_bq:plan.bs_bq + 3 * H])
        plan.tn("dYg", plan.dYg, 3 * H, Tq, 0, True, "xq", plan.xq, Cq, Tq + 1, False,
                plan.dWq, 2 * Cq, 3 * H, 2 * Cq, N, Tq, 1.0, 0)
        dWq = plan.dWq.view(3 * H, 2 * Cq)
        grads["rnn.layers.0.linear.weight"] = torch.cat([dWq[:, Cq:], dWq[:, :Cq]], 1)
        call("pase_transpose_pad", plan.Wq.reshape(-1), 2 * Cq, plan.WqT, 3 * H, 3 * H, 2 * Cq)
        plan.nt("dYg", plan.dYg, 3 * H, False, "WqT", plan.WqT, 3 * H, True, plan.dsrc, 2 * Cq,
                rows, 2 * Cq, 3 * H, 1.0, None, rows, rows, rows, 1, None, None, 0)
        last_src = dict(A=plan.dsrc[Cq:], a_ss=Tq * 2 * Cq, a_rs=2 * Cq,
                        B=plan.dsrc, b_ss=Tq * 2 * Cq, b_rs=2 * Cq, b_shift=1)
    else:
        last_src = dict(A=plan.dcat, a_ss=Tq * Kc, a_rs=Kc, B=None, b_ss=0, b_rs=0, b_shift=0)

    for l in range(plan.nblk - 1, -1, -1):
        g = G[l]
        C = g.Cout
        pre = "blocks.%d." % l
        mean, invstd, scale, shift = plan.bn[l][0], plan.bn[l][1], plan.bn[l][2], plan.bn[l][3]
        if l == plan.nblk - 1:
            s = dict(last_src)
            s.update(padL=0, padR=0)
        else:
            nx = G[l + 1]
            s = dict(A=plan.dxpad[l + 1], a_ss=nx.apad_floats, a_rs=C, padL=nx.padL,
                     padR=nx.padR, B=None, b_ss=0, b_rs=0, b_shift=0)
        if plan.skips and l + 1 < plan.nblk:
            pool, pd = plan.dcat[plan.col_off[l]:], plan.pool_d[l]
        else:
            pool, pd = None, 0
        if g.sinc:
            dst, d_ss = plan.dyz[l], g.Ty * C
        else:
            dst, d_ss = plan.dyz[l][(g.taps - 1) * C:], g.Pd * C
        o = plan.bs_off[l]
        S1, S2, dal, dbi = sb[o:o + C], sb[o + C:o + 2 * C], sb[o + 2 * C:o + 3 * C], \
            sb[o + 3 * C:o + 4 * C]
        call("pase_bn_prelu_bwd_reduce", plan.y[l], g.Ty * C, N, g.T_out, C, mean, invstd,
             scale, shift, P(pre + "act.weight"),
             s["A"], s["a_ss"], s["a_rs"], s["padL"], s["padR"],
             s["B"], s["b_ss"], s["b_rs"], s["b_shift"],
             pool, Tq * Kc, Kc, pd, Tq, dst, d_ss, S1, S2, dal)
        if training:
            a1, a2 = S1, S2
        else:
            a1, a2 = zeros[:C], zeros[C:2 * C]
        lo = plan.lo_of(("dyz", l), plan.dyz[l])
        dst_lo = None if lo is None else (lo if g.sinc else lo[(g.taps - 1) * C:])
        call("pase_bn_prelu_bwd_apply", plan.y[l], g.Ty * C, N, g.T_out, C, mean, invstd,
             P(pre + "norm.weight"), a1, a2, float(N * g.T_out), dst, d_ss,
             None if g.sinc else dbi, dst_lo)
        if g.sinc:
            plan.tn(("dyz", l), plan.dyz[l], g.Nn, g.rows_out, 0, False, ("apad", l), plan.apad[l],
                    g.lda, g.P, False, plan.dWt[l], g.K, g.Nn, g.K, N, g.rows_out, 1.0, 0)
            dlow = torch.empty_like(params[pre + "conv.low_hz_"])
            dband = torch.empty_like(params[pre + "conv.band_hz_"])
            call("pase_sinc_grad", plan.dWt[l], P(pre + "conv.low_hz_"),
                 P(pre + "conv.band_hz_"), mod._sinc_n, mod._sinc_win, dlow.reshape(-1),
                 dband.reshape(-1), g.Cout, g.k, g.fold, g.K, 50.0, 50.0, float(cfg["sr"]))
            grads[pre + "conv.low_hz_"] = dlow
            grads[pre + "conv.band_hz_"] = dband
        else:
            plan.tn(("dyz", l), plan.dyz[l], C, g.Pd, g.taps - 1, False, ("apad", l), plan.apad[l],
                    g.lda, g.P, False, plan.dWt[l], g.K, C, g.K, N, g.T_out, 1.0, 0)
            if l > 0:
                plan.nt(("dyz", l), plan.dyz[l], C, False, ("Wd", l), plan.Wd[l], g.taps * C, False,
                        plan.dxpad[l], g.s * g.Cin, N * g.Pd, g.s * g.Cin, g.taps * C, 1.0, None,
                        g.Pd, g.P, g.P, 1, None, None, 0)

    # conv weight gradients: GEMM layout -> parameter layout for every block in one launch,
    # into one per-call buffer (the gradients handed to autograd are views of it)
    jobs, off = [], 0
    for l, g in enumerate(G):
        if not g.sinc:
            cnt = g.Cout * g.Cin * g.k
            jobs.append((plan.dWt[l], off, None, None, g.Cout, g.Cin, g.k, 1, 1, cnt))
            off += cnt
    if jobs:
        dWflat = torch.empty(off, dtype=torch.float32, device=plan.device)
        plan.w_batch(2, jobs, dWflat)
        for (_, o, _, _, Cout, Cin, k, _, _, cnt), l in zip(jobs, [l for l, g in enumerate(G)
                                                                  if not g.sinc]):
            grads["blocks.%d.conv.weight" % l] = dWflat[o:o + cnt].view(Cout, Cin, k)

    # one cast for every small reduction (double accumulators -> fp32 gradients) into a
    # per-call vector; the per-parameter gradients are views of it (no copies)
    gv = torch.empty_like(plan.grad_vec)
    call("pase_cast_d2f", sb, gv, sb.numel(), 1.0)
    for l, g in enumerate(G):
        C, o = g.Cout, plan.bs_off[l]
        pre = "blocks.%d." % l
        grads[pre + "norm.bias"] = gv[o:o + C]
        grads[pre + "norm.weight"] = gv[o + C:o + 2 * C]
        grads[pre + "act.weight"] = gv[o + 2 * C:o + 3 * C]
        if not g.sinc:
            grads[pre + "conv.bias"] = gv[o + 3 * C:o + 4 * C]
    grads["W.bias"] = gv[plan.bs_bw:plan.bs_bw + emb]
    if plan.rnn:
        grads["rnn.layers.0.linear.bias"] = gv[plan.bs_bq:plan.bs_bq + 3 * plan.H]
    return grads


class _EncoderFn(torch.autograd.Function):
    """Whole-encoder autograd node: one forward / one backward sweep over the
    plan's buffers (no per-op autograd graph)."""

    @staticmethod
    def forward(ctx, x, mod, plan, training, names, *tensors):
        params = dict(zip(names, tensors))
        out, out_ntc = encoder_forward(plan, mod, x, params, training, True)
        ctx.mod, ctx.plan, ctx.training, ctx.names = mod, plan, training, names
        ctx.generation = plan.generation
        ctx.save_for_backward(*tensors)
        return out, out_ntc

    @staticmethod
    def backward(ctx, gout, gntc):
        plan = ctx.plan
        if plan.generation != ctx.generation:
            raise RuntimeError(
                "pase_b200: the encoder's activation buffers were overwritten by a later "
                "forward of the same (N,T) shape before this backward ran; run backward "
                "before the next forward (or build a second WaveFe instance).")
        params = dict(zip(ctx.names, ctx.saved_tensors))
        if gout is not None:
            gout = gout.contiguous()
        if gntc is not None:
            gntc = gntc.contiguous()
        grads = encoder_backward(plan, ctx.mod, params, gout, gntc, ctx.training)
        return (None, None, None, None, None) + tuple(grads.get(n) for n in ctx.names)
