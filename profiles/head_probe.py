"""GPU probe of the fused head kernel (csrc/head.cu): structured-input checks that localise an operand-layout
error (which of M / N / K is mis-mapped), a random-input check against a TF32-emulated fp64 product, the candidate
parity against decode_compact on the kernel's own head tensor, and timings at the BASELINE spp-608 shapes.

    python profiles/head_probe.py [stage ...]        stages: xrate struct small cand big time   (default: all, each in a
                                                     subprocess with a timeout so a trap in one does not stop the rest)
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SPP_ANCHORS = [[(116, 90), (156, 198), (373, 326)], [(30, 61), (62, 45), (59, 119)], [(10, 13), (16, 30), (33, 23)]]


def emit(**kw):
    print(json.dumps(kw), flush=True)


def tf32_trunc(t):
    import torch
    return (t.contiguous().view(torch.int32) & ~0x1fff).view(torch.float32)


def tf32_round(t):
    import torch
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1fff).view(torch.float32)


def run_head(x, w, bias, slope, spec, nc, head_out=True, buf=None, conf=0.3, row_off=0, rows=None):
    import torch
    from pytorch_yolo_b200 import ops
    n_out = spec.na * (nc + 5)
    n_pad = (n_out + 15) // 16 * 16
    wp = torch.zeros(n_pad, w.shape[1], device=x.device)
    wp[:n_out] = w
    hw = ops.HeadWeights(wp.contiguous(), bias.float().cpu().contiguous(), slope, n_out)
    ho = torch.full((x.shape[0], n_out, spec.ny, spec.nx), float("nan"), device=x.device) if head_out else None
    ops.head_decode_compact([x], [hw], [spec], [row_off], rows if rows is not None else spec.rows, nc, conf, buf,
                            head_outs=[ho], candidates=buf is not None)
    torch.cuda.synchronize()
    return ho


def ref_head(x, w, bias, slope, mode):
    import torch
    f = {"trunc": tf32_trunc, "round": tf32_round, "fp32": lambda t: t}[mode]
    xd, wd = f(x).double(), f(w).double()
    y = torch.einsum("oc,bcp->bop", wd, xd.flatten(2)) + bias.double().to(x.device)[None, :, None]
    y = torch.maximum(y, y * slope)
    return y.view(x.shape[0], w.shape[0], x.shape[2], x.shape[3])


def stage_struct():
    import torch
    from pytorch_yolo_b200 import ops
    dev = "cuda:0"
    nc, na = 80, 3
    B, Cc, ny, nx = 2, 64, 16, 20          # plane 320 = 2.5 tiles
    spec = ops.scale_spec(SPP_ANCHORS[0], ny, nx, 32 * max(ny, nx))
    n = na * (nc + 5)
    zero_b = torch.zeros(n)
    pos = torch.arange(ny * nx, device=dev, dtype=torch.float32).view(1, 1, ny, nx)
    cases = {
        # out[o, p] = o * C : N mapping
        "N": (torch.ones(B, Cc, ny, nx, device=dev), torch.arange(n, device=dev, dtype=torch.float32)[:, None].expand(n, Cc).contiguous()),
        # out[o, p] = p * C' : M mapping  (W = 1/64 exactly representable)
        "M": ((pos % 128).expand(B, Cc, ny, nx).contiguous(), torch.full((n, Cc), 1.0 / 64, device=dev)),
        # out[o, p] = o % C : K mapping
        "K": (torch.arange(Cc, device=dev, dtype=torch.float32).view(1, Cc, 1, 1).expand(B, Cc, ny, nx).contiguous(),
              (torch.arange(n, device=dev)[:, None] % Cc == torch.arange(Cc, device=dev)[None, :]).float().contiguous()),
        # batch mapping: out = b + 1
        "B": (torch.arange(1, B + 1, device=dev, dtype=torch.float32).view(B, 1, 1, 1).expand(B, Cc, ny, nx).contiguous(),
              torch.full((n, Cc), 1.0 / 64, device=dev)),
    }
    for name, (x, w) in cases.items():
        got = run_head(x, w, zero_b, 1.0, spec, nc)
        want = ref_head(x, w, zero_b, 1.0, "fp32").float()
        bad = (got != want) | torch.isnan(got)
        info = dict(stage="struct", case=name, mismatches=int(bad.sum()), total=got.numel())
        if bad.any():
            idx = bad.nonzero()[:6].tolist()
            info["first"] = [(i, float(got[tuple(i)]), float(want[tuple(i)])) for i in idx]
            info["bad_by_o"] = bad.sum(dim=(0, 2, 3))[:12].tolist()
            info["bad_by_pos"] = bad.flatten(2).sum(dim=(0, 1))[:40].tolist()
        emit(**info)


def stage_small():
    import torch
    from pytorch_yolo_b200 import ops
    dev = "cuda:0"
    torch.manual_seed(1)
    for (B, Cc, ny, nx, nc) in [(2, 64, 16, 20, 80), (3, 256, 38, 38, 80), (2, 96, 12, 12, 20), (1, 32, 8, 8, 1)]:
        spec = ops.scale_spec(SPP_ANCHORS[1], ny, nx, 16 * max(ny, nx))
        n = 3 * (nc + 5)
        x = torch.randn(B, Cc, ny, nx, device=dev)
        w = torch.randn(n, Cc, device=dev) / Cc ** 0.5
        bias = torch.randn(n)
        got = run_head(x, w, bias, 0.1, spec, nc).double()
        res = {}
        for mode in ("trunc", "round", "fp32"):
            want = ref_head(x, w, bias, 0.1, mode)
            res[mode] = float((got - want).abs().max())
        emit(stage="small", shape=[B, Cc, ny, nx, nc], max_abs_err=res, nan=int(torch.isnan(got).sum()))


def stage_cand():
    """Candidates of the fused kernel == decode_compact on the head tensor the same launch wrote (bit-exact)."""
    import torch
    from pytorch_yolo_b200 import ops
    dev = "cuda:0"
    torch.manual_seed(2)
    for (B, Cc, ny, nx, nc, conf) in [(4, 64, 16, 20, 80, 0.3), (8, 256, 76, 76, 80, 0.1), (2, 128, 12, 12, 20, 0.01), (2, 32, 8, 8, 1, 0.2)]:
        spec = ops.scale_spec(SPP_ANCHORS[2], ny, nx, 8 * max(ny, nx))
        n = 3 * (nc + 5)
        x = torch.randn(B, Cc, ny, nx, device=dev)
        w = torch.randn(n, Cc, device=dev) * (2.0 / Cc ** 0.5)
        w[4::nc + 5] *= 1.5
        bias = torch.randn(n) * 0.5
        bias[4::nc + 5] -= 4.0
        bias[20::nc + 5] += 30.0          # one saturated class logit per anchor: exercises the sigmoid-space rescan
        bias[(21 if nc > 20 else 6)::nc + 5] += 31.0
        buf = ops.Buffers(dev, B, spec.rows, nc)
        ho = run_head(x, w, bias, 1.0, spec, nc, buf=buf, conf=conf)
        cnt_f = buf.meta[:B].clone()
        ovf = int(buf.meta[B])
        box_f, meta_f = buf.cand_box.clone().view(B, -1, 4), buf.cand_meta.clone().view(B, -1, 4)
        ops.decode_compact([ho], [spec], nc, conf, buf)
        torch.cuda.synchronize()
        cnt_d = buf.meta[:B].clone()
        box_d, meta_d = buf.cand_box.view(B, -1, 4), buf.cand_meta.view(B, -1, 4)
        same = bool((cnt_f == cnt_d).all())
        exact = same
        if same:
            for b in range(B):
                k = int(cnt_f[b])
                of = meta_f[b, :k, 3].argsort()
                od = meta_d[b, :k, 3].argsort()
                exact &= bool((meta_f[b, :k][of] == meta_d[b, :k][od]).all()) and bool((box_f[b, :k][of].view(torch.int32) == box_d[b, :k][od].view(torch.int32)).all())
        emit(stage="cand", shape=[B, Cc, ny, nx, nc], conf=conf, counts_fused=cnt_f.tolist()[:8], counts_decode=cnt_d.tolist()[:8],
             overflow=ovf, bit_exact=exact)


def _spp_inputs(B, dev, nc=80):
    """Feature maps ~ N(0,1) and head weights scaled per output channel so that the head tensor follows SYNTH-A
    (xy ~ N(0,1), wh ~ N(0,0.5^2), obj ~ N(-7,3^2), cls ~ N(-2,2^2)): ~2 % of the anchors pass conf 0.3."""
    import torch
    from pytorch_yolo_b200 import ops
    shapes = [(1024, 19), (512, 38), (256, 76)]
    specs = [ops.scale_spec(SPP_ANCHORS[k], g, g, 608) for k, (_, g) in enumerate(shapes)]
    no = nc + 5
    std = torch.tensor(([1.0, 1.0, 0.5, 0.5, 3.0] + [2.0] * nc) * 3)
    mean = torch.tensor(([0.0, 0.0, 0.0, 0.0, -7.0] + [-2.0] * nc) * 3)
    feats, ws, bs = [], [], []
    g = torch.Generator(device=dev).manual_seed(3)
    for (Cc, gsz) in shapes:
        feats.append(torch.randn(B, Cc, gsz, gsz, device=dev, generator=g))
        ws.append(torch.randn(3 * no, Cc, device=dev, generator=g) * (std.to(dev)[:, None] / Cc ** 0.5))
        bs.append(mean.clone())
    return specs, feats, ws, bs


def stage_big():
    """spp-608 batch 8: fused scales (38^2, 76^2) + cuDNN conv + decode_compact for 19^2 vs the unfused path."""
    import torch
    from pytorch_yolo_b200 import ops
    dev = "cuda:0"
    B, nc = 8, 80
    specs, feats, ws, bs = _spp_inputs(B, dev)
    for k in (1, 2):
        got = run_head(feats[k], ws[k], bs[k], 0.1, specs[k], nc).double()
        want = ref_head(feats[k], ws[k], bs[k], 0.1, "trunc")
        emit(stage="big", scale=k, max_abs_err_vs_trunc=float((got - want).abs().max()),
             max_abs_err_vs_fp32=float((got - ref_head(feats[k], ws[k], bs[k], 0.1, "fp32")).abs().max()))
    # fused candidates + NMS vs the unfused path on the kernel's own head tensors, all three scales (19^2 unfused in both)
    rows = sum(sp.rows for sp in specs)
    offs = [0, specs[0].rows, specs[0].rows + specs[1].rows]
    hws = []
    for k in range(3):
        wp = torch.zeros(256, ws[k].shape[1], device=dev)
        wp[:255] = ws[k]
        hws.append(ops.HeadWeights(wp, bs[k].float(), 1.0, 255))
    head0 = (torch.einsum("oc,bcp->bop", ws[0], feats[0].flatten(2)) + bs[0].to(dev)[None, :, None]).view(B, 255, 19, 19).contiguous()
    houts = [torch.empty(B, 255, sp.ny, sp.nx, device=dev) for sp in specs[1:]]
    buf = ops.Buffers(dev, B, rows, nc)
    ops.head_decode_compact(feats[1:], hws[1:], specs[1:], offs[1:], rows, nc, 0.3, buf, head_outs=houts)
    ops.decode_compact([head0], specs[:1], nc, 0.3, buf, row_offs=offs[:1], rows_per_img=rows, accumulate=True)
    out_f, row_f = buf.new_outputs()
    ops.nms(buf, 0.5, out_f, row_f)
    cand_f, kept_f, ovf = ops.read_counts(buf)
    cand_f, kept_f = cand_f.clone(), kept_f.clone()
    buf2 = ops.Buffers(dev, B, rows, nc)
    ops.decode_compact([head0] + houts, specs, nc, 0.3, buf2)
    out_d, row_d = buf2.new_outputs()
    ops.nms(buf2, 0.5, out_d, row_d)
    cand_d, kept_d, _ = ops.read_counts(buf2)
    same = bool((kept_f == kept_d).all()) and bool((cand_f == cand_d).all())
    if same:
        for b in range(B):
            n = int(kept_f[b])
            same &= bool((out_f[b, :n].view(torch.int32) == out_d[b, :n].view(torch.int32)).all()) and bool((row_f[b, :n] == row_d[b, :n]).all())
    emit(stage="big", what="fused(38,76)+ldg(19) -> nms vs decode_compact(3 scales) -> nms", cand=cand_f.tolist(), kept=kept_f.tolist(),
         overflow=ovf, bit_exact=same)


def stage_time():
    import torch
    from pytorch_yolo_b200 import ops
    dev = "cuda:0"
    B, nc = 64, 80
    specs, feats, ws, bs = _spp_inputs(B, dev)
    rows = sum(s.rows for s in specs)
    offs = [0, specs[0].rows, specs[0].rows + specs[1].rows]
    buf = ops.Buffers(dev, B, rows, nc)
    hws = []
    for k in range(3):
        wp = torch.zeros(256, ws[k].shape[1], device=dev)
        wp[:255] = ws[k]
        hws.append(ops.HeadWeights(wp, bs[k].float(), 1.0, 255))

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    for k in (1, 2):
        t = timeit(lambda: ops.head_decode_compact([feats[k]], [hws[k]], [specs[k]], [offs[k]], rows, nc, 0.3, buf))
        xb = feats[k].numel() * 4
        fl = 2.0 * B * specs[k].ny * specs[k].nx * 256 * feats[k].shape[1]
        emit(stage="time", what="fused head scale %d" % k, us=t, x_bytes=xb, x_gbs=xb / t / 1e3, tflops=fl / t / 1e6,
             cand=int(buf.meta[:B].sum()), overflow=int(buf.meta[B]))
        t1 = timeit(lambda: ops.head_decode_compact([feats[k]], [hws[k]], [specs[k]], [offs[k]], rows, nc, 0.3, buf, cta_pair=True))
        emit(stage="time", what="fused head scale %d, CTA-pair kernel" % k, us=t1, x_gbs=xb / t1 / 1e3, cand=int(buf.meta[:B].sum()))
        t = timeit(lambda: ops.head_decode_compact([feats[k]], [hws[k]], [specs[k]], [offs[k]], rows, nc, 0.3, buf, candidates=False))
        def prof(fl):
            return timeit(lambda: ops.head_decode_compact([feats[k]], [hws[k]], [specs[k]], [offs[k]], rows, nc, 0.3, buf, _profile_flags=fl,
                                                          cta_pair=False))
        t2, t3, t4, t5 = prof(0x100), prof(0x300), prof(0x500), prof(0x200)
        emit(stage="time", what="scale %d decomposition" % k, no_finish_us=t, mainloop_only_us=t2, mainloop_x_gbs=xb / t2 / 1e3,
             mainloop_no_w_us=t3, mainloop_no_x_us=t4, full_no_w_us=t5)
    t = timeit(lambda: ops.head_decode_compact(feats[1:], hws[1:], specs[1:], offs[1:], rows, nc, 0.3, buf))
    emit(stage="time", what="fused head scales 1+2 (two launches)", us=t)
    # 19x19 (C_in 1024): plane-padded copy, then the same kernel
    xp = torch.zeros(B, 1024, 364, device=dev)
    t_pad = timeit(lambda: ops.pad_feature(feats[0], out=xp))
    t0 = timeit(lambda: ops.head_decode_compact([xp], [hws[0]], [specs[0]], [offs[0]], rows, nc, 0.3, buf))
    def all3():
        ops.pad_feature(feats[0], out=xp)
        ops.head_decode_compact([xp] + feats[1:], hws, specs, offs, rows, nc, 0.3, buf)
    t3 = timeit(all3)
    emit(stage="time", what="19x19: pad copy + fused head", pad_us=t_pad, fused_us=t0, all_three_scales_us=t3,
         x_bytes_all=sum(f.numel() * 4 for f in feats), cand=int(buf.meta[:B].sum()))

    # the unfused path on the same inputs: cuDNN 1x1 conv (TF32) + bias + leaky, then decode_compact
    convs = []
    for k in range(3):
        c = torch.nn.Conv2d(ws[k].shape[1], 255, 1, bias=True).to(dev)
        with torch.no_grad():
            c.weight.copy_(ws[k].view(255, -1, 1, 1))
            c.bias.copy_(bs[k].to(dev))
        convs.append(c)
    with torch.no_grad():
        def heads_fn():
            return [convs[k](feats[k]) for k in range(3)]
        t_conv = timeit(heads_fn)
        heads = heads_fn()
        t_dec = timeit(lambda: ops.decode_compact(heads, specs, nc, 0.3, buf))
        t_dec12 = timeit(lambda: ops.decode_compact(heads[1:], specs[1:], nc, 0.3, buf, row_offs=offs[1:], rows_per_img=rows, accumulate=True))
        t_conv12 = timeit(lambda: [convs[k](feats[k]) for k in (1, 2)])
    emit(stage="time", what="unfused: torch (cuDNN) 1x1 conv + bias, then decode_compact", conv_3_scales_us=t_conv, conv_scales_1_2_us=t_conv12,
         decode_compact_3_scales_us=t_dec, decode_compact_scales_1_2_us=t_dec12, allow_tf32=torch.backends.cudnn.allow_tf32)


def stage_xrate():
    """The 76x76 / C_in 256 scale at batch 64: full kernel, main loop alone (no epilogue), X only / W only, for the single-CTA
    and the CTA-pair kernel.  YOLO_B200_LIB selects an alternative build of the library (kernel studies)."""
    import torch
    from pytorch_yolo_b200 import ops
    dev = "cuda:0"
    nc = 80
    for B in (64,):
        specs, feats, ws, bs = _spp_inputs(B, dev)
        rows = sum(s.rows for s in specs)
        offs = [0, specs[0].rows, specs[0].rows + specs[1].rows]
        buf = ops.Buffers(dev, B, rows, nc)
        k = 2
        wp = torch.zeros(256, ws[k].shape[1], device=dev)
        wp[:255] = ws[k]
        hw = ops.HeadWeights(wp, bs[k].float(), 1.0, 255)

        def t(fl, single):
            fn = lambda: ops.head_decode_compact([feats[k]], [hw], [specs[k]], [offs[k]], rows, nc, 0.3, buf, _profile_flags=fl, cta_pair=not single)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 20 * 1e3
        emit(stage="xrate", batch=B, env={k_: v_ for k_, v_ in os.environ.items() if k_.startswith("YB_HEAD")}, x_mb=feats[k].numel() * 4 / 1e6,
             single_full=t(0, True), single_mainloop=t(0x100, True), single_no_w=t(0x300, True), single_no_x=t(0x500, True), single_x_stream_no_mma=t(0xB00, True), single_xw_stream_no_mma=t(0x900, True), single_epilogue_only=t(0xE00, True), single_mma_epilogue=t(0x600, True), single_stream_epilogue_no_mma=t(0x800, True),
             pair_full=t(0, False), pair_mainloop=t(0x100, False))


STAGES = {"xrate": stage_xrate, "struct": stage_struct, "small": stage_small, "cand": stage_cand, "big": stage_big, "time": stage_time}

if __name__ == "__main__":
    args = sys.argv[1:]
    if len(args) == 2 and args[0] == "--one":
        STAGES[args[1]]()
        sys.exit(0)
    for st in (args or list(STAGES)):
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", st], timeout=240, capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                emit(stage=st, rc=r.returncode, stderr=r.stderr[-1500:])
        except subprocess.TimeoutExpired:
            emit(stage=st, error="timeout")
