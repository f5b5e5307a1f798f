"""Per-kernel timings of libdemf_b200.so at the DeMF geometries (development tool, GPU only).

    python tools/kbench.py [--B 8] [--reps 30] [--only msda,msda_self,encoder,fps,...]

CUDA events on the launching stream; an L2 flush (write of a 256 MB buffer) precedes every
timed launch unless --hot. Prints one JSON line per kernel/config: median and min ms plus the
algorithmic GB/s where one is defined.
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from demf_b200 import _lib, synth  # noqa: E402
from demf_b200.mm import point_ops as ops  # noqa: E402
from demf_b200.mm.ms_deform_attn import MultiScaleDeformableAttnFunction as MSDA  # noqa: E402


def timeit(fn, reps, flush):
    dev = torch.device("cuda")
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if junk is not None:
            junk.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)


def report(name, cfg, ms_med, ms_min, nbytes=None):
    rec = dict(kernel=name, **cfg, ms_median=round(ms_med, 5), ms_min=round(ms_min, 5))
    if nbytes:
        rec["alg_MB"] = round(nbytes / 1e6, 3)
        rec["GBps_median"] = round(nbytes / ms_med / 1e6, 1)
        rec["GBps_best"] = round(nbytes / ms_min / 1e6, 1)
    print(json.dumps(rec), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--hot", action="store_true", help="no L2 flush between launches")
    ap.add_argument("--only", default="")
    ap.add_argument("--sa-lanes", type=int, default=4)
    ap.add_argument("--sa-sleep", type=int, default=0)
    ap.add_argument("--gemm", default="tf32", choices=("tf32", "fp32"), help="library GEMM arithmetic (encoder)")
    ap.add_argument("--profile", action="store_true", help="encoder: print the per-kernel time table")
    args = ap.parse_args()
    only = set(filter(None, args.only.split(",")))
    want = lambda k: not only or k in only  # noqa: E731
    dev = torch.device("cuda:0")
    _lib.load().demf_sa_fused_tune(args.sa_lanes, args.sa_sleep)
    B, flush = args.B, not args.hot
    print(json.dumps(dict(device=torch.cuda.get_device_name(0), B=B, flush_l2=flush)), flush=True)

    if want("msda"):
        for name in ("S512", "REAL", "XL"):
            for P in (4, 2):
                Bm = B if name != "XL" else min(B, 4)
                v, sh, lsi, loc, at = (t.to(dev) for t in synth.make_msda_inputs(B=Bm, name=name, P=P))
                Q, H, D, L = 256, 8, 32, 4
                nbytes = Bm * Q * H * L * P * (4 * D * 4 + 12) + Bm * Q * H * D * 4
                med, mn = timeit(lambda: MSDA.apply(v, sh, lsi, loc, at, 64), args.reps, flush)
                report("msda_fwd", dict(pyramid=name, P=P, B=Bm), med, mn, nbytes)
                go = torch.randn(Bm, Q, H * D, device=dev)
                gv = torch.zeros_like(v)
                gl, ga = torch.empty_like(loc), torch.empty_like(at)
                lib = _lib.load()
                st = torch.cuda.current_stream().cuda_stream

                def bwd():
                    lib.demf_msda_bwd(v.data_ptr(), sh.data_ptr(), lsi.data_ptr(), loc.data_ptr(),
                                      at.data_ptr(), go.data_ptr(), Bm, v.shape[1], H, D, Q, L, P,
                                      gv.data_ptr(), gl.data_ptr(), ga.data_ptr(), st)
                med, mn = timeit(bwd, args.reps, flush)
                report("msda_bwd", dict(pyramid=name, P=P, B=Bm), med, mn, 2 * nbytes)
                del v, gv

    if want("msda_hbm"):
        # HBM-resident regime: XL pyramid at B=8 (2.85 GB of value, far beyond the 126 MB L2), the contract's
        # Q=256 and larger query counts (>= 4 waves of CTAs); inputs generated on the device
        name, Bm, H, D, L, P = "XL", B, 8, 32, 4, 4
        shapes = synth.PYRAMIDS[name]
        S = synth.pyramid_tokens(name)
        g = torch.Generator(device=dev).manual_seed(11)
        v = torch.randn(Bm, S, H, D, generator=g, device=dev)
        sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
        lsi = torch.cat([sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]])
        for Q in (256, 1024, 4096):
            ref = torch.rand(Bm, Q, 1, 1, 1, 2, generator=g, device=dev)
            loc = (ref + 0.05 * torch.randn(Bm, Q, H, L, P, 2, generator=g, device=dev)).contiguous()
            at = torch.softmax(torch.randn(Bm, Q, H, L * P, generator=g, device=dev), -1).view(Bm, Q, H, L, P).contiguous()
            nbytes = Bm * Q * H * L * P * (4 * D * 4 + 12) + Bm * Q * H * D * 4
            med, mn = timeit(lambda: MSDA.apply(v, sh, lsi, loc, at, 64), args.reps, flush)
            report("msda_fwd_hbm", dict(pyramid=name, P=P, B=Bm, Q=Q, value_GB=round(v.numel() * 4 / 1e9, 2)), med, mn, nbytes)
        del v

    if want("imgbranch"):
        from demf_b200 import engine
        torch.manual_seed(0)
        engine.set_gemm_precision(args.gemm)
        model = engine.build_demf_votenet(num_points=4, img_branch=True).to(dev).eval()
        img = torch.randn(B, 3, 512, 512, device=dev)
        metas = [dict(img_shape=(512, 512, 3), batch_input_shape=(512, 512)) for _ in range(B)]
        with torch.no_grad():
            for fn, nm in ((lambda: model.img_neck(model.img_backbone(img)), "resnet50+channelmapper"),
                           (lambda: model.extract_img_feat(img, metas), "image branch (backbone+neck+encoder)")):
                med, mn = timeit(fn, max(5, args.reps // 3), False)
                report(nm, dict(B=B, image="512x512", gemm=args.gemm, images_per_s=round(B / med * 1e3, 1)), med, mn)
        del model

    if want("gemm"):
        # the training GEMMs at batch-4 shapes (configs[3]): own tcgen05/TMA kernels vs the library (TF32)
        torch.backends.cuda.matmul.allow_tf32 = True
        shapes = [("SA1.l0", 524288, 8, 64), ("SA1.l1", 524288, 64, 64), ("SA1.l2", 524288, 64, 128),
                  ("SA2.l0", 131072, 132, 128), ("SA2.l1", 131072, 128, 128), ("SA2.l2", 131072, 128, 256),
                  ("SA3.l0", 32768, 260, 128), ("SA3.l2", 32768, 128, 256), ("agg.l0", 16384, 260, 256),
                  ("agg.l1", 16384, 256, 256), ("FP.l0", 4096, 512, 256), ("vote.l0", 4096, 256, 256)]
        only_layers = set(filter(None, os.environ.get("GEMM_LAYERS", "").split(",")))
        _lib.load().demf_gemm_tune(int(os.environ.get("GEMM_EPI", "2")))
        for name, R, K, N in shapes:
            if only_layers and name not in only_layers:
                continue
            x = torch.randn(R, K, device=dev)
            w = torch.randn(N, K, device=dev) / K ** 0.5
            dy = torch.randn(R, N, device=dev)
            dw = torch.zeros(N, K, device=dev)
            y = torch.empty(R, N, device=dev)
            dx = torch.empty(R, K, device=dev)
            st = ops.bn_rows_state(256, dev)
            st2 = ops.bn_rows_state(256, dev)
            bn_m, bn_i = torch.zeros(K, device=dev), torch.ones(K, device=dev)
            bn_g, bn_b = torch.ones(K, device=dev), torch.zeros(K, device=dev)
            nb = R * (K + N) * 4
            for label, fn in (("fwd", lambda: ops.gemm_rows_fwd(x, w, out=y)),
                              ("fwd+stats", lambda: ops.gemm_rows_fwd(x, w, bn_state=st, out=y)),
                              ("fwd torch", lambda: torch.mm(x, w.t(), out=y)),
                              ("dgrad", lambda: ops.gemm_rows_dgrad(dy, w, out=dx)),
                              ("dgrad+bn", (lambda: ops.gemm_rows_dgrad_bn(dy, w, x, bn_m, bn_i, bn_g, bn_b, st2))
                               if K in (64, 128, 256) else None),
                              ("dgrad torch", lambda: torch.mm(dy, w, out=dx)),
                              ("wgrad", lambda: ops.gemm_wgrad_(dw, dy, x)),
                              ("wgrad torch", lambda: dw.addmm_(dy.t(), x))):
                if fn is None:
                    continue
                med, mn = timeit(fn, args.reps, flush)
                report("gemm " + label, dict(layer=name, R=R, K=K, N=N), med, mn, nb)
        print(json.dumps(dict(gemm_error=ops.gemm_error())), flush=True)

    if want("msda_self"):
        # encoder regime (demf/modeling/layers/deform_detr_encoder.py): every pixel of the pyramid is a
        # query (Q = S) sampling around its own position. Compulsory HBM bytes = value + locations +
        # weights read once, output written once; `sample_MB` is what the gathers request from L1/L2.
        for name in ("S512", "REAL", "XL"):
            Bm = B if name != "XL" else 1
            shapes = synth.PYRAMIDS[name]
            S, H, D, L, P = synth.pyramid_tokens(name), 8, 32, 4, 4
            g = torch.Generator().manual_seed(5)
            v = torch.randn(Bm, S, H, D, generator=g).to(dev)
            ref = torch.cat([torch.stack(torch.meshgrid(
                (torch.arange(w) + 0.5) / w, (torch.arange(h) + 0.5) / h, indexing="xy"), -1).reshape(-1, 2)
                for h, w in shapes], 0)
            loc = (ref[None, :, None, None, None, :] + 0.03 * torch.randn(Bm, S, H, L, P, 2, generator=g)).to(dev)
            at = torch.softmax(torch.randn(Bm, S, H, L * P, generator=g), -1).view(Bm, S, H, L, P).to(dev)
            sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
            lsi = torch.cat([sh.new_zeros(1), sh.prod(1).cumsum(0)[:-1]])
            nbytes = v.numel() * 4 * 2 + loc.numel() * 4 + at.numel() * 4
            med, mn = timeit(lambda: MSDA.apply(v, sh, lsi, loc, at, 64), args.reps, flush)
            report("msda_self_fwd", dict(pyramid=name, B=Bm, Q=S, sample_MB=round(
                Bm * S * H * L * P * 4 * D * 4 / 1e6, 1)), med, mn, nbytes)
            del v, loc, at

    if want("encoder"):
        from demf_b200 import engine
        from demf_b200.mm.config import Config
        from demf_b200.mm.registry import build_head
        torch.manual_seed(0)
        engine.set_gemm_precision(args.gemm)
        enc = build_head(Config.fromfile(engine.CONFIG).img_encoder_cfg.to_dict())
        enc.init_weights()
        enc = enc.to(dev).eval()
        for name in ("S512", "REAL"):
            feats = [f.to(dev) for f in synth.make_pyramid(B, name)]
            metas = synth.make_img_metas(B, name)
            if args.profile:
                from torch.profiler import ProfilerActivity, profile
                with torch.no_grad():
                    enc(feats, metas)
                    torch.cuda.synchronize()
                    with profile(activities=[ProfilerActivity.CUDA]) as prof:
                        enc(feats, metas)
                        torch.cuda.synchronize()
                print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
            with torch.no_grad():
                med, mn = timeit(lambda: enc(feats, metas), max(5, args.reps // 3), False)
                report("img_encoder_eager", dict(pyramid=name, B=B, S=synth.pyramid_tokens(name), gemm=args.gemm), med, mn)
                side = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    for _ in range(3):
                        enc(feats, metas)
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    out = enc(feats, metas)  # noqa: F841
                med, mn = timeit(graph.replay, max(5, args.reps // 3), False)
                report("img_encoder_graph", dict(pyramid=name, B=B, S=synth.pyramid_tokens(name),
                                                 images_per_s=round(B / med * 1e3, 1)), med, mn)
            del graph, out

    pts = synth.make_points(B, 20000, seed=0, clustered=True)[..., :3].contiguous().to(dev)
    geoms = [(20000, 2048, 0.2, 64, 1), (2048, 1024, 0.4, 32, 128), (1024, 512, 0.8, 16, 256),
             (512, 256, 1.2, 16, 256), (1024, 256, 0.3, 16, 256)]
    if want("fps"):
        for N, m in ((20000, 2048), (2048, 1024), (1024, 512), (512, 256), (1024, 256)):
            x = pts[:, :N].contiguous()
            med, mn = timeit(lambda: ops.furthest_point_sample(x, m), max(5, args.reps // 3), False)
            report("fps", dict(B=B, N=N, m=m, us_per_iter=round(med * 1e3 / (m - 1), 3)), med, mn)
            if N >= 2048:
                grid = ops.ball_grid(x, 0.2 if N > 4096 else 0.4)
                med, mn = timeit(lambda: ops.furthest_point_sample_grid(x, m, grid), max(5, args.reps // 3), False)
                report("fps_grid", dict(B=B, N=N, m=m, us_per_iter=round(med * 1e3 / (m - 1), 3)), med, mn)
    if want("ball"):
        for N, M, r, ns, C in geoms:
            x = pts[:, :N].contiguous()
            c = x[:, :M].contiguous()
            med, mn = timeit(lambda: ops.ball_query(0.0, r, ns, x, c), args.reps, flush)
            report("ball_query", dict(B=B, N=N, M=M, ns=ns), med, mn, B * (N * 12 + M * 12 + M * ns * 4))
            f = torch.randn(B, C, N, device=dev)
            qg = ops.QueryAndGroup(r, ns, use_xyz=True, normalize_xyz=True)
            med, mn = timeit(lambda: qg(x, c, f), args.reps, flush)
            report("query_and_group", dict(B=B, N=N, M=M, ns=ns, C=C), med, mn,
                   B * (N * 12 + M * 12 + M * ns * 4 + C * N * 4 + (C + 3) * M * ns * 4))
    if want("sa"):
        # one set-abstraction level in eval mode: fused tcgen05 kernel vs the unfused rows path
        # (group rows kernel + 3 cuBLASLt TF32 GEMMs with bias+ReLU epilogue + amax)
        torch.backends.cuda.matmul.allow_tf32 = True
        widths_of = {1: (64, 64, 128), 128: (128, 128, 256), 256: (128, 128, 256)}
        cur = pts
        for li, (N, M, r, ns, C) in enumerate(geoms):
            widths = (256, 256, 256) if li == 4 else widths_of[C]
            x = pts[:, :N].contiguous() if li in (0, 4) else cur[:, :N].contiguous()
            c = ops.gather_rows(x, ops.furthest_point_sample(x, M)).contiguous()
            cur = c
            f = torch.randn(B, N, C, device=dev)
            K = ops.group_rows_width(C)
            cin = [K] + list(widths[:-1])
            ws = [torch.randn(co, ci, device=dev) / ci ** 0.5 for co, ci in zip(widths, cin)]
            bs = [torch.randn(co, device=dev) * 0.1 for co in widths]
            wpack, bias, wd = ops.sa_pack_mlp(ws, bs)
            rows = B * M * ns
            flops = 2.0 * rows * sum(co * ci for co, ci in zip(widths, cin))
            hbm = B * (N * 12 + N * C * 4 + M * 12 + M * widths[2] * 4)
            med, mn = timeit(lambda: ops.sa_fused(x, c, f, 0.0, r, ns, True, wpack, bias, wd),
                             args.reps, flush)
            report("sa_fused", dict(B=B, N=N, M=M, ns=ns, C=C, widths=widths,
                                    TFLOPs_median=round(flops / med / 1e9, 1)), med, mn, hbm)
            if N >= 2048:
                med, mn = timeit(lambda: ops.ball_grid(x, r), args.reps, flush)
                report("ball_grid_build", dict(B=B, N=N, r=r), med, mn, B * N * 28)
                grid = ops.ball_grid(x, r)
                med, mn = timeit(lambda: ops.sa_fused(x, c, f, 0.0, r, ns, True, wpack, bias, wd, grid=grid),
                                 args.reps, flush)
                report("sa_fused+grid", dict(B=B, N=N, M=M, ns=ns, C=C, widths=widths,
                                             TFLOPs_median=round(flops / med / 1e9, 1)), med, mn, hbm)
                med, mn = timeit(lambda: ops.ball_query_grid(0.0, r, ns, x, c, grid), args.reps, flush)
                report("ball_query_grid", dict(B=B, N=N, M=M, ns=ns), med, mn)
                med, mn = timeit(lambda: ops.ball_query(0.0, r, ns, x, c), args.reps, flush)
                report("ball_query_scan", dict(B=B, N=N, M=M, ns=ns), med, mn)

            def unfused():
                _, g = ops.query_and_group_rows(x, c, f, 0.0, r, ns, True)
                y = g.view(rows, K)
                for w, b in zip(ws, bs):
                    y = torch._addmm_activation(b, y, w.t())
                return y.view(B, M, ns, -1).amax(dim=2)
            with torch.no_grad():
                med, mn = timeit(unfused, args.reps, flush)
            report("sa_unfused", dict(B=B, N=N, M=M, ns=ns, C=C, widths=widths,
                                      TFLOPs_median=round(flops / med / 1e9, 1)), med, mn, hbm)
    if want("interp"):
        for n, m in ((512, 256), (1024, 512)):
            a, b = pts[:, :n].contiguous(), pts[:, 5000:5000 + m].contiguous()
            med, mn = timeit(lambda: ops.three_nn(a, b), args.reps, flush)
            report("three_nn", dict(B=B, n=n, m=m), med, mn)
            dist, idx = ops.three_nn(a, b)
            w = (1.0 / (dist + 1e-8))
            w = (w / w.sum(2, keepdim=True)).contiguous()
            f = torch.randn(B, 256, m, device=dev)
            med, mn = timeit(lambda: ops.three_interpolate(f, idx, w), args.reps, flush)
            report("three_interpolate", dict(B=B, n=n, m=m, C=256), med, mn,
                   B * 256 * (n + m) * 4 + B * n * 24)
    print(json.dumps(dict(launches=_lib.launch_count())))


if __name__ == "__main__":
    main()
