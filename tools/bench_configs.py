"""BASELINE configs 4 and 5 (single GPU): batched inference with GPU post-processing, and the high-resolution
head+loss sweep.  Prints one JSON object; results are recorded in RESULTS.md."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from db_text_minimal_b200 import DBLoss, DBTextModel, SegDetectorRepresenter, _lib, synth


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def config1(s=640):
    """BASELINE config 1: the reference's own CPU-runnable case, 1 x 3 x 640 x 640 eval forward + post-processing (thresh
    0.25, box_thresh 0.5, unclip 1.5), here on the GPU: latency of the model, of `SegDetectorRepresenter.__call__` in both
    modes on a synthetic probability map of that size, and of the two back to back (wall clock, host included)."""
    torch.manual_seed(0)
    model = DBTextModel(pretrained=False).cuda().eval()
    x = synth.images(1, s, s, 0).cuda()
    with torch.no_grad():
        ms_model = timed(lambda: model(x), 20, warm=5)
    P = torch.from_numpy(((synth.prob_map(s, s, 100) - 0.45) * 8).clip(0, 1))[None, None].cuda()
    rep = SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, unclip_ratio=1.5)
    shape = {"shape": [(s, s)]}
    out = {"workload": f"eval forward 1x3x{s}x{s} + post-processing (BASELINE config 1)", "model_ms": ms_model}
    for mode, key in ((False, "post_box_ms"), (True, "post_polygon_ms")):
        for _ in range(3):
            rep(shape, P, is_output_polygon=mode)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            boxes, _ = rep(shape, P, is_output_polygon=mode)
        torch.cuda.synchronize()
        out[key] = (time.perf_counter() - t0) * 1e3 / 20
    t0 = time.perf_counter()
    for _ in range(20):
        with torch.no_grad():
            model(x)
        rep(shape, P, is_output_polygon=False)
    torch.cuda.synchronize()
    out["model_plus_post_box_ms"] = (time.perf_counter() - t0) * 1e3 / 20
    out["img_s"] = 1e3 / out["model_plus_post_box_ms"]
    return out


def config4(n=64, s=1024):
    """BASELINE config 4: batched inference 64 x 1024 x 1024 with GPU binarize (0.25) + connected components + box score
    (0.5) + unclip ratio 1.5.  Model and post-processing are timed separately and together; the post-processing input is a
    synthetic blob map (a randomly initialised network's P sits near 0.5 everywhere: SURVEY section 8c)."""
    torch.manual_seed(0)
    model = DBTextModel(pretrained=False).cuda().eval()
    x = synth.images(n, s, s, 0).cuda()
    ms_model = timed(lambda: model(x), 5)
    maps = np.stack([((synth.prob_map(s, s, 100 + i) - 0.45) * 8).clip(0, 1) for i in range(8)])
    P = torch.from_numpy(np.concatenate([maps] * (n // 8)))[:, None].cuda()
    rep = SegDetectorRepresenter(thresh=0.25, box_thresh=0.5, unclip_ratio=1.5)
    shape = {"shape": [(s, s)] * n}
    for _ in range(2):
        rep(shape, P, is_output_polygon=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        boxes, scores = rep(shape, P, is_output_polygon=False)
    torch.cuda.synchronize()
    ms_boxes = (time.perf_counter() - t0) * 1e3 / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        y = model(x)
        rep(shape, P, is_output_polygon=False)
    torch.cuda.synchronize()
    ms_both = (time.perf_counter() - t0) * 1e3 / reps
    _lib.profile_enable(True)
    rep(shape, P, is_output_polygon=False)
    kern = _lib.profile_report()
    _lib.profile_enable(False)
    _, _, rec, nc = rep.front(P)
    px = n * s * s
    front_ms = sum(k["ms"] for k in kern if k["name"] != "ccl_points")
    dev_ms = sum(k["ms"] for k in kern)
    peak = 6549.8
    try:
        import json as _json
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            peak = _json.load(f)["hbm_gbs"]
    except Exception:
        pass
    return {"workload": f"eval forward {n}x3x{s}x{s} + GPU post-processing (thresh 0.25, box_thresh 0.5, unclip 1.5, box mode)",
            "model_ms": ms_model, "model_img_s": n / ms_model * 1e3,
            "post_boxes_ms": ms_boxes, "post_img_s": n / ms_boxes * 1e3,
            "model_plus_post_ms": ms_both, "e2e_img_s": n / ms_both * 1e3,
            "post_front_device_ms": front_ms, "post_device_ms_incl_border_points": dev_ms,
            "post_front_GBps_9B_per_px": 9 * px / front_ms / 1e6, "post_front_frac_of_hbm": 9 * px / front_ms / 1e6 / peak,
            "candidates_per_image": float(np.mean(nc)),
            "kept_boxes_per_image": float(np.mean([int((b.reshape(len(b), -1) != 0).any(1).sum()) for b in boxes])),
            "post_kernels_ms": {k["name"]: round(k["ms"], 4) for k in kern}}


def config5(n=16, sizes=(640, 768, 896, 1024, 1280, 1536)):
    """head tail (fwd + bwd) + DBLoss(OHEM 'none') fwd + bwd on synthetic ConvT1 outputs."""
    from db_text_minimal_b200 import _ops
    out = []
    for s in sizes:
        h2 = w2 = s // 2
        zt = (torch.randn((n, h2, w2, 128), device="cuda") * 1.5).to(torch.bfloat16)
        gamma, beta = torch.rand(128, device="cuda") + 0.5, torch.randn(128, device="cuda") * 0.3
        w2b, w2t = torch.randn((64, 1, 2, 2), device="cuda") * 0.15, torch.randn((64, 1, 2, 2), device="cuda") * 0.15
        b2b, b2t = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
        gts = torch.from_numpy(synth.gt_maps(n, s, s, 1)).cuda()
        crit = DBLoss(reduction="none")
        px = n * s * s

        def step():
            o, st = _ops.head_tail_fwd(zt, gamma, beta, None, None, True, w2b, w2t, b2b, b2t)
            o.requires_grad_(True)
            l = crit(o, gts)[-1]
            l.backward()
            _ops.head_tail_bwd(zt, gamma, st, w2b, w2t, o.detach(), o.grad)

        ms = timed(step, 5)
        _lib.profile_enable(True)
        step()
        kern = {k["name"]: k["ms"] for k in _lib.profile_report()}
        _lib.profile_enable(False)
        byts = {"head_tail_fwd": 76, "head_tail_bwd_reduce": 88, "head_tail_bwd_apply": 152, "dbloss_reduce": 28,
                "dbloss_select_pass2": 12, "dbloss_bwd": 40}
        dev = sum(kern.get(k, 0.0) for k in byts)
        tot_bytes = sum(byts.values()) * px
        out.append({"size": s, "batch": n, "ms_wall": ms, "ms_kernels": dev, "GBps": tot_bytes / dev / 1e6,
                    "frac_of_hbm_6549.8": tot_bytes / dev / 1e6 / 6549.8,
                    "per_kernel_GBps": {k: byts[k] * px / kern[k] / 1e6 for k in byts if k in kern}})
        del zt, gts
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    res = {}
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "4"):
        res["config4"] = config4()
    if which in ("all", "5"):
        res["config5"] = config5()
    print(json.dumps(res))
