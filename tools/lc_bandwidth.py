"""Achieved HBM bandwidth of the LinearCombination kernels (src/lc.rs:7-55 -> vec-ode_b200/csrc/lc.cu) on one large state.
Prints one JSON line per operation: algorithmic bytes (vector passes x 8 B per element) / CUDA-event time."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import vecode_b200 as vo

D = 1 << 26  # 512 MB per vector: > L2
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6454.0
for arith in ("strict", "fast"):
    ctx = vo.Context.on_torch_stream(0, arith=arith)
    vs = [vo.Ensemble.wrap_tensor(ctx, torch.rand(D, 1, device="cuda", dtype=torch.float64)) for _ in range(7)]  # one state of D components (N = 1)
    LC = vo.LinearCombination
    ops = {
        "scale (v *= k): 2 passes": (2, lambda: LC.scale(vs[0], 1.0000001)),
        "scalar_multiply_to (t = k v): 2 passes": (2, lambda: LC.scalar_multiply_to(vs[0], 0.5, vs[1])),
        "add_scalar_mul (v = v + k u): 3 passes": (3, lambda: LC.add_scalar_mul(vs[0], 1e-9, vs[1])),
        "add_assign_ref (v += u): 3 passes": (3, lambda: LC.add_assign_ref(vs[0], vs[1])),
        "delta (v -= y): 3 passes": (3, lambda: LC.delta(vs[0], vs[1])),
        "linear_combination, 5 terms: 6 passes (the reference's chain: 14)": (6, lambda: LC.linear_combination(vs[6], vs[1:6], [0.1, 0.2, 0.3, 0.4, 0.5])),
        "stage_combine, 5 terms + x0: 7 passes": (7, lambda: LC.stage_combine(vs[6], vs[1:6], [0.1, 0.2, 0.3, 0.4, 0.5], 1e-3, vs[0])),
    }
    if "--complex" in sys.argv:  # LinearCombination<Complex<f64>, V>: rows of interleaved (re, im) pairs (d = 1, n = D doubles = D/2 elements)
        del vs
        vs = [vo.Ensemble.wrap_tensor(ctx, torch.rand(1, D, device="cuda", dtype=torch.float64)) for _ in range(5)]
        ZC = vo.ComplexLinearCombination
        ops = {
            "complex scale (v *= k): 2 passes": (2, lambda: ZC.scale(vs[0], 1.0000001 + 1e-9j)),
            "complex scalar_multiply_to (t = k v): 2 passes": (2, lambda: ZC.scalar_multiply_to(vs[0], 0.5 - 0.25j, vs[1])),
            "complex add_scalar_mul (v = v + k u): 3 passes": (3, lambda: ZC.add_scalar_mul(vs[0], 1e-9 + 1e-9j, vs[1])),
            "complex linear_combination, 3 terms: 4 passes": (4, lambda: ZC.linear_combination(vs[4], vs[1:4], [0.1 + 0.2j, 0.3j, -0.5])),
        }
    for name, (passes, fn) in ops.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = passes * 8.0 * D / (ms * 1e-3) / 1e9
        print(json.dumps({"op": name, "arith": arith, "elements": D, "ms": round(ms, 4), "GB/s": round(gbs, 1), "frac_of_copy_peak": round(gbs / peak, 3)}), flush=True)
    del vs
