#!/usr/bin/env python
"""tools/exp_variants.py -- A/B the kernel variants on the bench workload in ONE process.

The library reads its tuning knobs (ADV_PIPE, ADV_E1_NG, ADV_E1_D, ADV_ND_NG, ...) from the
environment in adv_ctx_create, so every variant is a fresh context over the same device-resident
inputs.  Prints one JSON line per variant: ms per step and per kernel (CUDA events of the library's
phase markers), plus a checksum of the tendencies so that a variant that changes the result is
visible at once (all variants must print the same checksum).

    python tools/exp_variants.py [--side 613] [--steps 10] "ADV_PIPE=0" "ADV_PIPE=1 ADV_E1_D=2" ...
"""
import argparse
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=613)
    ap.add_argument("--nl", type=int, default=71)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tracers", type=int, default=2)
    ap.add_argument("--null-grad", action="store_true", help="edge_up_dn_grad = NULL: the library computes the gradients inside the step")
    ap.add_argument("--tra-diag", action="store_true", help="ltra_diag = .true.: tra_advhoriz / tra_advvert are produced as well")
    ap.add_argument("variants", nargs="*")
    a = ap.parse_args()
    import torch
    from fesom2_b200 import fields as F, mesh as M
    from fesom2_b200.driver import AdvB200
    dev = torch.device("cuda:0")
    g = M.synth_mesh(a.side, a.side, nl=a.nl)
    nb = M.nboundary_lay(g)
    st = F.make_state(g, dev)
    dt = F.cfl_dt(g, st, 0.3)
    tri = F.find_up_downwind_triangles(g)
    trs = []
    for k in range(a.tracers):
        trs += F.make_tracers_kind(g, k, dev, tri, hor="MFCT", ver="QR4C", lim="FCT")
    dh = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    dg = dict(tra_advhoriz=[torch.zeros_like(x) for x in dh], tra_advvert=[torch.zeros_like(x) for x in dv]) if a.tra_diag else {}
    knobs = set()
    for v in a.variants:
        for kv in v.split():
            knobs.add(kv.split("=")[0])
    for v in a.variants or [""]:
        for k in knobs:
            os.environ.pop(k, None)
        for kv in v.split():
            k, val = kv.split("=")
            os.environ[k] = val
        try:
            ctx = AdvB200(g, nb, device=0, max_tracers=a.tracers)
            if a.null_grad:
                ctx.set_gradient_mesh(tri)
                for t in trs:
                    t.edge_up_dn_grad = None
            for x in dh + dv:
                x.zero_()
            ctx.set_state(st)
            ctx.do_oce_adv_tra(dt, trs, dh, dv, **dg)       # one checked step from zero tendencies
            h = hashlib.sha1()
            for x in dh + dv:
                h.update(x.cpu().numpy().tobytes())
            for _ in range(3):
                ctx.set_state(st)
                ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False, **dg)
            ctx.synchronize()
            ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(ext)
            for _ in range(a.steps):
                ctx.set_state(st)
                ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False, **dg)
            ev1.record(ext)
            ctx.synchronize()
            torch.cuda.synchronize()
            ms = ev0.elapsed_time(ev1) / a.steps
            ctx.set_profiling(True)
            ph = np.zeros(4)
            try:
                for _ in range(5):
                    ctx.set_state(st)
                    ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False, **dg)
                    ctx.synchronize()
                    ph += np.array(ctx.phase_ms()[:4])
                ph /= 5
            except Exception:        # interleaved schedules (ADV_BAND) have no per-phase markers
                ph[:] = 0.0
            ctx.close()
            print(json.dumps({"variant": v, "ms_per_step": round(ms, 4), "E1": round(float(ph[0]), 4), "N1": round(float(ph[1]), 4),
                              "K2": round(float(ph[2]), 4), "K3": round(float(ph[3]), 4), "sha1": h.hexdigest()[:12]}), flush=True)
        except Exception as ex:
            print(json.dumps({"variant": v, "error": str(ex)[:300]}), flush=True)


if __name__ == "__main__":
    main()
