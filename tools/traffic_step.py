#!/usr/bin/env python
"""tools/traffic_step.py -- run ONE step of the bench workload between cudaProfilerStart/Stop so that

    ncu --profile-from-start off --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file out.csv \
        python tools/traffic_step.py [--side 613] [KEY=VAL ...]

lists the DRAM bytes of every launch of that step with the caches left as the previous launches
left them (the L2-resident band pipeline is only visible that way).  tools/traffic_sum.py adds them up."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--side", type=int, default=613)
    ap.add_argument("--nl", type=int, default=71)
    ap.add_argument("--tracers", type=int, default=2)
    ap.add_argument("--warm", type=int, default=2)
    ap.add_argument("--null-grad", action="store_true", help="edge_up_dn_grad = NULL: gradients computed inside the call")
    ap.add_argument("env", nargs="*")
    a = ap.parse_args()
    for kv in a.env:
        k, v = kv.split("=")
        os.environ[k] = v
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
    ctx = AdvB200(g, nb, device=0, max_tracers=a.tracers)
    if a.null_grad:
        ctx.set_gradient_mesh(tri)
        for t in trs:
            t.edge_up_dn_grad = None
    for _ in range(a.warm):
        ctx.set_state(st)
        ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False)
    ctx.synchronize()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ctx.set_state(st)
    ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False)
    ctx.synchronize()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ctx.close()


if __name__ == "__main__":
    main()
