import sys, os, numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
from common import make_case, run_oracle, to_device
from fesom2_b200 import mesh as M
from fesom2_b200.driver import AdvB200
g = M.synth_mesh(31, 27, nl=20, min_layers=4)
mode = sys.argv[1] if len(sys.argv) > 1 else "plain"
cases = [("UPW1", "UPW1", "FCT"), ("UPW1", "QR4C", "NON"), ("UPW1", "UPW1", "FCT"), ("MFCT", "QR4C", "FCT"), ("UPW1", "UPW1", "FCT")]
for hor, ver, lim in cases:
    st, trs, nb, dt = make_case(g, 2, hor, ver, lim, ph=0.25, pv=0.75)
    ora = run_oracle(g, st, trs, nb, dt)
    dev = torch.device("cuda:0")
    ctx = AdvB200(g, nb, device=0, max_tracers=2)
    st_d, trs_d = to_device(st, trs, dev)
    dh = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    dv = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]
    if mode == "sync":
        torch.cuda.synchronize()
    ctx.set_state(st_d)
    ctx.do_oce_adv_tra(dt, trs_d, dh, dv)
    f = ctx.get_work("adv_flux_hor", 0)
    print(mode, hor, ver, lim, "adf_h", np.abs(f).max(), "dh err", np.abs(dh[0].cpu().numpy() - ora.dttf_h[0]).max(),
          "dv err", np.abs(dv[0].cpu().numpy() - ora.dttf_v[0]).max(), flush=True)
    ctx.close()
