set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -3
python - <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0, "tests")
from fesom2_b200 import mesh as M, fields as F
from fesom2_b200.driver import AdvB200
g = M.synth_mesh(613, 613, nl=71); dev = torch.device("cuda:0")
st = F.make_state(g, dev); dt = F.cfl_dt(g, st, 0.3); tri = F.find_up_downwind_triangles(g, dev)
for lim, hor, ver in (("NON", "MFCT", "QR4C"), ("NON", "MUSCL", "PPM"), ("NON", "UPW1", "UPW1")):
    trs = [F.make_tracers_kind(g, k, dev, tri, hor=hor, ver=ver, lim=lim)[0] for k in range(2)]
    dh = [torch.zeros((g.Nh, g.L), dtype=torch.float64, device=dev) for _ in trs]; dv = [torch.zeros_like(x) for x in dh]
    ctx = AdvB200(g, M.nboundary_lay(g), max_tracers=2)
    for _ in range(3):
        ctx.set_state(st); ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False)
    ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        ctx.set_state(st); ctx.do_oce_adv_tra(dt, trs, dh, dv, sync=False)
    ctx.synchronize(); print(lim, hor, ver, "ms/step", 100 * (time.perf_counter() - t0), flush=True)
    ctx.close()
PY
