"""N ranks of a partitioned mesh inside ONE process: one context per rank, linked by the library's in-process
communicator (adv_ctx_comm_init_local: device-to-device halo copies ordered by CUDA events), every context
driven by its own host thread.  Runs on any number of GPUs >= 1 (rank r uses GPU r mod device_count), so the
multi-rank path -- boundary-first ordering, interior overlap, identity ranges with skip flags, the K3 pass over
halo nodes, exchange_elem -- is covered on the driver's one-GPU box too."""
from __future__ import annotations

import threading
from typing import List, Optional

import numpy as np
import torch

from common import to_device
from fesom2_b200 import fields as F
from fesom2_b200 import mesh as M


def run_local_ranks(g, part, st, trs, dt, nsteps: int = 1, null_grad: bool = False, tri: Optional[np.ndarray] = None,
                    exchange_inputs: bool = False, tra_diag: bool = False):
    """Returns one dict per rank: dh, dv, values (numpy), owned / all_nodes (global ids, 1-based), launches; with
    tra_diag also tah, tav = the ltra_diag arrays tra_advhoriz / tra_advvert (pre-filled with 7.0)."""
    from fesom2_b200.driver import AdvB200, comm_init_local
    world = int(np.asarray(part).max()) + 1
    ndev = torch.cuda.device_count()
    nb_g = M.nboundary_lay(g)
    ranks = []
    for r in range(world):
        loc = M.localize(g, part, r)
        lst, ltr = F.scatter_to_local(g, loc, st, trs)
        nb = nb_g[loc.myList_nod2D - 1]
        dev = torch.device(f"cuda:{r % ndev}")
        ctx = AdvB200(loc, nb, device=r % ndev, max_tracers=len(ltr))
        if null_grad:
            ctx.set_gradient_mesh(gmesh=M.gradient_mesh(g, tri, part, loc))
        st_d, trs_d = to_device(lst, ltr, dev)
        if null_grad:
            for t in trs_d:
                t.edge_up_dn_grad = None
        if exchange_inputs:                       # wipe the halo of the inputs: exchange_nod must restore it
            for t in trs_d:
                t.values[loc.N:] = -1.0e30
                t.valuesAB[loc.N:] = -1.0e30
        dh = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in ltr]
        dv = [torch.zeros((loc.Nh, loc.L), dtype=torch.float64, device=dev) for _ in ltr]
        tah = [torch.full((loc.Nh, loc.L), 7.0, dtype=torch.float64, device=dev) for _ in ltr] if tra_diag else None
        tav = [torch.full((loc.Nh, loc.L), 7.0, dtype=torch.float64, device=dev) for _ in ltr] if tra_diag else None
        ranks.append(dict(loc=loc, ctx=ctx, st=st_d, trs=trs_d, dh=dh, dv=dv, tah=tah, tav=tav, err=None))
    comm_init_local([rk["ctx"] for rk in ranks])

    def work(rk):
        try:
            ctx, trs_d, dh, dv = rk["ctx"], rk["trs"], rk["dh"], rk["dv"]
            torch.cuda.set_device(ctx.device)
            if exchange_inputs:
                ctx.exchange_nod([t.values for t in trs_d] + [t.valuesAB for t in trs_d], rk["loc"].L)
            ctx.set_state(rk["st"])
            for step in range(nsteps):
                if step > 0:
                    for x in dh + dv:
                        x.zero_()
                    torch.cuda.current_stream().synchronize()
                ctx.do_oce_adv_tra(dt, trs_d, dh, dv, tra_advhoriz=rk["tah"], tra_advvert=rk["tav"])
                if nsteps > 1:
                    ctx.update_values([t.values for t in trs_d], dh, dv)     # + exchange_nod(values)
                    ctx.synchronize()
            ctx.synchronize()
        except Exception as ex:          # surfaced by the caller; the peers time out in their rendezvous
            rk["err"] = ex

    threads = [threading.Thread(target=work, args=(rk,)) for rk in ranks]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for rk in ranks:
        if rk["err"] is not None:
            raise rk["err"]
    out = []
    for rk in ranks:
        loc, ctx = rk["loc"], rk["ctx"]
        res = dict(dh=[x.cpu().numpy() for x in rk["dh"]], dv=[x.cpu().numpy() for x in rk["dv"]],
                   values=[t.values.cpu().numpy() for t in rk["trs"]], owned=loc.myList_nod2D[:loc.N],
                   all_nodes=loc.myList_nod2D, launches=ctx.launch_count, N=loc.N)
        if tra_diag:
            res["tah"] = [x.cpu().numpy() for x in rk["tah"]]
            res["tav"] = [x.cpu().numpy() for x in rk["tav"]]
        if world > 1 and any(t.tra_adv_lim.strip() == "FCT" for t in rk["trs"]):   # a call without limiter exchanges nothing
            res["halo_stats"] = ctx.halo_stats()
        out.append(res)
    for rk in ranks:
        rk["ctx"].close()
    return out
